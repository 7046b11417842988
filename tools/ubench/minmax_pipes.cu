// Micro-benchmark: which pipes do the min/max flavours of sm_100a issue on, and do they overlap?
// Each kernel runs a long dependent-free stream of one instruction flavour (8 independent chains
// per thread), 1024 threads per SM-resident block, and reports warp-instructions per clock and SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_pipes minmax_pipes.cu && ./minmax_pipes
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

constexpr int ITER = 4096, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, uint32_t seed, long long* clk)
{
    uint32_t a[CH], b[CH];
    for (int i = 0; i < CH; ++i)
        a[i] = seed * (threadIdx.x + 1) + i * 977u, b[i] = seed ^ (i * 7919u + threadIdx.x);
    float fa[CH], fb[CH];
    for (int i = 0; i < CH; ++i)
        fa[i] = __uint_as_float((a[i] & 0x3FFFFFFFu) | 0x20000000u), fb[i] = __uint_as_float((b[i] & 0x3FFFFFFFu) | 0x20000000u);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
#pragma unroll
        for (int i = 0; i < CH; ++i)
        {
            if (MODE == 0) // VIMNMX.U32
                a[i] = min(a[i], b[i]), b[i] = max(b[i], a[(i + 1) % CH]);
            if (MODE == 1) // FMNMX
                fa[i] = fminf(fa[i], fb[i]), fb[i] = fmaxf(fb[i], fa[(i + 1) % CH]);
            if (MODE == 2) // HMNMX2.BF16
            {
                __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a[i]), y = *reinterpret_cast<__nv_bfloat162*>(&b[i]);
                __nv_bfloat162 z = *reinterpret_cast<__nv_bfloat162*>(&a[(i + 1) % CH]);
                x = __hmin2(x, y), y = __hmax2(y, z);
                a[i] = *reinterpret_cast<uint32_t*>(&x), b[i] = *reinterpret_cast<uint32_t*>(&y);
            }
            if (MODE == 3) // VIMNMX3.U32
                a[i] = __vimax3_u32(a[i], b[i], a[(i + 1) % CH]), b[i] = __vimin3_u32(b[i], a[i], b[(i + 1) % CH]);
            if (MODE == 4) // u16x2 min/max
                a[i] = __vminu2(a[i], b[i]), b[i] = __vmaxu2(b[i], a[(i + 1) % CH]);
            if (MODE == 5) // mix: VIMNMX.U32 + HMNMX2.BF16 (do they overlap?)
            {
                a[i] = min(a[i], a[(i + 1) % CH]);
                __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b[i]), z = *reinterpret_cast<__nv_bfloat162*>(&b[(i + 1) % CH]);
                y = __hmax2(y, z);
                b[i] = *reinterpret_cast<uint32_t*>(&y);
            }
            if (MODE == 6) // mix: VIMNMX.U32 + FMUL (different pipes: should overlap)
                a[i] = min(a[i], a[(i + 1) % CH]), fa[i] = fa[i] * fb[i];
            if (MODE == 7) // FMUL only
                fa[i] = fa[i] * fb[i], fb[i] = fb[i] * fa[(i + 1) % CH];
            if (MODE == 8) // HMNMX2 f16
            {
                __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&b[i]);
                __half2 z = *reinterpret_cast<__half2*>(&a[(i + 1) % CH]);
                x = __hmin2(x, y), y = __hmax2(y, z);
                a[i] = *reinterpret_cast<uint32_t*>(&x), b[i] = *reinterpret_cast<uint32_t*>(&y);
            }
            if (MODE == 9) // __vimax3_u16x2
                a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) % CH]), b[i] = __vimin3_u16x2(b[i], a[i], b[(i + 1) % CH]);
            if (MODE == 10) // FMNMX + HMNMX2.BF16 mix
            {
                fa[i] = fminf(fa[i], fa[(i + 1) % CH]);
                __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b[i]), z = *reinterpret_cast<__nv_bfloat162*>(&b[(i + 1) % CH]);
                y = __hmax2(y, z);
                b[i] = *reinterpret_cast<uint32_t*>(&y);
            }
            if (MODE == 11) // ISETP + SEL pair (predicate compare + select)
                a[i] = a[i] < b[i] ? a[(i + 1) % CH] : b[(i + 3) % CH], b[i] = b[i] + 1;
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < CH; ++i)
        s += a[i] + b[i] + __float_as_uint(fa[i]) + __float_as_uint(fb[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0)
        clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter)
{
    uint32_t* out;
    long long* clk;
    int nb = 148;
    cudaMalloc(&out, nb * 1024 * 4);
    cudaMalloc(&clk, nb * 8);
    bench<MODE><<<nb, 1024>>>(out, 12345u, clk);
    bench<MODE><<<nb, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, nb * 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    // 32 warps per SM = 8 per SMSP
    double winst = 8.0 * ITER * CH * ops_per_iter;
    printf("%-28s %8.0f cycles  %.3f warp-inst/clk/SMSP\n", name, cyc, winst / cyc);
    cudaFree(out), cudaFree(clk);
}

int main()
{
    run<0>("VIMNMX.U32", 2);
    run<1>("FMNMX", 2);
    run<2>("HMNMX2.BF16", 2);
    run<8>("HMNMX2.F16", 2);
    run<3>("VIMNMX3.U32", 2);
    run<4>("vminu2/vmaxu2", 2);
    run<9>("vimax3_u16x2", 2);
    run<5>("VIMNMX.U32 + HMNMX2.BF16", 2);
    run<10>("FMNMX + HMNMX2.BF16", 2);
    run<6>("VIMNMX.U32 + FMUL", 2);
    run<7>("FMUL", 2);
    run<11>("ISETP+SEL (+IADD)", 3);
    return 0;
}
