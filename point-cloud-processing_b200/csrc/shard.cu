// Multi-GPU plumbing on the device: the boundary strips a slab sends to its
// neighbours are cut out of the resident cloud by one ordered compaction instead of a chain of
// framework masking kernels with a host round trip.  Order = input order, so the local cloud a
// rank assembles (and with it every tie-break by original index) is reproducible.
#include "api_util.hpp"

using namespace pcpx;

namespace {

constexpr int kB     = 256;
constexpr int kItems = 16; // points per thread
constexpr int kTile  = kB * kItems;

__device__ __forceinline__ uint32_t classify(const float* __restrict__ xyz, uint32_t stride_f,
                                             uint32_t i, uint32_t n, int axis, float below,
                                             float above)
{
    if (i >= n)
        return 0u;
    float const c = __ldg(xyz + (size_t)i * stride_f + axis);
    return (c < below ? 1u : 0u) | (c > above ? 2u : 0u);
}

// per tile: how many points go to the "below" strip and to the "above" strip
__global__ void __launch_bounds__(kB) band_count_kernel(const float* __restrict__ xyz,
                                                        uint32_t stride_f, uint32_t n, int axis,
                                                        float below, float above,
                                                        uint32_t* __restrict__ tile_counts)
{
    uint32_t const base = blockIdx.x * kTile + threadIdx.x * kItems;
    uint32_t lo = 0, hi = 0;
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const f = classify(xyz, stride_f, base + r, n, axis, below, above);
        lo += f & 1u, hi += f >> 1;
    }
    __shared__ uint32_t s_lo, s_hi;
    if (threadIdx.x == 0)
        s_lo = 0, s_hi = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1)
        lo += __shfl_xor_sync(0xFFFFFFFFu, lo, o), hi += __shfl_xor_sync(0xFFFFFFFFu, hi, o);
    if ((threadIdx.x & 31) == 0)
        atomicAdd(&s_lo, lo), atomicAdd(&s_hi, hi);
    __syncthreads();
    if (threadIdx.x == 0)
        tile_counts[2 * blockIdx.x] = s_lo, tile_counts[2 * blockIdx.x + 1] = s_hi;
}

// exclusive scan of the tile counts by one block; totals -> counts_out[0..1] (64-bit)
__global__ void __launch_bounds__(1024) band_scan_kernel(uint32_t* tile_counts, uint32_t n_tiles,
                                                         unsigned long long* counts_out)
{
    __shared__ uint32_t warp_lo[32], warp_hi[32];
    __shared__ uint32_t carry_lo, carry_hi;
    if (threadIdx.x == 0)
        carry_lo = 0, carry_hi = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024)
    {
        uint32_t const i  = base + threadIdx.x;
        uint32_t const xl = i < n_tiles ? tile_counts[2 * i] : 0u;
        uint32_t const xh = i < n_tiles ? tile_counts[2 * i + 1] : 0u;
        uint32_t il = xl, ih = xh;
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const yl = __shfl_up_sync(0xFFFFFFFFu, il, o);
            uint32_t const yh = __shfl_up_sync(0xFFFFFFFFu, ih, o);
            if ((threadIdx.x & 31) >= o)
                il += yl, ih += yh;
        }
        if ((threadIdx.x & 31) == 31)
            warp_lo[threadIdx.x >> 5] = il, warp_hi[threadIdx.x >> 5] = ih;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            uint32_t wl = warp_lo[threadIdx.x], wh = warp_hi[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t const yl = __shfl_up_sync(0xFFFFFFFFu, wl, o);
                uint32_t const yh = __shfl_up_sync(0xFFFFFFFFu, wh, o);
                if (threadIdx.x >= o)
                    wl += yl, wh += yh;
            }
            warp_lo[threadIdx.x] = wl, warp_hi[threadIdx.x] = wh;
        }
        __syncthreads();
        uint32_t const w = threadIdx.x >> 5;
        if (i < n_tiles)
        {
            tile_counts[2 * i]     = carry_lo + (w ? warp_lo[w - 1] : 0u) + il - xl;
            tile_counts[2 * i + 1] = carry_hi + (w ? warp_hi[w - 1] : 0u) + ih - xh;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            carry_lo += warp_lo[31], carry_hi += warp_hi[31];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        counts_out[0] = carry_lo, counts_out[1] = carry_hi;
}

// ordered scatter: a thread owns kItems consecutive points, so its rank inside the tile is the
// exclusive prefix of the per-thread counts
__global__ void __launch_bounds__(kB) band_scatter_kernel(const float* __restrict__ xyz,
                                                          uint32_t stride_f, uint32_t n, int axis,
                                                          float below, float above,
                                                          const uint32_t* __restrict__ tile_offsets,
                                                          float* __restrict__ out_below,
                                                          float* __restrict__ out_above,
                                                          uint32_t capacity)
{
    uint32_t const base = blockIdx.x * kTile + threadIdx.x * kItems;
    uint32_t flags = 0, lo = 0, hi = 0;
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const f = classify(xyz, stride_f, base + r, n, axis, below, above);
        flags |= f << (2 * r);
        lo += f & 1u, hi += f >> 1;
    }
    // block-wide exclusive prefix of (lo, hi)
    __shared__ uint32_t warp_lo[kB / 32], warp_hi[kB / 32];
    uint32_t il = lo, ih = hi;
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const yl = __shfl_up_sync(0xFFFFFFFFu, il, o);
        uint32_t const yh = __shfl_up_sync(0xFFFFFFFFu, ih, o);
        if ((threadIdx.x & 31) >= o)
            il += yl, ih += yh;
    }
    if ((threadIdx.x & 31) == 31)
        warp_lo[threadIdx.x >> 5] = il, warp_hi[threadIdx.x >> 5] = ih;
    __syncthreads();
    uint32_t pl = tile_offsets[2 * blockIdx.x] + il - lo;
    uint32_t ph = tile_offsets[2 * blockIdx.x + 1] + ih - hi;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w)
        pl += warp_lo[w], ph += warp_hi[w];
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const f = (flags >> (2 * r)) & 3u;
        if (!f)
            continue;
        const float* src = xyz + (size_t)(base + r) * stride_f;
        float const x = src[0], y = src[1], z = src[2];
        if (f & 1u)
        {
            if (pl < capacity)
                out_below[3 * (size_t)pl] = x, out_below[3 * (size_t)pl + 1] = y,
                                       out_below[3 * (size_t)pl + 2] = z;
            ++pl;
        }
        if (f & 2u)
        {
            if (ph < capacity)
                out_above[3 * (size_t)ph] = x, out_above[3 * (size_t)ph + 1] = y,
                                       out_above[3 * (size_t)ph + 2] = z;
            ++ph;
        }
    }
}

} // namespace

extern "C" int pcpx_extract_bands(const float* xyz, size_t n, size_t stride_bytes, int axis,
                                  float below, float above, float* out_below, float* out_above,
                                  size_t capacity, uint64_t* counts, void* cuda_stream)
{
    return guarded([&] {
        if (!counts)
            fail(PCPX_ERR_INVALID_ARG, "counts is NULL");
        if (axis < 0 || axis > 2)
            fail(PCPX_ERR_INVALID_ARG, "axis must be 0, 1 or 2");
        if (n >= 0xFFFFFFFFull || capacity >= 0xFFFFFFFFull)
            fail(PCPX_ERR_UNSUPPORTED, "more than 2^32 - 2 points in one call");
        if (stride_bytes == 0)
            stride_bytes = 12;
        if (stride_bytes < 12 || stride_bytes % 4)
            fail(PCPX_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and >= 12");
        if (n && (!xyz || !is_device_pointer(xyz)))
            fail(PCPX_ERR_INVALID_ARG, "xyz must be device memory (this call is multi-GPU plumbing)");
        if (capacity && (!out_below || !out_above || !is_device_pointer(out_below) ||
                         !is_device_pointer(out_above)))
            fail(PCPX_ERR_INVALID_ARG, "out_below / out_above must be device memory");
        bool const counts_on_device = is_device_pointer(counts);
        cudaStream_t const s        = static_cast<cudaStream_t>(cuda_stream);
        uint32_t const n_tiles      = (uint32_t)((n + kTile - 1) / kTile);
        DevBuf<uint32_t> tiles(2 * (size_t)std::max(n_tiles, 1u));
        DevBuf<unsigned long long> totals;
        unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(counts);
        if (!counts_on_device)
        {
            totals.alloc(2);
            d_counts = totals.get();
        }
        uint32_t const sf = (uint32_t)(stride_bytes / 4);
        if (n_tiles)
            band_count_kernel<<<n_tiles, kB, 0, s>>>(xyz, sf, (uint32_t)n, axis, below, above,
                                                     tiles.get());
        band_scan_kernel<<<1, 1024, 0, s>>>(tiles.get(), n_tiles, d_counts);
        if (n_tiles && capacity)
            band_scatter_kernel<<<n_tiles, kB, 0, s>>>(xyz, sf, (uint32_t)n, axis, below, above,
                                                       tiles.get(), out_below, out_above,
                                                       (uint32_t)capacity);
        PCPX_CHECK_LAUNCH();
        if (!counts_on_device)
            PCPX_CUDA(cudaMemcpyAsync(counts, d_counts, 16, cudaMemcpyDeviceToHost, s));
        // the tile buffer goes back to the pool on return: the stream must be done with it
        PCPX_CUDA(cudaStreamSynchronize(s));
    });
}
