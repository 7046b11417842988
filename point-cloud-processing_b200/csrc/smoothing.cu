// The radius-search callers of the reference as device loops: bilateral
// filter of points / normals (algorithm/bilateral_filter.hpp) and WLOP (algorithm/wlop.hpp).
// Where the reference rebuilds a kd-tree per iteration and runs a `par` transform of range
// searches, this rebuilds the GPU index (index.cu, ~1 ms per 10 M points) and launches one
// thread-per-point kernel whose body is smoothing_core.cuh.  Everything stays on the device
// between iterations.  No CPU path.
#include <numeric>
#include <random>

#include "api_util.hpp"
#include "smoothing_core.cuh"

using namespace pcpx;

namespace {

constexpr int kBlock = 128;

inline uint32_t blocks(size_t n) { return (uint32_t)std::max<size_t>(1, (n + kBlock - 1) / kBlock); }

// attribute rows (input order, `stride_f` floats apart) -> float4 rows in the index's order
__global__ void __launch_bounds__(kBlock) gather3_kernel(const float4* __restrict__ pts, uint32_t n,
                                                         const float* __restrict__ in,
                                                         uint32_t stride_f,
                                                         float4* __restrict__ out)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= n)
        return;
    const float* r = in + (size_t)__float_as_uint(pts[t].w) * stride_f;
    out[t]         = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), 0.f);
}

// float4 rows in the index's order -> packed xyz rows in input order
__global__ void __launch_bounds__(kBlock) scatter3_kernel(const float4* __restrict__ pts, uint32_t n,
                                                          const float4* __restrict__ in,
                                                          float* __restrict__ out)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= n)
        return;
    float* r       = out + 3 * (size_t)__float_as_uint(pts[t].w);
    float4 const v = in[t];
    r[0] = v.x, r[1] = v.y, r[2] = v.z;
}

__global__ void __launch_bounds__(kBlock) take_rows_kernel(const float* __restrict__ xyz,
                                                           uint32_t stride_f,
                                                           const uint32_t* __restrict__ idx,
                                                           uint32_t m, float* __restrict__ out)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= m)
        return;
    const float* r = xyz + (size_t)idx[t] * stride_f;
    out[3 * (size_t)t] = r[0], out[3 * (size_t)t + 1] = r[1], out[3 * (size_t)t + 2] = r[2];
}

__global__ void __launch_bounds__(kBlock) bilateral_points_kernel(
    GridView g, uint32_t n, const float4* __restrict__ nrm_sorted, float sigmaf, float sigmag,
    float* __restrict__ out_xyz)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= n)
        return;
    float4 const s = __ldg(g.pts + t);
    float o[3];
    bilateral_point(g, nrm_sorted, s.x, s.y, s.z, sigmaf, sigmag, o);
    float* r = out_xyz + 3 * (size_t)__float_as_uint(s.w);
    r[0] = o[0], r[1] = o[1], r[2] = o[2];
}

// iterates in the index's order: in / out are both sorted-order float4 rows
__global__ void __launch_bounds__(kBlock) bilateral_normals_kernel(
    GridView g, uint32_t n, const float4* __restrict__ nrm_in, float sigmaf, float sigmag,
    float4* __restrict__ nrm_out)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= n)
        return;
    float4 const s  = __ldg(g.pts + t);
    float4 const ns = nrm_in[t];
    float o[3];
    bilateral_normal(g, nrm_in, s.x, s.y, s.z, ns.x, ns.y, ns.z, sigmaf, sigmag, o);
    nrm_out[t] = make_float4(o[0], o[1], o[2], 0.f);
}

__global__ void __launch_bounds__(kBlock) wlop_density_kernel(GridView g, uint32_t n, WlopParams w,
                                                              float* __restrict__ out_sorted)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= n)
        return;
    float4 const s = __ldg(g.pts + t);
    out_sorted[t]  = wlop_density(g, w, s.x, s.y, s.z);
}

__global__ void __launch_bounds__(kBlock) wlop_step_kernel(GridView gp,
                                                           const float* __restrict__ vj_sorted,
                                                           GridView gq,
                                                           const float* __restrict__ wi_sorted,
                                                           uint32_t nq, WlopParams w,
                                                           float* __restrict__ out_xyz)
{
    uint32_t const t = blockIdx.x * kBlock + threadIdx.x;
    if (t >= nq)
        return;
    float4 const q = __ldg(gq.pts + t);
    float o[3];
    wlop_step(gp, vj_sorted, gq, wi_sorted, w, q.x, q.y, q.z, o);
    float* r = out_xyz + 3 * (size_t)__float_as_uint(q.w);
    r[0] = o[0], r[1] = o[1], r[2] = o[2];
}

int pick_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        fail(PCPX_ERR_NO_DEVICE, "no CUDA device: libpcpx has no CPU path");
    if (device < 0)
        PCPX_CUDA(cudaGetDevice(&device));
    if (device >= ndev)
        fail(PCPX_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, ndev);
    return device;
}

size_t checked_stride(size_t stride_bytes, const char* what)
{
    if (stride_bytes == 0)
        stride_bytes = 12;
    if (stride_bytes < 12 || stride_bytes % 4)
        fail(PCPX_ERR_INVALID_ARG, "%s must be a multiple of 4 and >= 12", what);
    return stride_bytes;
}

std::unique_ptr<pcpx_index> index_on(const float* d_xyz, size_t n, int device)
{
    pcpx_index_params prm{};
    prm.device = device;
    return std::unique_ptr<pcpx_index>(build_index(d_xyz, n, 12, &prm));
}

// A stream of this call's own plus a start / stop event pair around everything it enqueues.
// The per-iteration indices run on their own streams; the call serialises them by
// synchronising each before the next is built, so the event pair brackets all device work.
struct CallClock
{
    cudaStream_t s = nullptr;
    Event t0, t1;
    CallClock()
    {
        PCPX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        t0.record(s);
    }
    ~CallClock()
    {
        if (s)
            cudaStreamDestroy(s);
    }
    void finish(float* out_ms)
    {
        t1.record(s);
        PCPX_CUDA(cudaStreamSynchronize(s));
        if (out_ms)
            *out_ms = elapsed_ms(t0, t1);
    }
};

// packed device copy of n xyz rows, whatever the source
void pack_rows(const float* p, size_t n, size_t stride_bytes, DevBuf<float>& dst, cudaStream_t s)
{
    dst.alloc(3 * n);
    cudaMemcpyKind const kind =
        is_device_pointer(p) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (stride_bytes == 12)
        PCPX_CUDA(cudaMemcpyAsync(dst.get(), p, 12 * n, kind, s));
    else
        PCPX_CUDA(cudaMemcpy2DAsync(dst.get(), 12, p, stride_bytes, 12, n, kind, s));
}

void copy_out(float* user, const float* d, size_t n_floats, cudaStream_t s)
{
    cudaMemcpyKind const kind =
        is_device_pointer(user) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    PCPX_CUDA(cudaMemcpyAsync(user, d, n_floats * 4, kind, s));
}

void check_bilateral_args(const float* xyz, size_t n, const float* normals, double sigmaf,
                          double sigmag, const float* out)
{
    if (n >= 0xFFFFFFFFull)
        fail(PCPX_ERR_UNSUPPORTED, "clouds of 2^32 - 1 points or more need 64-bit indices");
    if (n && (!xyz || !normals || !out))
        fail(PCPX_ERR_INVALID_ARG, "xyz / normals / out is NULL");
    // algorithm/bilateral_filter.hpp:332-335 asserts these
    if (!(sigmaf > 0.0) || !(sigmag > 0.0))
        fail(PCPX_ERR_INVALID_ARG, "sigmaf and sigmag must be positive");
}

} // namespace

extern "C" {

int pcpx_bilateral_filter_points(const float* xyz, size_t n, size_t stride_bytes,
                                 const float* normals, size_t normal_stride_bytes, double sigmaf,
                                 double sigmag, uint32_t iterations, int device, float* out_xyz,
                                 float* out_device_ms)
{
    return guarded([&] {
        check_bilateral_args(xyz, n, normals, sigmaf, sigmag, out_xyz);
        if (n == 0)
            return;
        stride_bytes        = checked_stride(stride_bytes, "stride_bytes");
        normal_stride_bytes = checked_stride(normal_stride_bytes, "normal_stride_bytes");
        device              = pick_device(device);
        ScopedDevice guard(device);
        CallClock clk;
        DevBuf<float> a, b;
        pack_rows(xyz, n, stride_bytes, a, clk.s);
        b.alloc(3 * n);
        InBuf nrm;
        nrm.stage(normals, n, normal_stride_bytes, 3, clk.s);
        DevBuf<float4> nrm_sorted(n);
        PCPX_CUDA(cudaStreamSynchronize(clk.s));
        float *cur = a.get(), *nxt = b.get();
        for (uint32_t it = 0; it < iterations; ++it)
        {
            auto ix = index_on(cur, n, device); // a fresh index per iteration, as :390-396
            gather3_kernel<<<blocks(n), kBlock, 0, ix->stream>>>(ix->grid.pts, (uint32_t)n, nrm.d,
                                                                 nrm.stride_f, nrm_sorted.get());
            bilateral_points_kernel<<<blocks(n), kBlock, 0, ix->stream>>>(
                ix->grid, (uint32_t)n, nrm_sorted.get(), (float)sigmaf, (float)sigmag, nxt);
            PCPX_CHECK_LAUNCH();
            PCPX_CUDA(cudaStreamSynchronize(ix->stream));
            std::swap(cur, nxt);
        }
        copy_out(out_xyz, cur, 3 * n, clk.s);
        clk.finish(out_device_ms);
    });
}

int pcpx_bilateral_filter_normals(const float* xyz, size_t n, size_t stride_bytes,
                                  const float* normals, size_t normal_stride_bytes, double sigmaf,
                                  double sigmag, uint32_t iterations, int device,
                                  float* out_normals, float* out_device_ms)
{
    return guarded([&] {
        check_bilateral_args(xyz, n, normals, sigmaf, sigmag, out_normals);
        if (n == 0)
            return;
        stride_bytes        = checked_stride(stride_bytes, "stride_bytes");
        normal_stride_bytes = checked_stride(normal_stride_bytes, "normal_stride_bytes");
        device              = pick_device(device);
        ScopedDevice guard(device);
        CallClock clk;
        pcpx_index_params prm{};
        prm.device = device;
        std::unique_ptr<pcpx_index> ix(build_index(xyz, n, stride_bytes, &prm)); // one index (:542)
        InBuf nrm;
        nrm.stage(normals, n, normal_stride_bytes, 3, ix->stream);
        DevBuf<float4> s0(n), s1(n);
        DevBuf<float> packed(3 * n);
        gather3_kernel<<<blocks(n), kBlock, 0, ix->stream>>>(ix->grid.pts, (uint32_t)n, nrm.d,
                                                             nrm.stride_f, s0.get());
        float4 *cur = s0.get(), *nxt = s1.get();
        for (uint32_t it = 0; it < iterations; ++it)
        {
            bilateral_normals_kernel<<<blocks(n), kBlock, 0, ix->stream>>>(
                ix->grid, (uint32_t)n, cur, (float)sigmaf, (float)sigmag, nxt);
            std::swap(cur, nxt);
        }
        scatter3_kernel<<<blocks(n), kBlock, 0, ix->stream>>>(ix->grid.pts, (uint32_t)n, cur,
                                                              packed.get());
        PCPX_CHECK_LAUNCH();
        copy_out(out_normals, packed.get(), 3 * n, ix->stream);
        PCPX_CUDA(cudaStreamSynchronize(ix->stream));
        clk.finish(out_device_ms);
    });
}

int pcpx_wlop(const float* xyz, size_t n, size_t stride_bytes, const uint32_t* initial_idx,
              size_t n_out, double mu, double h, uint32_t iterations, int uniform, uint32_t seed,
              int device, float* out_xyz, float* out_device_ms)
{
    return guarded([&] {
        if (n >= 0xFFFFFFFFull)
            fail(PCPX_ERR_UNSUPPORTED, "clouds of 2^32 - 1 points or more need 64-bit indices");
        // algorithm/wlop.hpp:305-307 asserts these
        if (n_out == 0 || n_out > n)
            fail(PCPX_ERR_INVALID_ARG, "n_out must satisfy 0 < n_out <= n");
        if (!(mu >= 0.0 && mu <= 0.5))
            fail(PCPX_ERR_INVALID_ARG, "mu must lie in [0, 0.5]");
        if (!(h > 0.0))
            fail(PCPX_ERR_INVALID_ARG, "h must be positive");
        if (!xyz || !out_xyz)
            fail(PCPX_ERR_INVALID_ARG, "xyz / out_xyz is NULL");
        stride_bytes = checked_stride(stride_bytes, "stride_bytes");
        device       = pick_device(device);
        ScopedDevice guard(device);
        CallClock clk;

        DevBuf<float> p_packed; // the input cloud P, packed, on the device
        pack_rows(xyz, n, stride_bytes, p_packed, clk.s);

        // start set X0 (:346-358)
        DevBuf<uint32_t> d_init(n_out);
        if (initial_idx)
        {
            cudaMemcpyKind const kind = is_device_pointer(initial_idx) ? cudaMemcpyDeviceToDevice
                                                                        : cudaMemcpyHostToDevice;
            if (kind == cudaMemcpyHostToDevice)
                for (size_t i = 0; i < n_out; ++i)
                    if (initial_idx[i] >= n)
                        fail(PCPX_ERR_INVALID_ARG, "initial_idx[%llu] = %u is not a point index",
                             (unsigned long long)i, initial_idx[i]);
            PCPX_CUDA(cudaMemcpyAsync(d_init.get(), initial_idx, n_out * 4, kind, clk.s));
            PCPX_CUDA(cudaStreamSynchronize(clk.s));
            if (kind == cudaMemcpyDeviceToDevice)
            {
                uint32_t const bad =
                    count_bad_indices(clk.s, d_init.get(), n_out, (uint32_t)n, 0u /* 0 is a point */);
                if (bad)
                    fail(PCPX_ERR_INVALID_ARG, "%u entries of initial_idx are not point indices", bad);
            }
        }
        else
        {
            std::vector<uint32_t> js(n);
            std::iota(js.begin(), js.end(), 0u);
            std::mt19937 gen(seed);
            std::shuffle(js.begin(), js.end(), gen);
            PCPX_CUDA(cudaMemcpyAsync(d_init.get(), js.data() + (n - n_out), n_out * 4,
                                      cudaMemcpyHostToDevice, clk.s));
            PCPX_CUDA(cudaStreamSynchronize(clk.s));
        }
        DevBuf<float> xa(3 * n_out), xb(3 * n_out);
        take_rows_kernel<<<blocks(n_out), kBlock, 0, clk.s>>>(p_packed.get(), 3u, d_init.get(),
                                                              (uint32_t)n_out, xa.get());
        PCPX_CHECK_LAUNCH();
        PCPX_CUDA(cudaStreamSynchronize(clk.s));

        WlopParams const w = wlop_params((float)h, (float)mu);
        auto ixp           = index_on(p_packed.get(), n, device); // p_kdtree (:365-369)
        DevBuf<float> vj, wi;
        if (uniform)
        {
            vj.alloc(n), wi.alloc(n_out);
            wlop_density_kernel<<<blocks(n), kBlock, 0, ixp->stream>>>(ixp->grid, (uint32_t)n, w,
                                                                       vj.get());
            PCPX_CHECK_LAUNCH();
        }
        PCPX_CUDA(cudaStreamSynchronize(ixp->stream));

        float *cur = xa.get(), *nxt = xb.get();
        for (uint32_t it = 0; it < iterations; ++it)
        {
            auto ixq = index_on(cur, n_out, device); // q_kdtree, rebuilt every iteration (:385)
            if (uniform)
                wlop_density_kernel<<<blocks(n_out), kBlock, 0, ixq->stream>>>(
                    ixq->grid, (uint32_t)n_out, w, wi.get());
            wlop_step_kernel<<<blocks(n_out), kBlock, 0, ixq->stream>>>(
                ixp->grid, uniform ? vj.get() : nullptr, ixq->grid, uniform ? wi.get() : nullptr,
                (uint32_t)n_out, w, nxt);
            PCPX_CHECK_LAUNCH();
            PCPX_CUDA(cudaStreamSynchronize(ixq->stream));
            std::swap(cur, nxt);
        }
        copy_out(out_xyz, cur, 3 * n_out, clk.s);
        clk.finish(out_device_ms);
    });
}

} // extern "C"
