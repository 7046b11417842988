// Query kernels: one thread per query, queries processed in Morton order so that the lanes of a
// warp walk the same few cells (their candidate loads coalesce into a handful of L1 lines).
// No tensor cores: nothing here is a dense contraction; the work is fp32 compares, 64-bit
// compare-selects and cached 16-byte loads.
#include <algorithm>

#include "big_k.cuh"
#include "normals_core.cuh"
#include "plan.hpp"
#include "query.hpp"
#include "radius_core.cuh"
#include "tree_core.cuh"

namespace pcpx {

Tuning& tuning()
{
    static Tuning t;
    return t;
}

namespace {

constexpr int kQBlock = 128;
// The kNN kernels are latency-bound (dependent L1 / local-memory loads in the span walk), so
// occupancy is bought with a register cap: 10 CTAs of 128 threads per SM (48 registers) for
// lists up to 16 entries measured best on B200 (4.3 ms vs 6.2 ms uncapped at k = 15, 10 M
// points); longer lists get proportionally more registers.
#ifndef PCPX_MIN_BLOCKS_SMALL_K
#define PCPX_MIN_BLOCKS_SMALL_K 10
#endif
__host__ __device__ constexpr int min_blocks_for(int K)
{
    // measured at 10 M points: k = 15 3.63 ms with 10 blocks (3.94 with 8, 3.75 with 12);
    // k = 8 (thick shell) 8.19 ms with 12 blocks against 8.68 with 10
    return K <= 8 ? 12 : (K <= 16 ? PCPX_MIN_BLOCKS_SMALL_K : (K <= 24 ? 9 : 7));
}

inline uint32_t grid_for(uint32_t n, int block) { return std::max(1u, (n + block - 1) / block); }

__device__ __forceinline__ bool fetch_query(const GridView& g, const QueryBatch& qb, uint32_t t,
                                            float& x, float& y, float& z, uint32_t& row)
{
    if (t >= qb.nq)
        return false;
    if (qb.q == nullptr)
    {
        float4 const c = __ldg(g.pts + t);
        x = c.x, y = c.y, z = c.z;
        row = __float_as_uint(c.w);
    }
    else
    {
        row            = qb.order ? qb.order[t] : t;
        const float* p = qb.q + (size_t)row * qb.stride_f;
        x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    }
    return true;
}

// ---- kNN-shaped kernels ----------------------------------------------------------------------
// Main pass: every query makes ONE attempt at the call's main level (chosen on the host from
// the per-level cell occupancy so that the level's 3x3x3 block holds a few times k points).
// All lanes of a warp therefore do the same amount of structure work, and the few queries whose
// neighbourhood reaches past their block (sparse regions, outliers) do not stall 31 other
// lanes: they are appended to a retry queue and answered by a second launch that walks to
// coarser levels.  Queries whose answer hinges on bit-equal distances take the exact 64-bit
// (distance, original index) search in a noinline function (register footprint stays off the
// fast path).
enum KnnMode
{
    MODE_KNN     = 0, // index rows (+ distances, counts)
    MODE_MEAN    = 1, // mean neighbour distance
    MODE_NORMALS = 2  // fused PCA normal (+ centroid)
};

struct KnnOutputs
{
    uint32_t* idx;
    float* d2;
    uint32_t* count;
    float* mean;
    float* normal;
    float* centroid;
    uint32_t* exact_counter; // queries answered by the exact tie path
    uint32_t* retry_items;   // queue of thread ids for the second launch
    uint32_t* retry_count;
};

__host__ __device__ constexpr int exact_k(int K) { return (K + 3) / 4 * 4; }

template <int K>
__device__ __noinline__ void knn_exact_row(const GridView& g, float x, float y, float z,
                                           uint32_t k, float eps, int start_level,
                                           uint32_t* idx_row, float* d2_row, uint32_t* out_count)
{
    TopK<exact_k(K)> top;
    knn_search<exact_k(K), TIE_ORIGINAL_INDEX>(g, x, y, z, k, eps, start_level, top, nullptr);
    uint32_t n = 0;
#pragma unroll
    for (int j = 0; j < exact_k(K); ++j)
        if ((uint32_t)j < k)
        {
            bool const valid = top.a[j] != kEmptyEntry;
            idx_row[j]       = valid ? (uint32_t)top.a[j] : PCPX_NO_NEIGHBOUR;
            if (d2_row)
                d2_row[j] = valid ? __uint_as_float((uint32_t)(top.a[j] >> 32)) : INFINITY;
            n += valid;
        }
    if (out_count)
        *out_count = n;
}

template <int K>
__device__ __noinline__ void normal_exact(const GridView& g, float x, float y, float z,
                                          uint32_t k, float eps, int start_level,
                                          float* out_normal_row, float* out_centroid_row)
{
    TopK<exact_k(K)> ids;
    int const level =
        knn_search<exact_k(K), TIE_ORIGINAL_INDEX>(g, x, y, z, k, eps, start_level, ids, nullptr);
    float n3[3], c3[3];
    normal_from_ids(g, query_cell(g, x, y, z), level, ids, k, n3, c3, nullptr);
    out_normal_row[0] = n3[0], out_normal_row[1] = n3[1], out_normal_row[2] = n3[2];
    if (out_centroid_row)
        out_centroid_row[0] = c3[0], out_centroid_row[1] = c3[1], out_centroid_row[2] = c3[2];
}

// second pass + output of one query whose first pass is final
template <int K, int MODE, class Region, class SL>
__device__ __forceinline__ void knn_finish(const GridView& g, const Region& region, int level,
                                           const SL& sl, float x, float y, float z,
                                           const TopD<K>& top, uint32_t k, float eps,
                                           uint32_t row, const KnnOutputs& out)
{
    if (MODE == MODE_MEAN)
    {
        out.mean[row] = mean_distance_d(top, k);
    }
    else if (MODE == MODE_KNN)
    {
        uint32_t* idx_row = out.idx + (size_t)row * k;
        float* d2_row     = out.d2 ? out.d2 + (size_t)row * k : nullptr;
        uint32_t* cnt     = out.count ? out.count + row : nullptr;
        if (!knn_two_pass_emit<K>(g, region, sl, x, y, z, top, k, eps, idx_row, d2_row, cnt))
        {
            knn_exact_row<K>(g, x, y, z, k, eps, level, idx_row, d2_row, cnt);
            if (out.exact_counter)
                atomicAdd(out.exact_counter, 1u);
        }
    }
    else
    {
        float* nrow = out.normal + 3 * (size_t)row;
        float* crow = out.centroid ? out.centroid + 3 * (size_t)row : nullptr;
        float n3[3], c3[3];
        if (normal_two_pass<K>(g, region, sl, x, y, z, top, k, eps, n3, c3, nullptr))
        {
            nrow[0] = n3[0], nrow[1] = n3[1], nrow[2] = n3[2];
            if (crow)
                crow[0] = c3[0], crow[1] = c3[1], crow[2] = c3[2];
        }
        else
        {
            // which of the equidistant points is a neighbour is decided by the original index
            normal_exact<K>(g, x, y, z, k, eps, level, nrow, crow);
            if (out.exact_counter)
                atomicAdd(out.exact_counter, 1u);
        }
    }
}

template <int K, int MODE, int RINGS>
__global__ void __launch_bounds__(kQBlock, min_blocks_for(K)) knn_main_kernel(
    GridView g, QueryBatch qb, uint32_t k, float eps, int level, KnnOutputs out)
{
    uint32_t const t = blockIdx.x * kQBlock + threadIdx.x;
    float x, y, z;
    uint32_t row;
    if (!fetch_query(g, qb, t, x, y, z, row))
        return;
    TopD<K> top;
    BlockGeom b;
    CellList cl;
    ShortListFor<K> sl;
    QueryCell const qc = query_cell(g, x, y, z);
    if (knn_attempt_dist<K, RINGS>(g, qc, level, x, y, z, k, eps, top, b, cl, sl, nullptr))
        knn_finish<K, MODE>(g, BlockRegion<RINGS>{b, level}, level, sl, x, y, z, top, k, eps, row,
                            out);
    else
        out.retry_items[atomicAdd(out.retry_count, 1u)] = t;
}

// The queries the main pass could not finish.  A near miss (the k-ball pokes just past the
// block) is usually final one level coarser, which is cheap; what is still open after that —
// sparse regions, outliers — descends the octree (tree_core.cuh) instead of trying ever coarser
// blocks whose cells hold thousands of points.
template <int K, int MODE, int RINGS>
__global__ void __launch_bounds__(kQBlock) knn_retry_kernel(GridView g, QueryBatch qb, uint32_t k,
                                                            float eps, int level, KnnOutputs out)
{
    uint32_t const n_retry = *out.retry_count;
    for (uint32_t i = blockIdx.x * kQBlock + threadIdx.x; i < n_retry; i += gridDim.x * kQBlock)
    {
        float x, y, z;
        uint32_t row;
        fetch_query(g, qb, out.retry_items[i], x, y, z, row);
        TopD<K> top;
        ShortListFor<K> sl;
        bool done = false;
        if (level > 0)
        {
            BlockGeom b;
            CellList cl;
            int const coarser = level - 1;
            if (knn_attempt_dist<K, RINGS>(g, query_cell(g, x, y, z), coarser, x, y, z, k, eps,
                                           top, b, cl, sl, nullptr))
            {
                knn_finish<K, MODE>(g, BlockRegion<RINGS>{b, coarser}, coarser, sl, x, y, z, top,
                                    k, eps, row, out);
                done = true;
            }
        }
        if (!done)
        {
            knn_tree_dist<K>(g, x, y, z, eps, top, sl, nullptr);
            knn_finish<K, MODE>(g, TreeRegion{}, level, sl, x, y, z, top, k, eps, row, out);
        }
    }
}

// ---- k beyond the register list (big_k.cuh) --------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(kQBlock) knn_big_kernel(
    GridView g, QueryBatch qb, uint32_t k, float eps, int level, uint32_t* __restrict__ out_idx,
    float* __restrict__ out_d2, uint32_t* __restrict__ out_count)
{
    float x, y, z;
    uint32_t row;
    if (!fetch_query(g, qb, blockIdx.x * kQBlock + threadIdx.x, x, y, z, row))
        return;
    uint64_t keys[CAP];
    BigHeap heap{keys, 0u, k};
    knn_search_big(g, x, y, z, eps, level, heap);
    size_t const base = (size_t)row * k;
    for (uint32_t j = 0; j < k; ++j)
    {
        bool const valid = j < heap.n;
        if (out_idx)
            out_idx[base + j] = valid ? (uint32_t)keys[j] : PCPX_NO_NEIGHBOUR;
        if (out_d2)
            out_d2[base + j] = valid ? __uint_as_float((uint32_t)(keys[j] >> 32)) : INFINITY;
    }
    if (out_count)
        out_count[row] = heap.n;
}

__global__ void __launch_bounds__(256) inverse_order_kernel(GridView g, uint32_t n_total,
                                                            uint32_t* __restrict__ inv)
{
    uint32_t const i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_total)
        inv[__float_as_uint(g.pts[i].w)] = i;
}

// PCA normal / mean distance from finished index rows (any k)
__global__ void __launch_bounds__(kQBlock) rows_to_normals_kernel(
    GridView g, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ inv, uint32_t nq,
    uint32_t k, float* __restrict__ out_centroid, float* __restrict__ out_normal)
{
    uint32_t const q = blockIdx.x * kQBlock + threadIdx.x;
    if (q >= nq)
        return;
    const uint32_t* row = idx + (size_t)q * k;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t n = 0;
    for (uint32_t j = 0; j < k && row[j] != PCPX_NO_NEIGHBOUR; ++j, ++n)
    {
        float4 const c = __ldg(g.pts + inv[row[j]]);
        sx += c.x, sy += c.y, sz += c.z;
    }
    float const inv_n = 1.f / (float)n;
    float const mx = sx * inv_n, my = sy * inv_n, mz = sz * inv_n;
    Sym3 m{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t j = 0; j < n; ++j)
    {
        float4 const c = __ldg(g.pts + inv[row[j]]);
        float const x = c.x - mx, y = c.y - my, z = c.z - mz;
        m.xx += x * x, m.xy += x * y, m.xz += x * z;
        m.yy += y * y, m.yz += y * z, m.zz += z * z;
    }
    float nx, ny, nz;
    smallest_eigenvector(m, nx, ny, nz, nullptr);
    out_normal[3 * (size_t)q] = nx, out_normal[3 * (size_t)q + 1] = ny,
                           out_normal[3 * (size_t)q + 2] = nz;
    if (out_centroid)
        out_centroid[3 * (size_t)q] = mx, out_centroid[3 * (size_t)q + 1] = my,
                                 out_centroid[3 * (size_t)q + 2] = mz;
}

__global__ void __launch_bounds__(kQBlock) rows_to_mean_kernel(
    const float* __restrict__ d2, uint32_t nq, uint32_t k, float* __restrict__ out_mean)
{
    uint32_t const q = blockIdx.x * kQBlock + threadIdx.x;
    if (q >= nq)
        return;
    const float* row = d2 + (size_t)q * k;
    float sum        = 0.f;
    uint32_t n       = 0;
    for (uint32_t j = 0; j < k && row[j] < INFINITY; ++j, ++n)
        sum = __fadd_rn(sum, __fsqrt_rn(row[j]));
    out_mean[q] = __fdiv_rn(sum, (float)n);
}

// ---- instrumentation: what the search does per query ---------------------------------------
template <int K, int RINGS>
__global__ void __launch_bounds__(kQBlock) knn_stats_kernel(
    GridView g, QueryBatch qb, uint32_t k, float eps, int level,
    unsigned long long* __restrict__ stats4)
{
    float x, y, z;
    uint32_t row;
    SearchStats st;
    bool const live = fetch_query(g, qb, blockIdx.x * kQBlock + threadIdx.x, x, y, z, row);
    if (live)
    {
        TopD<K> top;
        BlockGeom b;
        CellList cl;
        ShortListFor<K> sl;
        QueryCell const qc = query_cell(g, x, y, z);
        if (!knn_attempt_dist<K, RINGS>(g, qc, level, x, y, z, k, eps, top, b, cl, sl, &st) &&
            !(level > 0 &&
              knn_attempt_dist<K, RINGS>(g, qc, level - 1, x, y, z, k, eps, top, b, cl, sl, &st)))
            knn_tree_dist<K>(g, x, y, z, eps, top, sl, &st);
    }
    uint32_t const warp_max = __reduce_max_sync(0xFFFFFFFFu, st.candidates);
    unsigned long long v[4] = {st.candidates, st.lookups, st.attempts,
                               (threadIdx.x & 31) == 0 ? warp_max * 32ull : 0ull};
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], o);
        if ((threadIdx.x & 31) == 0 && v[i])
            atomicAdd(&stats4[i], v[i]);
    }
}

// ---- PCA of caller-supplied neighbourhoods ---------------------------------------------------
// common/normals/normal_estimation.hpp:41-77 per neighbourhood: fp32 mean, centred scatter,
// eigenvector of the smallest eigenvalue.  One thread per neighbourhood.
__global__ void __launch_bounds__(kQBlock) neighbourhood_normals_kernel(
    const float* __restrict__ nbr, const uint64_t* __restrict__ offsets, uint32_t n,
    float* __restrict__ out)
{
    uint32_t const i = blockIdx.x * kQBlock + threadIdx.x;
    if (i >= n)
        return;
    uint64_t const b = offsets[i], e = offsets[i + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (uint64_t j = b; j < e; ++j)
        sx += nbr[3 * j], sy += nbr[3 * j + 1], sz += nbr[3 * j + 2];
    float const inv = 1.f / (float)(e - b);
    float const mx = sx * inv, my = sy * inv, mz = sz * inv;
    Sym3 m{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint64_t j = b; j < e; ++j)
    {
        float const x = nbr[3 * j] - mx, y = nbr[3 * j + 1] - my, z = nbr[3 * j + 2] - mz;
        m.xx += x * x, m.xy += x * y, m.xz += x * z;
        m.yy += y * y, m.yz += y * z, m.zz += z * z;
    }
    float nx, ny, nz;
    smallest_eigenvector(m, nx, ny, nz, nullptr);
    out[3 * (size_t)i] = nx, out[3 * (size_t)i + 1] = ny, out[3 * (size_t)i + 2] = nz;
}

// ---- radius --------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQBlock) radius_count_kernel(
    GridView g, QueryBatch qb, const float* __restrict__ radii, float r,
    uint32_t* __restrict__ out_count)
{
    float x, y, z;
    uint32_t row;
    if (!fetch_query(g, qb, blockIdx.x * kQBlock + threadIdx.x, x, y, z, row))
        return;
    float const rq = radii ? radii[row] : r;
    uint32_t cnt   = 0;
    radius_visit(g, x, y, z, rq, [&](float4 const&, uint32_t) {
        ++cnt;
        return false;
    });
    out_count[row] = cnt;
}

__global__ void __launch_bounds__(kQBlock) radius_fill_kernel(
    GridView g, QueryBatch qb, const float* __restrict__ radii, float r,
    const uint64_t* __restrict__ offsets, uint32_t* __restrict__ out_idx)
{
    float x, y, z;
    uint32_t row;
    if (!fetch_query(g, qb, blockIdx.x * kQBlock + threadIdx.x, x, y, z, row))
        return;
    float const rq = radii ? radii[row] : r;
    uint64_t w     = offsets[row];
    radius_visit(g, x, y, z, rq, [&](float4 const& c, uint32_t) {
        out_idx[w++] = __float_as_uint(c.w);
        return false;
    });
}

// Density filter, first half: radius count with early exit at the threshold -> keep flag,
// written at the point's ORIGINAL position (examples/filter_point_cloud_noise_by_density.cpp:81-91).
__global__ void __launch_bounds__(kQBlock) density_keep_kernel(
    GridView g, uint32_t n_total, float r, uint32_t threshold, uint8_t* __restrict__ keep)
{
    uint32_t const t = blockIdx.x * kQBlock + threadIdx.x;
    if (t >= n_total)
        return;
    float4 const q = __ldg(g.pts + t);
    uint32_t cnt   = 0;
    if (threshold > 0)
        radius_visit(g, q.x, q.y, q.z, r, [&](float4 const&, uint32_t) {
            return ++cnt >= threshold;
        });
    keep[__float_as_uint(q.w)] = cnt >= threshold; // !(density < threshold)
}

// ---- scans / compaction / reduction --------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile  = kScanBlock * kScanItems;

template <typename T>
__global__ void __launch_bounds__(kScanBlock) tile_sums_kernel(
    const T* __restrict__ in, uint32_t n, uint64_t* __restrict__ tile_sums)
{
    uint32_t const base = blockIdx.x * kScanTile;
    uint32_t s          = 0;
    for (int r = 0; r < kScanItems; ++r)
    {
        uint32_t const i = base + r * kScanBlock + threadIdx.x;
        if (i < n)
            s += in[i];
    }
    __shared__ uint32_t sh[kScanBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint64_t tot = 0;
        for (int w = 0; w < kScanBlock / 32; ++w)
            tot += sh[w];
        tile_sums[blockIdx.x] = tot;
    }
}

// single CTA: exclusive scan of the tile sums in place; total appended at [n_tiles]
__global__ void __launch_bounds__(kScanBlock) scan_tile_sums_kernel(uint64_t* tile_sums,
                                                                    uint32_t n_tiles)
{
    __shared__ uint64_t wsum[kScanBlock / 32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += kScanBlock)
    {
        uint32_t const t = t0 + threadIdx.x;
        uint64_t const v = t < n_tiles ? tile_sums[t] : 0ull;
        uint64_t incl    = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint64_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o)
                incl += up;
        }
        if (lane == 31)
            wsum[warp] = incl;
        __syncthreads();
        uint64_t wbase = 0;
        for (int w = 0; w < warp; ++w)
            wbase += wsum[w];
        uint64_t const c = carry;
        if (t < n_tiles)
            tile_sums[t] = c + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1)
            carry = c + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        tile_sums[n_tiles] = carry;
}

template <typename T>
__global__ void __launch_bounds__(kScanBlock) tile_scan_kernel(
    const T* __restrict__ in, uint32_t n, const uint64_t* __restrict__ tile_sums,
    uint32_t n_tiles, uint64_t* __restrict__ out)
{
    // thread-blocked layout: thread owns kScanItems consecutive values
    uint32_t const base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r)
    {
        v[r] = base + r < n ? (uint32_t)in[base + r] : 0u;
        s += v[r];
    }
    __shared__ uint32_t wsum[kScanBlock / 32];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o)
            incl += up;
    }
    if (lane == 31)
        wsum[warp] = incl;
    __syncthreads();
    uint64_t run = tile_sums[blockIdx.x] + (incl - s);
    for (int w = 0; w < warp; ++w)
        run += wsum[w];
#pragma unroll
    for (int r = 0; r < kScanItems; ++r)
    {
        if (base + r < n)
            out[base + r] = run;
        run += v[r];
    }
    if (blockIdx.x == n_tiles - 1 && threadIdx.x == 0)
        out[n] = tile_sums[n_tiles];
}

// kept points out in ORIGINAL relative order (std::remove_if is stable)
__global__ void __launch_bounds__(kQBlock) compact_points_kernel(
    GridView g, uint32_t n_total, const uint8_t* __restrict__ keep,
    const uint64_t* __restrict__ scan, float* __restrict__ out_xyz)
{
    uint32_t const t = blockIdx.x * kQBlock + threadIdx.x;
    if (t >= n_total)
        return;
    float4 const p   = __ldg(g.pts + t);
    uint32_t const o = __float_as_uint(p.w);
    if (keep[o])
    {
        uint64_t const d   = scan[o];
        out_xyz[3 * d]     = p.x;
        out_xyz[3 * d + 1] = p.y;
        out_xyz[3 * d + 2] = p.z;
    }
}

// fixed-shape fp64 tree: per-CTA partial sums over a fixed slice, then one CTA adds the partials
// in index order -> bit-reproducible for a given n.  NaN means (0 neighbours) are skipped and
// counted out.
constexpr int kRedBlocks = 1024;
__global__ void __launch_bounds__(256) mean_partial_kernel(
    const float* __restrict__ v, uint32_t n, double* __restrict__ partial,
    uint32_t* __restrict__ partial_valid)
{
    uint64_t const per = ((uint64_t)n + kRedBlocks - 1) / kRedBlocks;
    uint64_t const b = per * blockIdx.x, e = min((uint64_t)n, b + per);
    double s   = 0.0;
    uint32_t c = 0;
    for (uint64_t i = b + threadIdx.x; i < e; i += 256)
    {
        float const f = v[i];
        if (f == f)
            s += (double)f, ++c;
    }
    __shared__ double sh[256];
    __shared__ uint32_t shc[256];
    sh[threadIdx.x] = s, shc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o)
            sh[threadIdx.x] += sh[threadIdx.x + o], shc[threadIdx.x] += shc[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        partial[blockIdx.x] = sh[0], partial_valid[blockIdx.x] = shc[0];
}
__global__ void mean_final_kernel(const double* partial, const uint32_t* partial_valid,
                                  double* out_sum, uint32_t* out_valid)
{
    double s   = 0.0;
    uint32_t c = 0;
    for (int i = 0; i < kRedBlocks; ++i)
        s += partial[i], c += partial_valid[i];
    *out_sum = s, *out_valid = c;
}

} // namespace

// register-list sizes that are compiled; a query with k neighbours runs with the smallest K >= k
static int list_size_for(uint32_t k)
{
    static const int sizes[] = {4, 8, 10, 12, 15, 16, 20, 24, 28, 30, 32};
    for (int s : sizes)
        if ((uint32_t)s >= k)
            return s;
    return 0;
}

#define PCPX_DISPATCH_K(KR, CALL)                                                              \
    switch (KR)                                                                                \
    {                                                                                          \
    case 4: { constexpr int KK = 4; CALL; } break;                                             \
    case 8: { constexpr int KK = 8; CALL; } break;                                             \
    case 10: { constexpr int KK = 10; CALL; } break;                                           \
    case 12: { constexpr int KK = 12; CALL; } break;                                           \
    case 15: { constexpr int KK = 15; CALL; } break;                                           \
    case 16: { constexpr int KK = 16; CALL; } break;                                           \
    case 20: { constexpr int KK = 20; CALL; } break;                                           \
    case 24: { constexpr int KK = 24; CALL; } break;                                           \
    case 28: { constexpr int KK = 28; CALL; } break;                                           \
    case 30: { constexpr int KK = 30; CALL; } break;                                           \
    case 32: { constexpr int KK = 32; CALL; } break;                                           \
    default: fail(PCPX_ERR_UNSUPPORTED, "k = %u is not supported (k <= %u)", k, kMaxK);        \
    }

// What a kNN-shaped call tries first (plan.hpp).
SearchPlan plan_for(const pcpx_index& ix, uint32_t k)
{
    PlanChoice const c = choose_plan(ix.n_indexed, ix.cells_per_level, ix.grid.lfine, k,
                                     (double)tuning().success_margin);
    return SearchPlan{c.level, c.rings};
}

// k > kMaxK: heap kernel -> index / distance rows -> (normals | means) from the rows
static void launch_knn_big(const pcpx_index& ix, const QueryBatch& qb, uint32_t k, float eps,
                           uint32_t* idx, float* d2, uint32_t* count)
{
    if (k > kBigKMax)
        fail(PCPX_ERR_UNSUPPORTED, "k = %u is not supported (k <= %u)", k, kBigKMax);
    int const level = plan_for(ix, k).level;
    dim3 const grid(grid_for(qb.nq, kQBlock));
    if (k <= 64)
        knn_big_kernel<64><<<grid, kQBlock, 0, ix.stream>>>(ix.grid, qb, k, eps, level, idx, d2,
                                                            count);
    else if (k <= 128)
        knn_big_kernel<128><<<grid, kQBlock, 0, ix.stream>>>(ix.grid, qb, k, eps, level, idx, d2,
                                                             count);
    else
        knn_big_kernel<256><<<grid, kQBlock, 0, ix.stream>>>(ix.grid, qb, k, eps, level, idx, d2,
                                                             count);
    PCPX_CHECK_LAUNCH();
}

template <int MODE>
static void launch_knn_shaped(const pcpx_index& ix, const QueryBatch& qb, uint32_t k, float eps,
                              KnnOutputs out, uint32_t* exact_counter, uint32_t* launches)
{
    if (k == 0)
        fail(PCPX_ERR_UNSUPPORTED, "k = 0 has no neighbourhood");
    if (k > kMaxK)
    {
        // rows are written per QUERY ROW (original order), so the row-wise epilogues index by row
        if (MODE == MODE_KNN)
            launch_knn_big(ix, qb, k, eps, out.idx, out.d2, out.count);
        else if (MODE == MODE_MEAN)
        {
            DevBuf<float> d2((size_t)qb.nq * k);
            launch_knn_big(ix, qb, k, eps, nullptr, d2.get(), nullptr);
            rows_to_mean_kernel<<<grid_for(qb.nq, kQBlock), kQBlock, 0, ix.stream>>>(
                d2.get(), qb.nq, k, out.mean);
            PCPX_CHECK_LAUNCH();
            PCPX_CUDA(cudaStreamSynchronize(ix.stream));
        }
        else
        {
            DevBuf<uint32_t> idx((size_t)qb.nq * k), inv(std::max<uint64_t>(ix.n_input, 1));
            launch_knn_big(ix, qb, k, eps, idx.get(), nullptr, nullptr);
            inverse_order_kernel<<<grid_for((uint32_t)ix.n_input, 256), 256, 0, ix.stream>>>(
                ix.grid, (uint32_t)ix.n_input, inv.get());
            PCPX_CHECK_LAUNCH();
            rows_to_normals_kernel<<<grid_for(qb.nq, kQBlock), kQBlock, 0, ix.stream>>>(
                ix.grid, idx.get(), inv.get(), qb.nq, k, out.centroid, out.normal);
            PCPX_CHECK_LAUNCH();
            PCPX_CUDA(cudaStreamSynchronize(ix.stream));
        }
        return;
    }
    uint32_t const kr     = (uint32_t)list_size_for(k);
    SearchPlan const plan = plan_for(ix, k);
    DevBuf<uint32_t> retry_items(qb.nq), retry_count(1);
    PCPX_CUDA(cudaMemsetAsync(retry_count.get(), 0, 4, ix.stream));
    out.exact_counter = exact_counter;
    out.retry_items   = retry_items.get();
    out.retry_count   = retry_count.get();
    dim3 const grid(grid_for(qb.nq, kQBlock));
    dim3 const retry_grid(std::min<uint32_t>(grid.x, 148u * 16u));
    if (plan.rings >= 2)
    {
        PCPX_DISPATCH_K(kr, (knn_main_kernel<KK, MODE, 2><<<grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, out)));
        PCPX_CHECK_LAUNCH();
    }
    else
    {
        PCPX_DISPATCH_K(kr, (knn_main_kernel<KK, MODE, 1><<<grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, out)));
        PCPX_CHECK_LAUNCH();
    }
    if (plan.rings >= 2)
    {
        PCPX_DISPATCH_K(kr, (knn_retry_kernel<KK, MODE, 2><<<retry_grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, out)));
    }
    else
    {
        PCPX_DISPATCH_K(kr, (knn_retry_kernel<KK, MODE, 1><<<retry_grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, out)));
    }
    PCPX_CHECK_LAUNCH();
    if (launches)
        *launches += 2;
    // the queue is read by the retry kernel: keep it until the stream is idle
    PCPX_CUDA(cudaStreamSynchronize(ix.stream));
}

void launch_knn(const pcpx_index& ix, const QueryBatch& qb, uint32_t k, float eps, uint32_t* idx,
                float* d2, uint32_t* count, uint32_t* exact_counter)
{
    if (qb.nq == 0 || k == 0)
        return;
    KnnOutputs out{};
    out.idx = idx, out.d2 = d2, out.count = count;
    launch_knn_shaped<MODE_KNN>(ix, qb, k, eps, out, exact_counter, nullptr);
}

void launch_mean_distance(const pcpx_index& ix, const QueryBatch& qb, uint32_t k, float eps,
                          float* means)
{
    if (qb.nq == 0)
        return;
    KnnOutputs out{};
    out.mean = means;
    launch_knn_shaped<MODE_MEAN>(ix, qb, k, eps, out, nullptr, nullptr);
}

void launch_normals(const pcpx_index& ix, const QueryBatch& qb, uint32_t k, float eps,
                    float* centroids, float* normals, uint32_t* exact_counter)
{
    if (qb.nq == 0)
        return;
    KnnOutputs out{};
    out.normal = normals, out.centroid = centroids;
    launch_knn_shaped<MODE_NORMALS>(ix, qb, k, eps, out, exact_counter, nullptr);
}

void launch_normals_from_neighbourhoods(cudaStream_t stream, const float* nbr_xyz,
                                        const uint64_t* offsets, uint32_t n, float* normals)
{
    if (n == 0)
        return;
    neighbourhood_normals_kernel<<<grid_for(n, kQBlock), kQBlock, 0, stream>>>(nbr_xyz, offsets, n,
                                                                                normals);
    PCPX_CHECK_LAUNCH();
}

void launch_knn_stats(const pcpx_index& ix, uint32_t k, float eps, unsigned long long* stats4)
{
    if (k == 0 || k > kMaxK) // instrumentation of the register-list path only
        fail(PCPX_ERR_UNSUPPORTED, "k = %u is not supported (1 <= k <= %u)", k, kMaxK);
    QueryBatch qb{nullptr, 3u, nullptr, (uint32_t)ix.n_input};
    if (qb.nq == 0)
        return;
    uint32_t const kr     = (uint32_t)list_size_for(k);
    SearchPlan const plan = plan_for(ix, k);
    dim3 const grid(grid_for(qb.nq, kQBlock));
    if (plan.rings >= 2)
    {
        PCPX_DISPATCH_K(kr, (knn_stats_kernel<KK, 2><<<grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, stats4)));
    }
    else
    {
        PCPX_DISPATCH_K(kr, (knn_stats_kernel<KK, 1><<<grid, kQBlock, 0, ix.stream>>>(
                                ix.grid, qb, k, eps, plan.level, stats4)));
    }
    PCPX_CHECK_LAUNCH();
}

void launch_radius_count(const pcpx_index& ix, const QueryBatch& qb, const float* radii, float r,
                         uint32_t* count)
{
    if (qb.nq == 0)
        return;
    radius_count_kernel<<<grid_for(qb.nq, kQBlock), kQBlock, 0, ix.stream>>>(ix.grid, qb, radii, r,
                                                                              count);
    PCPX_CHECK_LAUNCH();
}

void launch_radius_fill(const pcpx_index& ix, const QueryBatch& qb, const float* radii, float r,
                        const uint64_t* offsets, uint32_t* idx)
{
    if (qb.nq == 0)
        return;
    radius_fill_kernel<<<grid_for(qb.nq, kQBlock), kQBlock, 0, ix.stream>>>(ix.grid, qb, radii, r,
                                                                             offsets, idx);
    PCPX_CHECK_LAUNCH();
}

void launch_density_keep(const pcpx_index& ix, float r, uint32_t threshold, uint8_t* keep)
{
    uint32_t const n = (uint32_t)ix.n_input;
    if (n == 0)
        return;
    density_keep_kernel<<<grid_for(n, kQBlock), kQBlock, 0, ix.stream>>>(ix.grid, n, r, threshold,
                                                                          keep);
    PCPX_CHECK_LAUNCH();
}

template <typename T>
static void exclusive_scan(const pcpx_index& ix, const T* in, uint32_t n, uint64_t* out)
{
    if (n == 0)
    {
        PCPX_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), ix.stream));
        return;
    }
    uint32_t const n_tiles = (n + kScanTile - 1) / kScanTile;
    DevBuf<uint64_t> sums(n_tiles + 1);
    tile_sums_kernel<T><<<n_tiles, kScanBlock, 0, ix.stream>>>(in, n, sums.get());
    PCPX_CHECK_LAUNCH();
    scan_tile_sums_kernel<<<1, kScanBlock, 0, ix.stream>>>(sums.get(), n_tiles);
    PCPX_CHECK_LAUNCH();
    tile_scan_kernel<T><<<n_tiles, kScanBlock, 0, ix.stream>>>(in, n, sums.get(), n_tiles, out);
    PCPX_CHECK_LAUNCH();
    PCPX_CUDA(cudaStreamSynchronize(ix.stream)); // `sums` is freed on return
}

void launch_exclusive_scan_u32(const pcpx_index& ix, const uint32_t* in, uint32_t n, uint64_t* out)
{
    exclusive_scan<uint32_t>(ix, in, n, out);
}
void launch_exclusive_scan_u8(const pcpx_index& ix, const uint8_t* in, uint32_t n, uint64_t* out)
{
    exclusive_scan<uint8_t>(ix, in, n, out);
}

void launch_compact_points(const pcpx_index& ix, const uint8_t* keep, const uint64_t* scan,
                           float* out_xyz)
{
    uint32_t const n = (uint32_t)ix.n_input;
    if (n == 0)
        return;
    compact_points_kernel<<<grid_for(n, kQBlock), kQBlock, 0, ix.stream>>>(ix.grid, n, keep, scan,
                                                                            out_xyz);
    PCPX_CHECK_LAUNCH();
}

void launch_mean_reduce(const pcpx_index& ix, const float* v, uint32_t n, double* out_sum,
                        uint32_t* out_valid)
{
    DevBuf<double> partial(kRedBlocks);
    DevBuf<uint32_t> pvalid(kRedBlocks);
    mean_partial_kernel<<<kRedBlocks, 256, 0, ix.stream>>>(v, n, partial.get(), pvalid.get());
    PCPX_CHECK_LAUNCH();
    mean_final_kernel<<<1, 1, 0, ix.stream>>>(partial.get(), pvalid.get(), out_sum, out_valid);
    PCPX_CHECK_LAUNCH();
    PCPX_CUDA(cudaStreamSynchronize(ix.stream));
}

} // namespace pcpx
