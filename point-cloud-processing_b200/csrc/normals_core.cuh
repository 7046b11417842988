// Per-query epilogues on top of the top-K list: PCA normal / tangent plane and mean neighbour
// distance.  __host__ __device__ (see grid_core.cuh).
#pragma once
#include "eig3.cuh"
#include "knn_core.cuh"

namespace pcpx {

// common/normals/normal_estimation.hpp:41-77 on the neighbours whose SORTED POSITIONS are the
// low words of top.a[0..k): fp32 mean (sum / n), centred un-normalised scatter, eigenvector of
// the smallest eigenvalue.  Returns the number of neighbours used.
template <int K>
PCPX_HD uint32_t normal_from_positions(const GridView& g, const TopK<K>& top, uint32_t k,
                                       float* n3, float* centroid3, float* gap)
{
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t n = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if ((uint32_t)j < k && top.a[j] != kEmptyEntry)
        {
            float4 const c = load_pt(g.pts + (uint32_t)top.a[j]);
            sx += c.x, sy += c.y, sz += c.z;
            ++n;
        }
    float const inv = 1.f / (float)n; // n == 0 -> inf, mean NaN like Eigen's empty mean()
    float const mx = sx * inv, my = sy * inv, mz = sz * inv;
    Sym3 m{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < K; ++j)
        if ((uint32_t)j < k && top.a[j] != kEmptyEntry)
        {
            float4 const c = load_pt(g.pts + (uint32_t)top.a[j]);
            float const x = c.x - mx, y = c.y - my, z = c.z - mz;
            m.xx += x * x, m.xy += x * y, m.xz += x * z;
            m.yy += y * y, m.yz += y * z, m.zz += z * z;
        }
    smallest_eigenvector(m, n3[0], n3[1], n3[2], gap);
    if (centroid3)
        centroid3[0] = mx, centroid3[1] = my, centroid3[2] = mz;
    return n;
}

// Rare path of the normals kernel: the k-th and (k+1)-th distances are bit-equal, so WHICH of
// the tied points belongs to the neighbourhood is decided by the original index (the parity
// contract), which the position-keyed list does not carry.  `ids` is the exact (d2, original
// index) list; the block at `level` is walked once more and the members are accumulated in
// two passes exactly like normal_from_positions.
template <int K, class F>
PCPX_HD void for_each_in_block(const GridView& g, const QueryCell& qc, int level, F&& f)
{
    int const sh = g.lcap - level;
    uint32_t const cx = qc.ux >> sh, cy = qc.uy >> sh, cz = qc.uz >> sh;
    uint32_t const last = (1u << level) - 1u;
    uint64_t const key0 = cell_key(level, cx, cy, cz);
    for (int i = 0; i < 27; ++i)
    {
        Offset3 const o = block27_offset(i);
        int const dx = o.dx, dy = o.dy, dz = o.dz;
        if ((dx < 0 && cx == 0u) || (dx > 0 && cx == last) || (dy < 0 && cy == 0u) ||
            (dy > 0 && cy == last) || (dz < 0 && cz == 0u) || (dz > 0 && cz == last))
            continue;
        uint32_t start, count;
        if (!find_cell(g, key0 + key_delta(dx, dy, dz), start, count))
            continue;
        for (uint32_t p = start; p < start + count; ++p)
            f(load_pt(g.pts + p));
    }
}

template <int K>
PCPX_HD bool is_member(const TopK<K>& ids, uint32_t k, uint32_t id)
{
    bool m = false;
#pragma unroll
    for (int j = 0; j < K; ++j)
        m = m || ((uint32_t)j < k && ids.a[j] != kEmptyEntry && (uint32_t)ids.a[j] == id);
    return m;
}

template <int K>
PCPX_HD uint32_t normal_from_ids(const GridView& g, const QueryCell& qc, int level,
                                 const TopK<K>& ids, uint32_t k, float* n3, float* centroid3,
                                 float* gap)
{
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t n = 0;
    for_each_in_block<K>(g, qc, level, [&](float4 const& c) {
        if (is_member(ids, k, f2u(c.w)))
        {
            sx += c.x, sy += c.y, sz += c.z;
            ++n;
        }
    });
    float const inv = 1.f / (float)n;
    float const mx = sx * inv, my = sy * inv, mz = sz * inv;
    Sym3 m{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for_each_in_block<K>(g, qc, level, [&](float4 const& c) {
        if (is_member(ids, k, f2u(c.w)))
        {
            float const x = c.x - mx, y = c.y - my, z = c.z - mz;
            m.xx += x * x, m.xy += x * y, m.xz += x * z;
            m.yy += y * y, m.yz += y * z, m.zz += z * z;
        }
    });
    smallest_eigenvector(m, n3[0], n3[1], n3[2], gap);
    if (centroid3)
        centroid3[0] = mx, centroid3[1] = my, centroid3[2] = mz;
    return n;
}

// algorithm/average_distance_to_neighbors.hpp:56-70: sequential fp32 sum of sqrt(d2), nearest ->
// furthest, divided by the neighbour count (0 neighbours -> 0/0 = NaN like the reference).
template <int K>
PCPX_HD float mean_distance(const TopK<K>& top, uint32_t k)
{
    float sum  = 0.f;
    uint32_t n = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if ((uint32_t)j < k && top.a[j] != kEmptyEntry)
        {
#ifdef __CUDA_ARCH__
            sum = __fadd_rn(sum, __fsqrt_rn(u2f((uint32_t)(top.a[j] >> 32))));
#else
            sum = sum + sqrtf(u2f((uint32_t)(top.a[j] >> 32)));
#endif
            ++n;
        }
#ifdef __CUDA_ARCH__
    return __fdiv_rn(sum, (float)n);
#else
    return sum / (float)n;
#endif
}

} // namespace pcpx
