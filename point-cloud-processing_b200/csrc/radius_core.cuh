// Per-query sphere range search over the multi-level grid.  Predicate: pcp::sphere_t::contains
// (common/sphere.hpp:27-35), fl(d2) <= fl(r * r), the query point itself included.
#pragma once
#include "knn_core.cuh"

namespace pcpx {

// Coarsest-needed level: the finest one whose 3x3x3 block around the query provably contains
// the whole ball (every point outside the block has fl(d2) >= block_lb2 > r*r).
PCPX_HD int radius_level(const GridView& g, const QueryCell& qc, float qx, float qy, float qz,
                         float r, float rr, BlockGeom& b)
{
    float const ratio = fabsf(r) / g.extent;
    int l;
    if (!(ratio < 1.f))
        l = 0;
    else if (!(ratio > 0.f))
        l = g.lfine;
    else
    {
        l = -ilogbf(ratio) - 1; // 2^-l > ratio  <=>  cell side > r
        l = l < 0 ? 0 : (l > g.lfine ? g.lfine : l);
    }
    for (;; --l)
    {
        b = block_geom(g, qc, l, qx, qy, qz);
        if (l == 0 || rr < b.block_lb2)
            break;
    }
    return l;
}

// Calls f(point, sorted position) for every indexed point inside the ball; f returns true to
// stop early.  Cells whose conservative lower bound exceeds r*r are skipped.
template <class F>
PCPX_HD void radius_visit(const GridView& g, float qx, float qy, float qz, float r, F&& f)
{
    float const rr     = fmul_x(r, r);
    QueryCell const qc = query_cell(g, qx, qy, qz);
    BlockGeom b;
    int const l         = radius_level(g, qc, qx, qy, qz, r, rr, b);
    uint64_t const key0 = cell_key(l, b.cx, b.cy, b.cz);
#pragma unroll 1
    for (int i = 0; i < 27; ++i)
    {
        Offset3 const o = block27_offset(i);
        int const dx = o.dx, dy = o.dy, dz = o.dz;
        if ((dx < 0 && b.cx == 0u) || (dx > 0 && b.cx == b.last) || (dy < 0 && b.cy == 0u) ||
            (dy > 0 && b.cy == b.last) || (dz < 0 && b.cz == 0u) || (dz > 0 && b.cz == b.last))
            continue;
        float const sx = dx < 0 ? b.sm[0] : (dx > 0 ? b.sp[0] : 0.f);
        float const sy = dy < 0 ? b.sm[1] : (dy > 0 ? b.sp[1] : 0.f);
        float const sz = dz < 0 ? b.sm[2] : (dz > 0 ? b.sp[2] : 0.f);
        if (fadd_x(fadd_x(sx, sy), sz) > rr)
            continue;
        uint32_t start, count;
        if (!find_cell(g, key0 + key_delta(dx, dy, dz), start, count))
            continue;
        for (uint32_t p = start; p < start + count; ++p)
        {
            float4 const c = load_pt(g.pts + p);
            float const d2 = sqdist_x(fsub_x(c.x, qx), fsub_x(c.y, qy), fsub_x(c.z, qz));
            if (d2 <= rr)
                if (f(c, p))
                    return;
        }
    }
}

} // namespace pcpx
