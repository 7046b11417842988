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
// stop early.  The cells of the block whose conservative lower bound exceeds r * r are dropped
// before they are looked up (the unrolled lookups of knn_core.cuh: collect_block27); the
// remaining spans are walked FLAT — one loop over all candidates of the lane, the next point
// already in flight — so a warp pays max over lanes of the candidate total, not the sum over
// cells of the max over lanes of the cell size.  Visiting order: cells in block order (own,
// faces, edges, corners), points of a cell in sorted order.
template <class F>
PCPX_HD void radius_visit(const GridView& g, float qx, float qy, float qz, float r, F&& f)
{
    float const rr     = fmul_x(r, r);
    QueryCell const qc = query_cell(g, qx, qy, qz);
    BlockGeom b;
    int const l = radius_level(g, qc, qx, qy, qz, r, rr, b);
    CellList cl;
    collect_block27(g, b, l, cl, nullptr, rr);
    int e = 0;
    uint32_t p = 0, pend = 0;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;)
    {
        if (p >= pend)
        {
            uint32_t const s = cl.start[e], en = cl.end[e];
            ++e;
            if (s == kSpanEnd) // the sentinel record: every span has been walked
                return;
            p = s, pend = en;
            c = load_pt(g.pts + p);
        }
        float4 const a = c;
        c = load_pt(g.pts + p + 1); // unconditional: the array is padded (kPtsPad)
        float const d2 = sqdist_x(fsub_x(a.x, qx), fsub_x(a.y, qy), fsub_x(a.z, qz));
        if (d2 <= rr)
            if (f(a, p))
                return;
        ++p;
    }
}

// The same visit with the cells looked up one at a time, each walked before the next is looked
// up.  For callers that usually stop after a few hits (the density filter's threshold) the 27
// lookups up front of radius_visit are wasted work: density filter on the 10 M noise mix 1.31 ms
// this way against 2.07 ms with the flat walk; WLOP (two visits per point, long callbacks)
// 95 ms against 112 ms.
template <class F>
PCPX_HD void radius_visit_lazy(const GridView& g, float qx, float qy, float qz, float r, F&& f)
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
        if (outside_block_near(b, o.dx, o.dy, o.dz) || cell_lb2_near(b, o.dx, o.dy, o.dz) > rr)
            continue;
        uint32_t start, count;
        if (!find_cell(g, key0 + key_delta(o.dx, o.dy, o.dz), start, count))
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
