// Depth-first descent of the linear octree (levels 0 .. lfine of the cell table) for the queries
// the block search cannot answer cheaply: points far from every surface (outliers, uniform
// noise), whose k-ball only becomes final at a level where the 27 cells of a block hold
// thousands of points.  The descent opens a cell's eight children nearest-first and prunes by
// the same conservative float bounds as the block walk, so it touches the few fine cells the
// ball really reaches — what the reference's best-first octree traversal does
// (octree/linked_octree_node.hpp:525-566), without its heap.
#pragma once
#include "knn_core.cuh"

namespace pcpx {

constexpr int kTreeStack         = 64;
constexpr uint32_t kTreeLeafSize = 32; // cells this small are scanned, not opened

struct TreeNode
{
    uint64_t key; // cell_key(level, cx, cy, cz)
    uint32_t start, count;
    float lb2;
};

// conservative squared distances from q to the low / high half of the cell [lo, lo + 2 h') per axis
PCPX_HD void child_axis_bounds(float q, float lo, float hc, float d2x, float& a_lo, float& a_hi)
{
    float const mid = lo + hc, hi = mid + hc;
    float dl        = fmaxf(fmaxf(lo - q, q - mid), 0.f) - d2x;
    float dh        = fmaxf(fmaxf(mid - q, q - hi), 0.f) - d2x;
    dl              = dl > 0.f ? dl : 0.f;
    dh              = dh > 0.f ? dh : 0.f;
    a_lo = fmul_x(dl, dl), a_hi = fmul_x(dh, dh);
}

// leaf(start, count) for every leaf span whose bound does not exceed bound() at the time it is
// reached; bound() may shrink while the walk proceeds.
template <class Bound, class Leaf>
PCPX_HD void tree_walk(const GridView& g, float qx, float qy, float qz, Bound&& bound, Leaf&& leaf)
{
    TreeNode stack[kTreeStack];
    int sp = 0;
    {
        uint32_t start, count;
        if (!find_cell(g, cell_key(0, 0u, 0u, 0u), start, count))
            return; // empty index
        stack[sp++] = TreeNode{cell_key(0, 0u, 0u, 0u), start, count, 0.f};
    }
    float const d2x = 2.f * g.delta;
    while (sp > 0)
    {
        TreeNode const nd = stack[--sp];
        if (nd.lb2 > bound()) // equal: a tie may hide there
            continue;
        int const level = (int)(nd.key >> 57);
        if (level >= g.lfine || nd.count <= kTreeLeafSize || sp + 8 > kTreeStack)
        {
            leaf(nd.start, nd.count);
            continue;
        }
        uint32_t const cx = (uint32_t)(nd.key & 0x7FFFFu), cy = (uint32_t)((nd.key >> 19) & 0x7FFFFu),
                       cz = (uint32_t)((nd.key >> 38) & 0x7FFFFu);
        float const hc = ldexpf(g.extent, -(level + 1)); // child cell side
        float ax[2], ay[2], az[2];
        child_axis_bounds(qx, g.ox + (float)(2u * cx) * hc, hc, d2x, ax[0], ax[1]);
        child_axis_bounds(qy, g.oy + (float)(2u * cy) * hc, hc, d2x, ay[0], ay[1]);
        child_axis_bounds(qz, g.oz + (float)(2u * cz) * hc, hc, d2x, az[0], az[1]);
        // existing children, pushed farthest first so that the nearest is opened next
        TreeNode kids[8];
        int nk = 0;
        for (int c = 0; c < 8; ++c)
        {
            int const bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
            float const lb = fadd_x(fadd_x(ax[bx], ay[by]), az[bz]);
            if (lb > bound())
                continue;
            uint64_t const key = cell_key(level + 1, 2u * cx + (uint32_t)bx, 2u * cy + (uint32_t)by,
                                          2u * cz + (uint32_t)bz);
            uint32_t start, count;
            if (!find_cell(g, key, start, count))
                continue;
            kids[nk++] = TreeNode{key, start, count, lb};
        }
        for (int placed = 0; placed < nk; ++placed)
        {
            int far = -1;
            for (int c = 0; c < nk; ++c)
                if (kids[c].count != 0u && (far < 0 || kids[c].lb2 > kids[far].lb2))
                    far = c;
            stack[sp++]     = kids[far];
            kids[far].count = 0u; // taken
        }
    }
}

// Pass 1 over the whole cloud by descent: on return `top` holds the K smallest distances and the
// answer is final (no block boundary to check).
template <int K, class SL>
PCPX_HD void knn_tree_dist(const GridView& g, float qx, float qy, float qz, float eps,
                           TopD<K>& top, SL& sl, SearchStats* st)
{
    top.reset();
    sl.n        = 0;
    sl.overflow = false;
    if (st)
        st->attempts++;
    tree_walk(
        g, qx, qy, qz, [&] { return top.worst(); },
        [&](uint32_t start, uint32_t count) {
            for (uint32_t p = start; p < start + count; ++p)
            {
                float const d2 = candidate_d2(load_pt(g.pts + p), qx, qy, qz, eps);
                if (d2 <= top.worst() && d2 < INFINITY)
                    sl.push(p);
                top.insert(d2);
            }
            if (st)
                st->candidates += count, st->lookups++;
        });
}

// the whole cloud as a pass-2 region
struct TreeRegion
{
    template <class F>
    PCPX_HD void walk(const GridView& g, float qx, float qy, float qz, float tau, float eps,
                      F&& f) const
    {
        tree_walk(
            g, qx, qy, qz, [&] { return tau; },
            [&](uint32_t start, uint32_t count) {
                for (uint32_t p = start; p < start + count; ++p)
                    offer_within(load_pt(g.pts + p), p, qx, qy, qz, tau, eps, f);
            });
    }
};

} // namespace pcpx
