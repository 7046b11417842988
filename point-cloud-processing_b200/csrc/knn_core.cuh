// Per-query exact kNN over the multi-level grid: register-resident sorted top-K list and the
// 3x3x3 block walk with conservative float pruning.  __host__ __device__ so tests/emu can run
// it on the CPU; the product only reaches it through the __global__ kernels in query.cu.
//
// Result contract (what the reference's octree/kd-tree return, made deterministic):
//   the k eligible points with the smallest (fp32 squared distance, tie-break id), ascending;
//   eligible = NOT (|dx| < eps && |dy| < eps && |dz| < eps)
//   (octree/linked_octree_node.hpp:525-543, kdtree/linked_kdtree.hpp:461-488).
#pragma once
#include "grid_core.cuh"

namespace pcpx {

constexpr uint64_t kEmptyEntry = 0xFFFFFFFFFFFFFFFFull;

// What the low 32 bits of a list entry carry (and therefore what breaks distance ties).
enum TieId
{
    TIE_ORIGINAL_INDEX = 0, // the parity contract: (d2, original index)
    TIE_SORTED_POSITION = 1 // (d2, position in the Morton-sorted array): cheap gathers
};

struct SearchStats
{
    uint32_t candidates = 0; // points whose distance was evaluated
    uint32_t lookups    = 0; // hash-table cell lookups
    uint32_t attempts   = 0; // levels tried (1 = answered at the first level)
};

// Sorted ascending; entry = (bits(d2) << 32) | id.  d2 >= 0, so unsigned order of the bits is
// the fp32 order and ONE 64-bit compare gives the lexicographic (d2, id) order.
template <int K>
struct TopK
{
    uint64_t a[K];

    PCPX_HD void reset()
    {
#pragma unroll
        for (int j = 0; j < K; ++j)
            a[j] = kEmptyEntry;
    }
    PCPX_HD uint64_t last() const { return a[K - 1]; }
    // fp32 view of the largest kept distance; NaN while the list is not full, which makes
    // every `x > worst` / `worst < x` test false (= "cannot prune / not done").
    PCPX_HD float worst_d2() const { return u2f((uint32_t)(a[K - 1] >> 32)); }
    // entry k-1 without dynamic register indexing.  The launchers round K up to a multiple of 4
    // (kNN: K >= k, normals: K >= k + 1), so the entry is one of the last five.
    PCPX_HD uint64_t kth(uint32_t k) const
    {
        uint32_t const back = (uint32_t)K - k;
        uint64_t r          = a[K - 1];
#pragma unroll
        for (int b = 1; b <= 4; ++b)
            if (K - 1 - b >= 0)
                r = back == (uint32_t)b ? a[K - 1 - b >= 0 ? K - 1 - b : 0] : r;
        return r;
    }
    // precondition: key < a[K-1]
    PCPX_HD void insert(uint64_t key)
    {
#pragma unroll
        for (int j = K - 1; j > 0; --j)
        {
            bool const lt_prev = key < a[j - 1];
            bool const lt_cur  = key < a[j];
            a[j]               = lt_prev ? a[j - 1] : (lt_cur ? key : a[j]);
        }
        a[0] = key < a[0] ? key : a[0];
    }
};

// Finest level whose own cell holds at least `min_count` points (0 = root).
PCPX_HD int select_level(const GridView& g, const QueryCell& c, uint32_t min_count,
                         SearchStats* st)
{
    int l = g.lfine;
    for (; l > 0; --l)
    {
        int const sh = g.lcap - l;
        uint32_t start, count;
        if (st)
            st->lookups++;
        if (find_cell(g, cell_key(l, c.ux >> sh, c.uy >> sh, c.uz >> sh), start, count) &&
            count >= min_count)
            break;
    }
    return l;
}

// Geometry of the query inside its own cell at one level: conservative distances to the six
// faces of the cell (for per-cell bounds) and to the faces of the 3x3x3 block (termination).
struct BlockGeom
{
    uint32_t cx, cy, cz, last; // cell coordinates, 2^l - 1
    float sm[3], sp[3];        // squared conservative distance to the -/+ neighbour slab per axis
    float block_lb2;           // squared lower bound on the distance to any point outside the block
};

PCPX_HD BlockGeom block_geom(const GridView& g, const QueryCell& c, int l, float qx, float qy,
                             float qz)
{
    BlockGeom b;
    int const sh = g.lcap - l;
    b.cx = c.ux >> sh, b.cy = c.uy >> sh, b.cz = c.uz >> sh;
    b.last          = (1u << l) - 1u;
    float const h   = ldexpf(g.extent, -l);
    float const d2x = 2.f * g.delta;
    float const q[3]  = {qx, qy, qz};
    float const o[3]  = {g.ox, g.oy, g.oz};
    uint32_t const cc[3] = {b.cx, b.cy, b.cz};
    float lb = INFINITY;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
    {
        float const lo = o[ax] + (float)cc[ax] * h;
        float fm       = (q[ax] - lo) - d2x;       // to the low face of the own cell
        float fp       = ((lo + h) - q[ax]) - d2x; // to the high face
        fm             = fm > 0.f ? fm : 0.f;
        fp             = fp > 0.f ? fp : 0.f;
        b.sm[ax]       = fmul_x(fm, fm);
        b.sp[ax]       = fmul_x(fp, fp);
        // block faces one cell further out; a face on the grid boundary constrains nothing
        float const bm = cc[ax] == 0u ? INFINITY : fmaxf(fm + h - d2x, 0.f);
        float const bp = cc[ax] == b.last ? INFINITY : fmaxf(fp + h - d2x, 0.f);
        lb             = fminf(lb, fminf(bm, bp));
    }
    b.block_lb2 = fmul_x(lb, lb);
    return b;
}

// Evaluate one candidate against the list.
template <int K, int TIE>
PCPX_HD void offer(TopK<K>& top, const float4& c, uint32_t pos, float qx, float qy, float qz,
                   float eps)
{
    float const dx = fsub_x(c.x, qx), dy = fsub_x(c.y, qy), dz = fsub_x(c.z, qz);
    float const d2 = sqdist_x(dx, dy, dz);
    // common/vector3d_queries.hpp:31-35,59-63: strict <, all three axes
    bool const excluded = fabsf(dx) < eps && fabsf(dy) < eps && fabsf(dz) < eps;
    uint32_t const id   = TIE == TIE_ORIGINAL_INDEX ? f2u(c.w) : pos;
    uint64_t const key  = ((uint64_t)f2u(d2) << 32) | id;
    if (!excluded && key < top.last())
        top.insert(key);
}

// The search proper.  On return `top` holds the K best entries (first k are the answer).
// min_count steers the start level (cells with about k/2 points make the first 3x3x3 block
// succeed for most queries); correctness does not depend on it.
// Returns the level at which the answer was found (every list entry lies in that level's block).
template <int K, int TIE>
PCPX_HD int knn_search(const GridView& g, float qx, float qy, float qz, uint32_t k, float eps,
                       uint32_t min_count, TopK<K>& top, SearchStats* st)
{
    QueryCell const qc = query_cell(g, qx, qy, qz);
    int l              = select_level(g, qc, min_count, st);
    for (;; --l)
    {
        top.reset();
        if (st)
            st->attempts++;
        BlockGeom const b  = block_geom(g, qc, l, qx, qy, qz);
        uint64_t const key0 = cell_key(l, b.cx, b.cy, b.cz);
#pragma unroll 1
        for (int i = 0; i < 27; ++i)
        {
            Offset3 const o = block27_offset(i);
            int const dx = o.dx, dy = o.dy, dz = o.dz;
            if ((dx < 0 && b.cx == 0u) || (dx > 0 && b.cx == b.last) ||
                (dy < 0 && b.cy == 0u) || (dy > 0 && b.cy == b.last) ||
                (dz < 0 && b.cz == 0u) || (dz > 0 && b.cz == b.last))
                continue;
            float const sx = dx < 0 ? b.sm[0] : (dx > 0 ? b.sp[0] : 0.f);
            float const sy = dy < 0 ? b.sm[1] : (dy > 0 ? b.sp[1] : 0.f);
            float const sz = dz < 0 ? b.sm[2] : (dz > 0 ? b.sp[2] : 0.f);
            float const lb2 = fadd_x(fadd_x(sx, sy), sz);
            if (lb2 > top.worst_d2()) // a cell at exactly the worst distance may hold a tie
                continue;
            uint32_t start, count;
            if (st)
                st->lookups++;
            if (!find_cell(g, key0 + key_delta(dx, dy, dz), start, count))
                continue;
            if (st)
                st->candidates += count;
            for (uint32_t p = start; p < start + count; ++p)
                offer<K, TIE>(top, load_pt(g.pts + p), p, qx, qy, qz, eps);
        }
        if (l == 0)
            break; // the root cell holds every indexed point
        float const kth = u2f((uint32_t)(top.kth(k) >> 32));
        if (kth < b.block_lb2) // strictly closer than anything outside the block
            break;
    }
    return l;
}

} // namespace pcpx
