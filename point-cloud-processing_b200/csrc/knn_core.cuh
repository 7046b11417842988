// Per-query exact kNN over the multi-level grid: register-resident sorted top-K list and the
// 3x3x3 block walk with conservative float pruning.  __host__ __device__ so tests/emu can run
// it on the CPU; the product only reaches it through the __global__ kernels in query_body.inc.
//
// Result contract (what the reference's octree/kd-tree return, made deterministic):
//   the k eligible points with the smallest (fp32 squared distance, tie-break id), ascending;
//   eligible = NOT (|dx| < eps && |dy| < eps && |dz| < eps)
//   (octree/linked_octree_node.hpp:525-543, kdtree/linked_kdtree.hpp:461-488).
#pragma once
#include "grid_core.cuh"

namespace pcpx {

constexpr uint64_t kEmptyEntry = 0xFFFFFFFFFFFFFFFFull;

// What the low 32 bits of a 64-bit list entry carry (and therefore what breaks distance ties).
enum TieId
{
    TIE_ORIGINAL_INDEX = 0 // the parity contract: (d2, original index)
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

// Geometry of the query inside its own cell at one level: conservative distances to the six
// faces of the cell (for per-cell bounds) and to the faces of the 3x3x3 block (termination).
struct BlockGeom
{
    uint32_t cx, cy, cz, last; // cell coordinates, 2^l - 1
    float sm[3], sp[3];        // squared conservative distance to the slab of cells at offset -1 / +1
    float sm2[3], sp2[3];      // ... at offset -2 / +2
    float block_lb2;           // squared lower bound on the distance to anything outside the 3^3 block
    float block_lb2_r2;        // ... outside the 5^3 block
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
    float lb1 = INFINITY, lb2 = INFINITY;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
    {
        float const lo = o[ax] + (float)cc[ax] * h;
        float fm       = (q[ax] - lo) - d2x;       // to the low face of the own cell
        float fp       = ((lo + h) - q[ax]) - d2x; // to the high face
        fm             = fm > 0.f ? fm : 0.f;
        fp             = fp > 0.f ? fp : 0.f;
        // one cell further out: the slabs at offset +-2 and the faces of the 3^3 block
        float const gm = fmaxf(fm + h - d2x, 0.f), gp = fmaxf(fp + h - d2x, 0.f);
        // two cells further out: the faces of the 5^3 block
        float const hm = fmaxf(gm + h - d2x, 0.f), hp = fmaxf(gp + h - d2x, 0.f);
        b.sm[ax] = fmul_x(fm, fm), b.sp[ax] = fmul_x(fp, fp);
        b.sm2[ax] = fmul_x(gm, gm), b.sp2[ax] = fmul_x(gp, gp);
        // a block face on (or beyond) the grid boundary constrains nothing
        lb1 = fminf(lb1, fminf(cc[ax] == 0u ? INFINITY : gm, cc[ax] == b.last ? INFINITY : gp));
        lb2 = fminf(lb2, fminf(cc[ax] <= 1u ? INFINITY : hm,
                               cc[ax] + 1u >= b.last ? INFINITY : hp));
    }
    b.block_lb2    = fmul_x(lb1, lb1);
    b.block_lb2_r2 = fmul_x(lb2, lb2);
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
    uint32_t const id   = f2u(c.w);
    (void)pos;
    uint64_t const key  = ((uint64_t)f2u(d2) << 32) | id;
    if (!excluded && key < top.last())
        top.insert(key);
}

// The search proper.  On return `top` holds the K best entries (first k are the answer).
// start_level only steers the cost (a level whose cells hold a fraction of k points makes the
// first 3x3x3 block succeed for most queries); correctness does not depend on it.
// Returns the level at which the answer was found (every list entry lies in that level's block).
template <int K, int TIE>
PCPX_HD int knn_search(const GridView& g, float qx, float qy, float qz, uint32_t k, float eps,
                       int start_level, TopK<K>& top, SearchStats* st)
{
    QueryCell const qc = query_cell(g, qx, qy, qz);
    int l              = start_level;
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


// =============================================================================================
// Two-pass search (the fast path of every kNN-shaped kernel).
//
// Pass 1 keeps only the K smallest DISTANCES: a sorted fp32 list updated by a branch-free
// min/max chain — a[j] = max(a[j-1], min(a[j], d)) is the sorted insert when d is new and the
// identity when d >= a[K-1] — 2 instructions per slot and no divergence, against ~8 per slot for
// the 64-bit (distance, id) list above.  Pass 2 walks the same block again with tau = the k-th
// distance: every point with d2 <= tau is a neighbour; its rank is the number of list entries
// strictly below d2.  Bit-equal distances (inside the list or across its boundary) make ranks
// ambiguous; they are detected and such queries fall back to the exact 64-bit search, so the
// (distance, original index) contract holds for every query.
// =============================================================================================
template <int K>
struct TopD
{
    float a[K];

    PCPX_HD void reset()
    {
#pragma unroll
        for (int j = 0; j < K; ++j)
            a[j] = INFINITY;
    }
    PCPX_HD float worst() const { return a[K - 1]; }
    PCPX_HD float kth(uint32_t k) const // a[k - 1]; the launchers keep K - k <= 4
    {
        uint32_t const back = (uint32_t)K - k;
        float r             = a[K - 1];
#pragma unroll
        for (int b = 1; b <= 4; ++b)
            if (K - 1 - b >= 0)
                r = back == (uint32_t)b ? a[K - 1 - b >= 0 ? K - 1 - b : 0] : r;
        return r;
    }
    PCPX_HD void insert(float d)
    {
#pragma unroll
        for (int j = K - 1; j > 0; --j)
            a[j] = fmaxf(a[j - 1], fminf(a[j], d));
        a[0] = fminf(a[0], d);
    }
    // Two candidates at once: with lo <= hi, the j-th smallest of list + {lo, hi} is
    // max(min(a[j], lo), min(a[j-1], hi), a[j-2]) — 2 min + one 3-input max per slot for TWO
    // candidates (1.5 ALU ops per slot and candidate instead of 2).
    PCPX_HD void insert2(float d0, float d1)
    {
        float const lo = fminf(d0, d1), hi = fmaxf(d0, d1);
#pragma unroll
        for (int j = K - 1; j >= 2; --j)
            a[j] = fmaxf(fmaxf(fminf(a[j], lo), fminf(a[j - 1], hi)), a[j - 2]);
        if (K >= 2)
            a[K >= 2 ? 1 : 0] = fmaxf(fminf(a[K >= 2 ? 1 : 0], lo), fminf(a[0], hi));
        a[0] = fminf(a[0], lo);
    }
    // number of entries strictly below d (= rank of a neighbour at distance d)
    PCPX_HD uint32_t rank_of(float d) const
    {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            r += a[j] < d;
        return r;
    }
    // two of the first k entries are bit-equal (finite): ranks would collide
    PCPX_HD bool has_internal_tie(uint32_t k) const
    {
        bool t = false;
#pragma unroll
        for (int j = 0; j + 1 < K; ++j)
            t = t || ((uint32_t)(j + 1) < k && a[j] == a[j + 1] && a[j] < INFINITY);
        return t;
    }
};

PCPX_HD bool outside_axis(uint32_t c, uint32_t last, int d)
{
    return d < 0 ? c < (uint32_t)(-d) : c + (uint32_t)d > last;
}
PCPX_HD bool outside_block(const BlockGeom& b, int dx, int dy, int dz)
{
    return outside_axis(b.cx, b.last, dx) || outside_axis(b.cy, b.last, dy) ||
           outside_axis(b.cz, b.last, dz);
}
// |d| <= 1 fast paths (the 3^3 block): cheaper than the general forms below
PCPX_HD bool outside_block_near(const BlockGeom& b, int dx, int dy, int dz)
{
    return (dx < 0 && b.cx == 0u) || (dx > 0 && b.cx == b.last) || (dy < 0 && b.cy == 0u) ||
           (dy > 0 && b.cy == b.last) || (dz < 0 && b.cz == 0u) || (dz > 0 && b.cz == b.last);
}
PCPX_HD float cell_lb2_near(const BlockGeom& b, int dx, int dy, int dz)
{
    float const sx = dx < 0 ? b.sm[0] : (dx > 0 ? b.sp[0] : 0.f);
    float const sy = dy < 0 ? b.sm[1] : (dy > 0 ? b.sp[1] : 0.f);
    float const sz = dz < 0 ? b.sm[2] : (dz > 0 ? b.sp[2] : 0.f);
    return fadd_x(fadd_x(sx, sy), sz);
}
PCPX_HD float axis_lb2(const BlockGeom& b, int ax, int d)
{
    return d == 0 ? 0.f
                  : (d == -1 ? b.sm[ax]
                             : (d == 1 ? b.sp[ax] : (d < 0 ? b.sm2[ax] : b.sp2[ax])));
}
// conservative squared lower bound on the distance to any point of the cell at offset (dx,dy,dz)
PCPX_HD float cell_lb2(const BlockGeom& b, int dx, int dy, int dz)
{
    return fadd_x(fadd_x(axis_lb2(b, 0, dx), axis_lb2(b, 1, dy)), axis_lb2(b, 2, dz));
}

// Visiting order of the 5^3 block: own cell, ring 1 (6 face, 12 edge, 8 corner neighbours), then
// the 98 cells of ring 2 by increasing lower bound.  Entry = (dx+2) | (dy+2) << 3 | (dz+2) << 6.
#define PCPX_RING_CODES {146, 145, 147, 138, 154, 82, 210, 137, 139, 153, 155, 81, 83, 209, 211, 74, 90, 202, 218, 73, 75, 89, 91, 201, 203, 217, 219, 144, 130, 18, 274, 162, 148, 136, 80, 208, 152, 129, 17, 273, 161, 66, 194, 10, 266, 26, 282, 98, 226, 131, 19, 275, 163, 140, 84, 212, 156, 72, 200, 88, 216, 65, 193, 9, 265, 25, 281, 97, 225, 67, 195, 11, 267, 27, 283, 99, 227, 76, 204, 92, 220, 128, 16, 272, 160, 2, 258, 34, 290, 132, 20, 276, 164, 64, 192, 8, 264, 24, 280, 96, 224, 1, 257, 33, 289, 3, 259, 35, 291, 68, 196, 12, 268, 28, 284, 100, 228, 0, 256, 32, 288, 4, 260, 36, 292}
static const uint16_t kRingCodesHost[125] = PCPX_RING_CODES;
#ifdef __CUDACC__
static __constant__ uint16_t kRingCodesDev[125] = PCPX_RING_CODES;
#endif
constexpr int kRing1End = 27, kRing2End = 125;
PCPX_HD Offset3 ring_offset(int i)
{
#ifdef __CUDA_ARCH__
    int const c = kRingCodesDev[i];
#else
    int const c = kRingCodesHost[i];
#endif
    Offset3 o;
    o.dx = (c & 7) - 2, o.dy = ((c >> 3) & 7) - 2, o.dz = ((c >> 6) & 7) - 2;
    return o;
}

// Up to 27 non-empty cells of a block, in visiting order, with their conservative squared lower
// bounds.  Filled by a lock-step sweep of table lookups (every lane of a warp does the same thing
// at the same time); the candidate loops then run FLAT per lane over these spans, so a warp's
// cost is max over lanes of (total candidates) rather than the sum over cells of max over lanes
// of (cell size).  Lives in local memory (dynamically indexed), which L1 caches.
// Entry n is a sentinel (start = end = kSpanEnd, bound -1: never pruned) that ends the walk, so
// the walk tests no counter; entry n + 1 exists only so that the walk's one-ahead prefetch of
// the record after the sentinel stays inside the arrays.
constexpr uint32_t kSpanEnd = 0xFFFFFFFFu;
struct CellList
{
    uint32_t start[29], end[29];
    float lb2[29];
    int n;

    PCPX_HD void close(int count)
    {
        n            = count;
        start[count] = kSpanEnd, end[count] = kSpanEnd;
        lb2[count]   = -1.f;
    }
};

#ifndef PCPX_COLLECT_UNROLL
#define PCPX_COLLECT_UNROLL 27
#endif
constexpr int kCollectUnroll = PCPX_COLLECT_UNROLL;

// Rings 0-1 (the 3^3 block): every in-grid cell is looked up, in visiting order.
// `bound`: cells whose lower bound exceeds it are dropped before they cost a lookup (the radius
// search knows r * r up front; the kNN search passes nothing and prunes while it walks).
PCPX_HD void collect_block27(const GridView& g, const BlockGeom& b, int level, CellList& cl,
                             SearchStats* st, float bound = INFINITY)
{
    uint64_t const key0 = cell_key(level, b.cx, b.cy, b.cz);
    int n               = 0;
    // Fully unrolled (kCollectUnroll): offsets, key deltas and the grid-edge tests become
    // immediates, which halves the instructions of a lookup.  The candidate walk is NOT inside
    // this loop, so the unrolled code is 27 short lookups, not 27 copies of the hot loop.
#pragma unroll kCollectUnroll
    for (int i = 0; i < 27; ++i)
    {
        Offset3 const o = block27_offset(i);
        if (outside_block_near(b, o.dx, o.dy, o.dz))
            continue;
        float const lb = cell_lb2_near(b, o.dx, o.dy, o.dz);
        if (lb > bound)
            continue;
        uint32_t start, count;
        if (st)
            st->lookups++;
        if (!find_cell(g, key0 + key_delta(o.dx, o.dy, o.dz), start, count))
            continue;
        cl.start[n] = start;
        cl.end[n]   = start + count;
        cl.lb2[n]   = lb;
        ++n;
    }
    cl.close(n);
}

// Ring 2, in chunks of up to 27 spans starting at ring offset `i` (advanced): cells outside the
// grid, cells whose bound already exceeds `worst`, and empty cells are dropped — most of the 98
// cells never cost a table lookup because rings 0-1 have already tightened `worst`.
PCPX_HD void collect_ring2(const GridView& g, const BlockGeom& b, int level, int& i, float worst,
                           CellList& cl, SearchStats* st)
{
    uint64_t const key0 = cell_key(level, b.cx, b.cy, b.cz);
    int n               = 0;
#pragma unroll 1
    for (; i < kRing2End && n < 27; ++i)
    {
        Offset3 const o = ring_offset(i);
        if (outside_block(b, o.dx, o.dy, o.dz))
            continue;
        float const lb = cell_lb2(b, o.dx, o.dy, o.dz);
        if (lb > worst) // equal: a tie may hide there
            continue;
        uint32_t start, count;
        if (st)
            st->lookups++;
        if (!find_cell(g, key0 + key_delta(o.dx, o.dy, o.dz), start, count))
            continue;
        cl.start[n] = start;
        cl.end[n]   = start + count;
        cl.lb2[n]   = lb;
        ++n;
    }
    cl.close(n);
}

// Candidates that were at or below the list's worst distance when pass 1 met them — a superset
// of the final neighbours (the worst distance only shrinks) and typically ~k (1 + ln(n / k)) of
// the n candidates, so pass 2 revisits about half of them.  More than kShortMax entries set
// `overflow` and pass 2 walks the block again.
template <int CAP>
struct ShortListT
{
    static constexpr int capacity = CAP;
    uint32_t pos[CAP]; // sorted positions
    uint32_t n;
    bool overflow;

    PCPX_HD void push(uint32_t p)
    {
        if (n < (uint32_t)CAP)
            pos[n] = p;
        else
            overflow = true;
        n += 1;
    }
};
// capacity for a K-entry list: ~3 K covers k (1 + ln(n / k)) for blocks of a few k candidates
template <int K>
using ShortListFor = ShortListT<(K <= 16 ? 48 : (K <= 24 ? 72 : 96))>;

PCPX_HD float candidate_d2(const float4& c, float qx, float qy, float qz, float eps)
{
    float const dx = fsub_x(c.x, qx), dy = fsub_x(c.y, qy), dz = fsub_x(c.z, qz);
    float d2       = sqdist_x(dx, dy, dz);
    // exclusion box, common/vector3d_queries.hpp:31-35,59-63 (strict <, all axes): all three
    // below eps <=> the largest is
    if (fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz)) < eps)
        d2 = INFINITY;
    return d2;
}

// Pass 1 over one chunk of spans: distances of every candidate go through the sorted list, two
// at a time, with the next pair already in flight; candidates at or below the running worst
// distance are remembered for pass 2.  Spans whose bound exceeds the running worst are skipped.
//
// The short list's fill count lives in a register for the whole walk (as a struct member it is
// reloaded from local memory after every store into pos[]), and the two pushes of a step are
// predicated stores behind ONE capacity test instead of two branches each.
template <int K, class SL>
PCPX_HD void knn_scan_dist(const GridView& g, const CellList& cl, float qx, float qy, float qz,
                           float eps, TopD<K>& top, SL& sl, SearchStats* st)
{
    int e = 0;                // next span to enter
    uint32_t p = 0, pend = 0; // position inside the current span
    uint32_t sn = sl.n;
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    // the next span's record is fetched from local memory one span ahead, so that entering it
    // does not wait on the load; the sentinel record makes every load unconditional
    float nlb   = cl.lb2[0];
    uint32_t ns = cl.start[0], nen = cl.end[0];
    for (;;)
    {
        if (p >= pend)
        {
            uint32_t s, en;
            for (;;)
            {
                float const lb = nlb;
                s = ns, en = nen;
                ++e;
                nlb = cl.lb2[e], ns = cl.start[e], nen = cl.end[e];
                if (!(lb > top.worst())) // equal: a tie may hide there
                    break;
            }
            if (s == kSpanEnd)
                break;
            p = s, pend = en;
            c0 = load_pt(g.pts + p);
            c1 = load_pt(g.pts + p + 1);
        }
        // The loads run up to three entries past the span (and, at the very end, past the last
        // point: the array is padded by kPtsPad) rather than test for its end; what they fetch
        // there is never used — the second candidate is masked by has1 and the prefetched pair is
        // overwritten when the next span is entered.
        bool const has1 = p + 1 < pend;
        float4 const a0 = c0, a1 = c1;
        c0 = load_pt(g.pts + p + 2); // prefetch the next pair
        c1 = load_pt(g.pts + p + 3);
        float const d0 = candidate_d2(a0, qx, qy, qz, eps);
        float const d1x = candidate_d2(a1, qx, qy, qz, eps);
        float const d1  = has1 ? d1x : INFINITY;
        float const w = fminf(top.worst(), 3.402823466e+38f); // finite: +inf (excluded) never passes
        bool const k0 = d0 <= w, k1 = d1 <= w;
        if (sn + 2u <= (uint32_t)SL::capacity)
        {
            if (k0)
                sl.pos[sn] = p;
            sn += k0;
            if (k1)
                sl.pos[sn] = p + 1;
            sn += k1;
        }
        else // the list is (nearly) full: pass 2 will walk the region again
        {
            if (k0 && sn < (uint32_t)SL::capacity)
                sl.pos[sn] = p;
            sn += k0;
            if (k1 && sn < (uint32_t)SL::capacity)
                sl.pos[sn] = p + 1;
            sn += k1;
        }
        top.insert2(d0, d1);
        if (st)
            st->candidates += has1 ? 2 : 1;
        p += 2;
    }
    sl.n        = sn;
    sl.overflow = sl.overflow || sn > (uint32_t)SL::capacity;
}

// What a kNN-shaped call tries first: a (2 * rings + 1)^3 block at `level`.
struct SearchPlan
{
    int level;
    int rings; // 1 or 2
};

// One attempt: pass 1 over the block of `rings` rings at `level`.  Ring 2 is only collected
// after rings 0-1 have gone through the list, so its cells are pruned against an already tight
// worst distance BEFORE they cost a table lookup.  Returns true when the answer is final: the
// k-th distance is strictly below the distance to anything outside the block (or the level is
// the root, which holds every point).
template <int K, int RINGS, class SL>
PCPX_HD bool knn_attempt_dist(const GridView& g, const QueryCell& qc, int level, float qx,
                              float qy, float qz, uint32_t k, float eps, TopD<K>& top,
                              BlockGeom& b, CellList& cl, SL& sl, SearchStats* st)
{
    top.reset();
    sl.n        = 0;
    sl.overflow = false;
    if (st)
        st->attempts++;
    b = block_geom(g, qc, level, qx, qy, qz);
    collect_block27(g, b, level, cl, st);
    knn_scan_dist<K>(g, cl, qx, qy, qz, eps, top, sl, st);
    if (RINGS >= 2)
    {
        // ring 2 is collected only now, so that it sees the worst distance left by rings 0-1
        int i = kRing1End;
        while (i < kRing2End)
        {
            collect_ring2(g, b, level, i, top.worst(), cl, st);
            knn_scan_dist<K>(g, cl, qx, qy, qz, eps, top, sl, st);
        }
    }
    return level == 0 || top.kth(k) < (RINGS >= 2 ? b.block_lb2_r2 : b.block_lb2);
}

// Where a finished pass 1 looked: pass 2 walks the same region again when the short list
// overflowed.  walk(g, q, tau, eps, f) calls f(point, position, d2, dx, dy, dz) for every
// eligible point of the region with d2 <= tau.
template <class F>
PCPX_HD void offer_within(const float4& c, uint32_t p, float qx, float qy, float qz, float tau,
                          float eps, F&& f)
{
    float const dx = fsub_x(c.x, qx), dy = fsub_x(c.y, qy), dz = fsub_x(c.z, qz);
    float const d2 = sqdist_x(dx, dy, dz);
    bool const excluded = fabsf(dx) < eps && fabsf(dy) < eps && fabsf(dz) < eps;
    if (!excluded && d2 <= tau)
        f(c, p, d2, dx, dy, dz);
}

template <int RINGS>
struct BlockRegion
{
    BlockGeom b;
    int level;

    template <class F>
    PCPX_HD void walk(const GridView& g, float qx, float qy, float qz, float tau, float eps,
                      F&& f) const
    {
        uint64_t const key0 = cell_key(level, b.cx, b.cy, b.cz);
        int const i_end     = RINGS >= 2 ? kRing2End : kRing1End;
#pragma unroll 1
        for (int i = 0; i < i_end; ++i)
        {
            Offset3 const o = ring_offset(i);
            if (outside_block(b, o.dx, o.dy, o.dz) || cell_lb2(b, o.dx, o.dy, o.dz) > tau)
                continue;
            uint32_t start, count;
            if (!find_cell(g, key0 + key_delta(o.dx, o.dy, o.dz), start, count))
                continue;
            for (uint32_t p = start; p < start + count; ++p)
                offer_within(load_pt(g.pts + p), p, qx, qy, qz, tau, eps, f);
        }
    }
};

// Pass 2: over the short list when it is complete, else over the region again.
template <class Region, class SL, class F>
PCPX_HD void for_each_within(const GridView& g, const Region& region, const SL& sl, float qx,
                             float qy, float qz, float tau, float eps, F&& f)
{
    if (!sl.overflow)
    {
        for (uint32_t j = 0; j < sl.n; ++j)
        {
            uint32_t const p = sl.pos[j];
            offer_within(load_pt(g.pts + p), p, qx, qy, qz, tau, eps, f);
        }
        return;
    }
    region.walk(g, qx, qy, qz, tau, eps, f);
}

} // namespace pcpx
