// Warp-cooperative exact kNN: one WARP per query, for the queries the fast passes hand on (the
// tile pass's near misses and tie cases, sparse tiles, outliers; the block search's failures).
//
// A thread-per-query search is a serial chain of dependent loads — 27 table probes, a hundred
// candidates, a second pass — and with a few hundred such queries left the GPU waits 0.2 ms for
// the slowest thread, twice (queue kernel, then retry kernel).  Here the 32 lanes share one
// query, so a round of the search is ONE memory latency:
//
//   * the octree (levels 0 .. lfine of the cell table) is walked depth-first with a per-warp
//     stack in shared memory; a step pops up to four cells and looks up their 32 children at
//     once, one per lane, each with its conservative squared lower bound (same float slack as
//     every other bound in the library, grid_core.cuh);
//   * children that are leaves (finest level, or <= 32 points) are scanned together: their points
//     are spread over the lanes by a prefix over the leaf sizes, 32 candidates per round;
//   * the current k nearest live as ONE 64-bit (bits(d2), original index) key per lane, sorted
//     across the warp; candidates that beat the current k-th key are collected one per lane and
//     merged in 32 at a time: sorted by a 15-step shuffle bitonic network, min with the reversed
//     list, 5 merge steps;
//   * the remaining children go on the stack farthest first, so the nearest is opened next, and
//     every popped cell is pruned against the k-th distance found so far (a cell whose bound
//     EQUALS it is still opened: a tie with a smaller index may hide there).
//
// The first descent starts from the 3x3x3 block of the query's cell one level above the call's
// main level — a near miss is final there after two or three rounds (k-th distance strictly
// inside the block) —; what is not restarts from a block three levels coarser (cells 8 x larger),
// and so on up to the root, each attempt pruned by the k-th distance of the one before.  Keys are
// exact and totally ordered (distinct original indices), so rows come out in the (d2, original
// index) order of the parity contract with no tie handling at all.  Device only (shuffles); the GPU parity suite runs every cloud of the
// oracle tests through it (tuning "warp_all").
#pragma once
#include "normals_core.cuh"

namespace pcpx {

constexpr int kWarpStackCap      = 576;
constexpr uint32_t kWarpLeafSize = 32;
constexpr uint32_t kFullMask     = 0xFFFFFFFFu;

struct WarpStack
{
    uint64_t key[kWarpStackCap]; // cell_key(level, cx, cy, cz)
    float lb2[kWarpStackCap];
};

#ifdef __CUDACC__

// compare-exchange with the lane `xm` away: keeps the smaller (or larger) key and its payload
__device__ __forceinline__ void warp_cex(uint64_t& key, uint32_t& pay, int xm, bool keep_min)
{
    uint64_t const ok = __shfl_xor_sync(kFullMask, key, xm);
    uint32_t const op = __shfl_xor_sync(kFullMask, pay, xm);
    bool const take   = keep_min ? ok < key : ok > key;
    key               = take ? ok : key;
    pay               = take ? op : pay;
}

// ascending across the lanes
__device__ __forceinline__ void warp_sort32(uint64_t& key, uint32_t& pay, int lane)
{
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1)
            warp_cex(key, pay, j, ((lane & k2) == 0) == ((lane & j) == 0));
}

// list (ascending across the lanes) <- the 32 smallest of list + batch
__device__ __forceinline__ void warp_merge32(uint64_t& lkey, uint32_t& lpay, uint64_t bkey,
                                             uint32_t bpay, int lane)
{
    warp_sort32(bkey, bpay, lane);
    uint64_t const rk = __shfl_sync(kFullMask, bkey, 31 - lane);
    uint32_t const rp = __shfl_sync(kFullMask, bpay, 31 - lane);
    if (rk < lkey)
        lkey = rk, lpay = rp;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1)
        warp_cex(lkey, lpay, j, (lane & j) == 0);
}

// conservative squared distance from q to the cell (level, cx, cy, cz): nominal faces in float,
// every gap reduced by 2 delta (child_axis_bounds of tree_core.cuh, one cell at a time)
__device__ __forceinline__ float warp_cell_lb2(const GridView& g, float h, uint32_t cx, uint32_t cy,
                                               uint32_t cz, float qx, float qy, float qz)
{
    float const d2x = 2.f * g.delta;
    float const lx = g.ox + (float)cx * h, ly = g.oy + (float)cy * h, lz = g.oz + (float)cz * h;
    float dx = fmaxf(fmaxf(lx - qx, qx - (lx + h)), 0.f) - d2x;
    float dy = fmaxf(fmaxf(ly - qy, qy - (ly + h)), 0.f) - d2x;
    float dz = fmaxf(fmaxf(lz - qz, qz - (lz + h)), 0.f) - d2x;
    dx = dx > 0.f ? dx : 0.f, dy = dy > 0.f ? dy : 0.f, dz = dz > 0.f ? dz : 0.f;
    return fadd_x(fadd_x(fmul_x(dx, dx), fmul_x(dy, dy)), fmul_x(dz, dz));
}

struct WarpKnn
{
    uint64_t key; // lane j: the j-th smallest (bits(d2) << 32 | original index), kEmptyEntry when none
    uint32_t pos; // its position in the sorted point array

    __device__ __forceinline__ float kth_d2(uint32_t k) const
    {
        uint64_t const kk = __shfl_sync(kFullMask, key, (int)k - 1);
        return kk == kEmptyEntry ? INFINITY : __uint_as_float((uint32_t)(kk >> 32));
    }
};

// One descent.  `ckey / clb2 / cvalid`: this lane's cell of the initial batch.  `prune`: squared
// distance no neighbour can exceed (INFINITY when unknown).  On return the list holds the k
// nearest eligible points among everything under the initial cells whose bound is <= prune.
// `min_total`: when the initial cells hold fewer points than this the descent is abandoned
// before anything is scanned (the list stays empty).
__device__ __forceinline__ void warp_descend(const GridView& g, WarpStack& st, float qx, float qy,
                                             float qz, uint32_t k, float eps, float prune,
                                             uint32_t min_total, uint64_t ckey, float clb2,
                                             bool cvalid, WarpKnn& top, int lane)
{
    top.key = kEmptyEntry, top.pos = 0u;
    int sp  = 0;
    float tau = prune;
    for (;;)
    {
        // ---- look the batch up ----
        uint32_t start = 0, count = 0;
        bool found = false;
        if (cvalid && clb2 <= tau)
            found = find_cell(g, ckey, start, count);
        if (min_total != 0u)
        {
            if (__reduce_add_sync(kFullMask, found ? count : 0u) < min_total)
                return;
            min_total = 0u;
        }
        int const level = (int)(ckey >> 57);
        bool const room = sp + 32 <= kWarpStackCap;
        bool const leaf = found && (level >= g.lfine || count <= kWarpLeafSize || !room);
        bool inner      = found && !leaf;
        // ---- scan the leaves, 32 candidates per round ----
        uint32_t const cnt = leaf ? count : 0u;
        uint32_t incl      = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const up = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o)
                incl += up;
        }
        uint32_t const excl = incl - cnt, total = __shfl_sync(kFullMask, incl, 31);
        // Candidates that beat the current k-th key ("survivors") are collected, one per lane,
        // and merged into the list 32 at a time: once the list is full only a few candidates of
        // a round survive, and a merge per round would mostly sort empty keys.
        uint64_t pkey = kEmptyEntry; // pending survivors: lanes [0, np)
        uint32_t ppos = 0u, np = 0u;
        uint64_t kk   = __shfl_sync(kFullMask, top.key, (int)k - 1);
        for (uint32_t base = 0; base < total; base += 32)
        {
            uint32_t const f = base + (uint32_t)lane;
            int src          = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
            {
                uint32_t const v = __shfl_sync(kFullMask, excl, src + step);
                src += v <= f ? step : 0;
            }
            uint32_t const s = __shfl_sync(kFullMask, start, src), e = __shfl_sync(kFullMask, excl, src);
            uint64_t bkey = kEmptyEntry;
            uint32_t bpos = 0u;
            if (f < total)
            {
                bpos           = s + (f - e);
                float4 const c = __ldg(g.pts + bpos);
                float const d2 = candidate_d2(c, qx, qy, qz, eps);
                if (d2 < INFINITY)
                    bkey = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
            }
            uint32_t const m = __ballot_sync(kFullMask, bkey < kk);
            if (__popc(m) >= 12)
            {
                // many survivors (the list is still filling): the round goes in as it is
                warp_merge32(top.key, top.pos, bkey < kk ? bkey : kEmptyEntry, bpos, lane);
                kk = __shfl_sync(kFullMask, top.key, (int)k - 1);
                continue;
            }
            for (uint32_t mm = m; mm != 0u; mm &= mm - 1u) // a few: appended one by one
            {
                int const j       = __ffs((int)mm) - 1;
                uint64_t const vk = __shfl_sync(kFullMask, bkey, j);
                uint32_t const vp = __shfl_sync(kFullMask, bpos, j);
                if ((uint32_t)lane == np)
                    pkey = vk, ppos = vp;
                if (++np == 32u)
                {
                    warp_merge32(top.key, top.pos, pkey, ppos, lane);
                    pkey = kEmptyEntry, np = 0u;
                    kk   = __shfl_sync(kFullMask, top.key, (int)k - 1);
                }
            }
        }
        if (np != 0u)
            warp_merge32(top.key, top.pos, (uint32_t)lane < np ? pkey : kEmptyEntry, ppos, lane);
        tau = fminf(tau, top.kth_d2(k));
        // ---- the other children: on the stack, farthest first ----
        inner              = inner && clb2 <= tau;
        uint32_t const im  = __ballot_sync(kFullMask, inner);
        if (im != 0u)
        {
            // rank among the cells that go on the stack, farthest first (ties: lower lane first)
            uint32_t rank = 0u;
            for (uint32_t mm = im; mm != 0u; mm &= mm - 1u)
            {
                int const j    = __ffs((int)mm) - 1;
                float const lj = __shfl_sync(kFullMask, clb2, j);
                rank += (lj > clb2 || (lj == clb2 && j < lane)) ? 1u : 0u;
            }
            if (inner)
                st.key[sp + (int)rank] = ckey, st.lb2[sp + (int)rank] = clb2;
            sp += __popc(im);
        }
        __syncwarp();
        // ---- next batch: the children of the (up to) four topmost cells still worth opening;
        // entries whose bound the k-th distance has overtaken are dropped a window at a time ----
        uint64_t pk = 0;
        float plb   = 0.f;
        int P       = 0;
        while (sp > 0)
        {
            int const idx     = sp - 1 - lane;
            bool const live   = idx >= 0 && st.lb2[idx] <= tau;
            uint32_t const m  = __ballot_sync(kFullMask, live);
            int const n_live  = __popc(m);
            if (n_live == 0)
            {
                sp = sp > 32 ? sp - 32 : 0;
                continue;
            }
            P = n_live < 4 ? n_live : 4;
            // lane of the p-th live entry (p = lane >> 3), and of the last one taken
            uint32_t t = m;
            for (int i = 0; i < (lane >> 3); ++i)
                t &= t - 1u;
            int const mine = (lane >> 3) < P ? __ffs((int)t) - 1 : -1;
            uint32_t u = m;
            for (int i = 1; i < P; ++i)
                u &= u - 1u;
            int const last = __ffs((int)u) - 1;
            if (mine >= 0)
                pk = st.key[sp - 1 - mine], plb = st.lb2[sp - 1 - mine];
            // fewer than four live entries in the window: the rest of it is dead too
            int const used = n_live <= 4 ? 32 : last + 1;
            sp             = sp > used ? sp - used : 0;
            break;
        }
        __syncwarp();
        if (P == 0)
            break;
        cvalid = (lane >> 3) < P;
        cvalid = cvalid && plb <= tau;
        int const cl = (int)(pk >> 57) + 1;
        uint32_t const cx = 2u * (uint32_t)(pk & 0x7FFFFu) + (uint32_t)(lane & 1),
                       cy = 2u * (uint32_t)((pk >> 19) & 0x7FFFFu) + (uint32_t)((lane >> 1) & 1),
                       cz = 2u * (uint32_t)((pk >> 38) & 0x7FFFFu) + (uint32_t)((lane >> 2) & 1);
        ckey = cell_key(cl, cx, cy, cz);
        clb2 = warp_cell_lb2(g, ldexpf(g.extent, -cl), cx, cy, cz, qx, qy, qz);
    }
}

// The whole search of one query by one warp: blocks of 3x3x3 cells around the query at levels
// level - 1, level - 4, ... (a block is final when the k-th distance lies strictly inside it;
// a block with fewer than k + 1 points is not even scanned), then the whole tree.  Every failed
// attempt leaves its k-th distance as the pruning bound of the next.  Returns true when the
// first attempt was final.
__device__ __forceinline__ bool warp_knn(const GridView& g, WarpStack& st, float qx, float qy,
                                         float qz, uint32_t k, float eps, int level, WarpKnn& top,
                                         int lane)
{
    QueryCell const qc = query_cell(g, qx, qy, qz);
    float prune        = INFINITY;
    bool first         = true;
    for (int ls = level - 1; ls > 0; ls -= 3, first = false)
    {
        BlockGeom const b = block_geom(g, qc, ls, qx, qy, qz);
        int const dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
        int const cx = (int)b.cx + dx, cy = (int)b.cy + dy, cz = (int)b.cz + dz;
        bool const valid = lane < 27 && cx >= 0 && cy >= 0 && cz >= 0 && cx <= (int)b.last &&
                           cy <= (int)b.last && cz <= (int)b.last;
        uint64_t const key = cell_key(ls, (uint32_t)(valid ? cx : 0), (uint32_t)(valid ? cy : 0),
                                      (uint32_t)(valid ? cz : 0));
        float const lb2 = warp_cell_lb2(g, ldexpf(g.extent, -ls), (uint32_t)(valid ? cx : 0),
                                        (uint32_t)(valid ? cy : 0), (uint32_t)(valid ? cz : 0), qx,
                                        qy, qz);
        warp_descend(g, st, qx, qy, qz, k, eps, prune, k + 1u, key, lb2, valid, top, lane);
        float const kd2 = top.kth_d2(k);
        if (kd2 < b.block_lb2) // strictly inside the block: nothing outside can tie
            return first;
        prune = fminf(prune, kd2);
    }
    warp_descend(g, st, qx, qy, qz, k, eps, prune, 0u, cell_key(0, 0u, 0u, 0u), 0.f, lane == 0,
                 top, lane);
    return false;
}

#endif // __CUDACC__

} // namespace pcpx
