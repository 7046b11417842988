// Tile-cooperative exact kNN: the fast path of every kNN-shaped call whose queries are the
// indexed points themselves (estimate_normals, mean neighbour distance, kNN graphs).
//
// Every query of one finest cell sees the same 3x3x3 block, so the per-thread search of
// knn_core.cuh re-derives 27 table lookups and a span list per THREAD that are identical for all
// queries of the cell.  Here a CTA owns a *tile* — one cell two levels above the call's main
// level, i.e. 4x4x4 main-level cells, a contiguous run of the sorted point array — and does the
// structure work ONCE per tile:
//
//   1. 216 table lookups (the 6x6x6 main-level cells: the tile plus one cell of halo on every
//      side) give the spans of every point a tile query can need;
//   2. those points are copied into shared memory RE-BINNED into a dense local grid: SA sub-bins
//      per main-level cell along a, SB along b, whole cells along the third axis (c = the axis
//      along which the tile's neighbourhood is thinnest), rows along a contiguous — so the
//      candidates of a query are a handful of contiguous shared-memory ranges, found by two
//      loads from a 864-entry start table instead of hash probes;
//   3. one thread per query lists, row by row, only the bins its current search ball touches
//      (a disc, not a box; the ball shrinks as soon as the list is full) and selects the k
//      nearest with sorting-network batches: no local memory, no per-thread span lists.
//
// Layouts (template parameter S): 1 = whole cells (1 x 1), 2 = 2 x 2 sub-bins, 4 = 4 x 1 (four
// sub-bins along the row axis, whole cells along b).  When SB == 1 a row of bins is a row of
// CELLS, so the start of every cell in the staged array follows from the cell counts alone: the
// staging is ONE pass over global memory (a warp per occupied cell, its points ranked into the
// cell's sub-bins by ballots) instead of count -> scan -> place, and the row table and the query
// segments are known before a single point has been loaded.
//
// The top-k list is KL (>= k + 1) 32-bit keys in registers: (bits(d2) & ~mask) | staged position.
// d2 >= 0, so unsigned order of the bits is the fp32 order; the low `mask` bits are replaced by
// the position of the candidate in shared memory, which makes ONE pass enough (no second pass to
// recover identities) at the price of comparing distances truncated to (23 - bits(mask)) mantissa
// bits.  That is still exact:
//   * truncation is monotone, so every member of the k smallest KEYS has d2 <= every non-member
//     — unless the k-th and (k+1)-th keys agree in all kept bits; that case is detected and the
//     query goes to the retry queue (the exact per-thread search of knn_core.cuh);
//   * the order among members is re-established from the exact distances (and original indices)
//     of the k winners when a call needs it (kNN rows, mean distance).
// A query is final when the k-th key's upper bound lies inside the ball that was scanned and that
// ball lies inside the staged region; everything else (sparse neighbourhoods, outliers, tiles
// whose region does not fit) is appended to the same retry queue.
//
// __host__ __device__ throughout (see grid_core.cuh): tests/emu runs the same phases on the CPU.
#pragma once
#include "normals_core.cuh"

namespace pcpx {

constexpr int kTileShift      = 2;                   // tile = cell at (main level - 2)
constexpr int kTileCells      = 1 << kTileShift;     // main-level cells per tile and axis
constexpr int kRegionCells    = kTileCells + 2;      // + one cell of halo on every side
constexpr int kRegionCellCount = kRegionCells * kRegionCells * kRegionCells; // 216
constexpr uint32_t kKeyEmpty  = 0xFFFFFFFFu;
constexpr uint32_t kTilePad   = 2;                   // readable entries behind the staged points
constexpr int kTileCandCap    = 32;                  // candidates one query lists per round
constexpr uint32_t kNoSelf    = 0xFFFFFFFFu;

template <int S>
struct TileDims
{
    static_assert(S == 1 || S == 2 || S == 4, "layouts: 1 (1 x 1), 2 (2 x 2), 4 (4 x 1)");
    static constexpr int SA = S, SB = S == 4 ? 1 : S; // sub-bins per cell along a / along b
    static constexpr int sub = SA * SB;
    static constexpr int na = kRegionCells * SA, nb = kRegionCells * SB, nc = kRegionCells;
    static constexpr int bins = na * nb * nc;
    // rows (fixed b bin and c cell) that hold the tile's own points, i.e. the queries
    static constexpr int segs = kTileCells * kTileCells * SB;
    static constexpr bool one_pass = SB == 1;
};

// per call, computed on the host (make_tile_params)
struct TileParams
{
    int level;           // main level L
    int rows_b, rows_c;  // rows on either side of the query's own row that a scan can touch
    uint32_t max_points; // capacity of the staged region
    uint32_t key_mask;   // low key bits that carry the staged position
    float h;             // cell side at L
    float bins_per_len;  // SA * 2^L / extent
    float len_per_bin;   // h / SA
    float bins_per_len_b; // SB * 2^L / extent
    float len_per_bin_b; // h / SB
    float cells_per_len; // 2^L / extent
    float delta_bins;    // g.delta in a bin units: slack of every float-evaluated bin bound
    float delta_bins_b;  // ... in b bin units
    float delta_cells;   // g.delta in cell units
    float scan_cap;      // largest scan radius, in units of h
    uint32_t first_cap;  // candidates listed before the ball is first shrunk
    int threads;         // threads of the CTA
    uint32_t min_queries; // tiles with fewer queries are handed on without staging
    uint32_t by_position; // output rows indexed by sorted position (QueryBatch::by_position)
};

template <int S>
inline TileParams make_tile_params(const GridView& g, int level, uint32_t max_points,
                                   float scan_cap)
{
    using D = TileDims<S>;
    TileParams tp;
    tp.level         = level;
    tp.max_points    = max_points;
    uint32_t bits    = 1;
    while ((1u << bits) < max_points + kTilePad)
        ++bits;
    tp.key_mask      = (1u << bits) - 1u;
    tp.h             = ldexpf(g.extent, -level);
    tp.cells_per_len = ldexpf(1.f, level) / g.extent;
    tp.bins_per_len  = tp.cells_per_len * (float)D::SA;
    tp.len_per_bin   = tp.h / (float)D::SA;
    tp.bins_per_len_b = tp.cells_per_len * (float)D::SB;
    tp.len_per_bin_b = tp.h / (float)D::SB;
    tp.delta_bins    = g.delta * tp.bins_per_len * 1.5f + 1e-6f;
    tp.delta_bins_b  = g.delta * tp.bins_per_len_b * 1.5f + 1e-6f;
    tp.delta_cells   = g.delta * tp.cells_per_len * 1.5f + 1e-6f;
    tp.scan_cap      = scan_cap;
    tp.rows_b        = (int)ceilf(scan_cap * (float)D::SB);
    tp.rows_c        = (int)ceilf(scan_cap);
    tp.first_cap     = (uint32_t)kTileCandCap;
    tp.threads       = 96;
    tp.min_queries   = 0;
    tp.by_position   = 0;
    return tp;
}

// per tile, written by one thread (tile_phase_plan)
struct TileGeom
{
    int32_t r0[3];      // region origin in main-level cells along a, b, c (may be -1)
    int32_t ax[3];      // world axis (0 = x, 1 = y, 2 = z) of a, b, c
    uint32_t n_occ;     // region cells that hold points
    uint32_t n_points;  // points of the region (staged)
    int32_t fallback;   // the region does not fit: every query of the tile goes to the retry queue
};

// pointers into the CTA's shared memory (or the emulator's arrays)
struct TileSmem
{
    float4* P;            // max_points + kTilePad, bin order
    uint32_t* F;          // bins + 1: start of every bin (F[bins] = n_points)
    uint32_t* cstart;     // kRegionCellCount: span of every region cell in the sorted array ...
    uint32_t* ccount;     // kRegionCellCount: ... cell id = (z * 6 + y) * 6 + x, world axes
    uint8_t* occ;         // kRegionCellCount: ids of the cells that hold points, ascending
    TileGeom* geom;
    uint32_t* rowmask;    // nc words: bit ib of word ic = row (ic, ib) holds points
    uint32_t* seg_off;    // segs + 1: prefix over the tile's own row segments (the queries)
    uint16_t* cl;         // kTileCandCap x nthreads candidate positions, [j * nthreads + tid]
    uint32_t* gpos;       // max_points: position of every staged point in the sorted global array
};

// tile coordinates (cell coordinates at the tile level), 21 bits each
PCPX_HD uint64_t tile_pack(uint32_t x, uint32_t y, uint32_t z)
{
    return (uint64_t)x | ((uint64_t)y << 21) | ((uint64_t)z << 42);
}
PCPX_HD uint32_t tile_x(uint64_t t) { return (uint32_t)t & 0x1FFFFFu; }
PCPX_HD uint32_t tile_y(uint64_t t) { return (uint32_t)(t >> 21) & 0x1FFFFFu; }
PCPX_HD uint32_t tile_z(uint64_t t) { return (uint32_t)(t >> 42) & 0x1FFFFFu; }

PCPX_HD float axis_of(float x, float y, float z, int ax) { return ax == 0 ? x : (ax == 1 ? y : z); }
PCPX_HD int32_t axis_of_i(int32_t x, int32_t y, int32_t z, int ax)
{
    return ax == 0 ? x : (ax == 1 ? y : z);
}
// weight of world axis `ax` in a region cell id ((z * 6 + y) * 6 + x)
PCPX_HD int tile_axis_weight(int ax)
{
    return ax == 0 ? 1 : (ax == 1 ? kRegionCells : kRegionCells * kRegionCells);
}

// ---- staging phases (a barrier between consecutive phases) -----------------------------------
//
//   lookup  one table lookup per region cell (216, main level): the spans of exactly the points a
//           tile query can need;
//   plan    the occupied cells compacted (device: ballots by the first warp), the thin axis chosen,
//           the tile handed on when its region does not fit.  One-pass layouts (SB == 1): the same
//           warp goes on to scan the cell counts in (c, b, a) order — F[first bin of a cell] = the
//           cell's start in the staged array —, and derives the row table and the query segments;
//   then, one-pass layouts:
//   place   a warp per occupied cell: loads the cell's points (the loads of four cells in flight
//           at once), ranks them into the cell's sub-bins by ballots, writes them and the
//           sub-bin starts;
//   other layouts:
//   count   one thread per occupied cell: how many of its points fall into each of its sub-bins
//           (every bin belongs to exactly one cell: plain stores, no atomics);
//   scan    inclusive scan of the bin counts: F[i] = start of bin i;
//   place   the same thread copies its cell's points to their bins in the order of the sorted
//           array — the staged order is a function of the input alone;
//   rows    row table and query segments.

// phase 0: look up the region cells (other layouts: and clear the bin table)
template <int S>
PCPX_HD void tile_phase_lookup(const GridView& g, const TileParams& tp, const TileSmem& sm,
                               uint64_t tile_xyz, int tid, int nthreads)
{
    if (!TileDims<S>::one_pass)
    {
        for (int i = tid; i <= TileDims<S>::bins; i += nthreads)
            sm.F[i] = 0u;
        if (tid < TileDims<S>::nc)
            sm.rowmask[tid] = 0u;
    }
    int const last = (1 << tp.level) - 1;
    int const x0 = kTileCells * (int)tile_x(tile_xyz) - 1, y0 = kTileCells * (int)tile_y(tile_xyz) - 1,
              z0 = kTileCells * (int)tile_z(tile_xyz) - 1;
    // three cells per thread and round: their first probes are in flight together
    constexpr int U = 3;
    for (int c0 = tid; c0 < kRegionCellCount; c0 += nthreads * U)
    {
        uint32_t slot[U], klo[U], khi[U];
        HashSlot sl[U];
        bool in[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            int const c  = c0 + u * nthreads;
            int const cx = x0 + c % kRegionCells, cy = y0 + (c / kRegionCells) % kRegionCells,
                      cz = z0 + c / (kRegionCells * kRegionCells);
            in[u] = c < kRegionCellCount && cx >= 0 && cy >= 0 && cz >= 0 && cx <= last &&
                    cy <= last && cz <= last;
            uint64_t const key =
                cell_key(tp.level, (uint32_t)(in[u] ? cx : 0), (uint32_t)(in[u] ? cy : 0),
                         (uint32_t)(in[u] ? cz : 0));
            slot[u] = hash_slot(key, g.table_size);
            klo[u] = (uint32_t)key, khi[u] = (uint32_t)(key >> 32);
            if (in[u])
                sl[u] = load_slot(g.table + slot[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            int const c = c0 + u * nthreads;
            if (c >= kRegionCellCount)
                continue;
            uint32_t start = 0, count = 0;
            if (in[u])
                for (;;) // linear probing (find_cell), the first probe already loaded
                {
                    if (sl[u].key_hi == khi[u] && sl[u].key_lo == klo[u])
                    {
                        start = sl[u].start, count = sl[u].count;
                        break;
                    }
                    if (sl[u].key_hi == kEmptyKeyHi)
                        break;
                    slot[u] = slot[u] + 1 == g.table_size ? 0u : slot[u] + 1;
                    sl[u]   = load_slot(g.table + slot[u]);
                }
            sm.cstart[c] = start;
            sm.ccount[c] = count;
        }
    }
}

// One-pass layouts, first warp (device) / one thread (host), after the plan is written: cell
// starts, row table, query segments.  F[j * SA + s] = start of cell j (j = (ic * 6 + ib) * 6 + ia)
// for every sub-bin s; the place phase overwrites s >= 1 of the occupied cells.
template <int S>
PCPX_HD void tile_plan_prefix(const TileSmem& sm, int lane)
{
    using D = TileDims<S>;
    TileGeom const tg = *sm.geom;
    if (tg.fallback != 0)
        return;
    int const wa = tile_axis_weight(tg.ax[0]), wb = tile_axis_weight(tg.ax[1]),
              wc = tile_axis_weight(tg.ax[2]);
    auto count_of = [&](int j) {
        int const ia = j % kRegionCells, ib = (j / kRegionCells) % kRegionCells,
                  ic = j / (kRegionCells * kRegionCells);
        return sm.ccount[ia * wa + ib * wb + ic * wc];
    };
#ifdef __CUDA_ARCH__
    constexpr int per = (kRegionCellCount + 31) / 32;
    uint32_t cnt[per], s = 0;
#pragma unroll
    for (int i = 0; i < per; ++i)
    {
        int const j = lane * per + i;
        cnt[i]      = j < kRegionCellCount ? count_of(j) : 0u;
        s += cnt[i];
    }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o)
            incl += up;
    }
    uint32_t run = incl - s;
#pragma unroll
    for (int i = 0; i < per; ++i)
    {
        int const j = lane * per + i;
        if (j < kRegionCellCount)
        {
#pragma unroll
            for (int q = 0; q < D::SA; ++q)
                sm.F[j * D::SA + q] = run;
        }
        run += cnt[i];
    }
    if (lane == 31)
        sm.F[D::bins] = run;
    __syncwarp();
    if (lane < D::nc)
    {
        uint32_t m = 0;
        for (int ib = 0; ib < D::nb; ++ib)
        {
            int const r = lane * D::nb + ib;
            if (sm.F[(r + 1) * D::na] > sm.F[r * D::na])
                m |= 1u << ib;
        }
        sm.rowmask[lane] = m;
    }
    {
        uint32_t len = 0;
        if (lane < D::segs)
        {
            int const base = ((1 + lane / kTileCells) * D::nb + 1 + lane % kTileCells) * D::na;
            len            = sm.F[base + D::SA + kTileCells * D::SA] - sm.F[base + D::SA];
        }
        uint32_t sc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, sc, o);
            if (lane >= o)
                sc += up;
        }
        if (lane < D::segs)
            sm.seg_off[lane] = sc - len;
        if (lane == D::segs - 1)
            sm.seg_off[D::segs] = sc;
    }
#else
    (void)lane;
    uint32_t run = 0;
    for (int j = 0; j < kRegionCellCount; ++j)
    {
        for (int q = 0; q < D::SA; ++q)
            sm.F[j * D::SA + q] = run;
        run += count_of(j);
    }
    sm.F[D::bins] = run;
    for (int ic = 0; ic < D::nc; ++ic)
    {
        uint32_t m = 0;
        for (int ib = 0; ib < D::nb; ++ib)
            if (sm.F[(ic * D::nb + ib + 1) * D::na] > sm.F[(ic * D::nb + ib) * D::na])
                m |= 1u << ib;
        sm.rowmask[ic] = m;
    }
    uint32_t q = 0;
    for (int seg = 0; seg < D::segs; ++seg)
    {
        int const base  = ((1 + seg / kTileCells) * D::nb + 1 + seg % kTileCells) * D::na;
        sm.seg_off[seg] = q;
        q += sm.F[base + D::SA + kTileCells * D::SA] - sm.F[base + D::SA];
    }
    sm.seg_off[D::segs] = q;
#endif
}

// phase 1: compact the occupied cells, choose the axes (+ tile_plan_prefix)
template <int S>
PCPX_HD void tile_phase_plan(const TileParams& tp, const TileSmem& sm, uint64_t tile_xyz, int tid)
{
    uint32_t occ[3] = {0u, 0u, 0u}; // bit i of occ[axis]: some cell at coordinate i holds points
    uint32_t n = 0, total = 0;
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 800
    if (tid >= 32)
        return;
    for (int c0 = 0; c0 < kRegionCellCount; c0 += 32)
    {
        int const c        = c0 + tid;
        uint32_t const cnt = c < kRegionCellCount ? sm.ccount[c] : 0u;
        uint32_t const m   = __ballot_sync(0xFFFFFFFFu, cnt != 0u);
        if (cnt)
            sm.occ[n + (uint32_t)__popc(m & ((1u << tid) - 1u))] = (uint8_t)c;
        n += (uint32_t)__popc(m);
        total += __reduce_add_sync(0xFFFFFFFFu, cnt);
        occ[0] |= __reduce_or_sync(0xFFFFFFFFu, cnt ? 1u << (c % kRegionCells) : 0u);
        occ[1] |= __reduce_or_sync(0xFFFFFFFFu, cnt ? 1u << ((c / kRegionCells) % kRegionCells) : 0u);
        occ[2] |= __reduce_or_sync(0xFFFFFFFFu, cnt ? 1u << (c / (kRegionCells * kRegionCells)) : 0u);
    }
#else
    if (tid != 0)
        return;
    for (int c = 0; c < kRegionCellCount; ++c)
    {
        uint32_t const cnt = sm.ccount[c];
        if (cnt == 0)
            continue;
        occ[0] |= 1u << (c % kRegionCells), occ[1] |= 1u << ((c / kRegionCells) % kRegionCells);
        occ[2] |= 1u << (c / (kRegionCells * kRegionCells));
        sm.occ[n++] = (uint8_t)c;
        total += cnt;
    }
#endif
    if (tid == 0)
    {
        auto pop = [](uint32_t m) {
            uint32_t r = 0;
            for (int i = 0; i < kRegionCells; ++i)
                r += (m >> i) & 1u;
            return r;
        };
        uint32_t const ex = pop(occ[0]), ey = pop(occ[1]), ez = pop(occ[2]);
        // c = the thinnest axis (ties: z, then y), a = the widest of the other two (ties: the lower)
        int c = 2;
        if (ey < ez)
            c = 1;
        if (ex < (c == 2 ? ez : ey))
            c = 0;
        int a = c == 0 ? 1 : 0, b = c == 2 ? 1 : 2;
        uint32_t const ea = a == 0 ? ex : ey, eb = b == 1 ? ey : ez;
        if (eb > ea)
        {
            int const t = a;
            a = b, b = t;
        }
        TileGeom& tg = *sm.geom;
        tg.ax[0] = a, tg.ax[1] = b, tg.ax[2] = c;
        int32_t const tx = (int32_t)tile_x(tile_xyz), ty = (int32_t)tile_y(tile_xyz),
                      tz = (int32_t)tile_z(tile_xyz);
        tg.r0[0] = kTileCells * axis_of_i(tx, ty, tz, a) - 1;
        tg.r0[1] = kTileCells * axis_of_i(tx, ty, tz, b) - 1;
        tg.r0[2] = kTileCells * axis_of_i(tx, ty, tz, c) - 1;
        tg.n_occ    = n;
        tg.n_points = total;
        tg.fallback = total > tp.max_points ? 1 : 0;
    }
    if constexpr (TileDims<S>::one_pass)
    {
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
        tile_plan_prefix<S>(sm, tid);
    }
}

// What a thread needs to bin the points of one cell: the first bin of the cell and the sub-bin
// of a point inside its cell.  (x - o) * (scale * SA) == SA * ((x - o) * scale) exactly (SA a
// power of two), so a point's sub-bin always lies inside the cell the index assigned it to.
template <int S>
struct TileCellBinner
{
    using D = TileDims<S>;
    float oa, ob, scale_a, scale_b;
    uint32_t top_a, top_b;
    int sh, A, B;
    int wx, wy, wz; // weight of a cell's world coordinates in its first bin's index

    PCPX_HD TileCellBinner(const GridView& g, const TileParams& tp, const TileGeom& tg)
    {
        A = tg.ax[0], B = tg.ax[1];
        oa = axis_of(g.ox, g.oy, g.oz, A), ob = axis_of(g.ox, g.oy, g.oz, B);
        scale_a = g.scale * (float)D::SA, scale_b = g.scale * (float)D::SB;
        top_a   = ((1u << g.lcap) * (uint32_t)D::SA) - 1u;
        top_b   = ((1u << g.lcap) * (uint32_t)D::SB) - 1u;
        sh      = g.lcap - tp.level;
        auto weight = [&](int axis) {
            return A == axis ? D::SA : (B == axis ? D::SB * D::na : D::na * D::nb);
        };
        wx = weight(0), wy = weight(1), wz = weight(2);
    }
    PCPX_HD int first_bin(int cell) const
    {
        return (cell % kRegionCells) * wx + ((cell / kRegionCells) % kRegionCells) * wy +
               (cell / (kRegionCells * kRegionCells)) * wz;
    }
    PCPX_HD static uint32_t part(float x, float o, float scale, uint32_t top, int sh, uint32_t n)
    {
        float t = (x - o) * scale;
        t       = t > 0.f ? t : 0.f;
        uint32_t u = (uint32_t)t;
        u          = u < top ? u : top;
        return (u >> sh) & (n - 1u);
    }
    PCPX_HD uint32_t sub_a(const float4& p) const
    {
        return D::SA == 1 ? 0u : part(axis_of(p.x, p.y, p.z, A), oa, scale_a, top_a, sh, D::SA);
    }
    PCPX_HD uint32_t sub_b(const float4& p) const
    {
        return D::SB == 1 ? 0u : part(axis_of(p.x, p.y, p.z, B), ob, scale_b, top_b, sh, D::SB);
    }
    // 0 .. SA * SB - 1: sub-bin of a point inside its cell, a fastest
    PCPX_HD uint32_t sub(const float4& p) const { return sub_a(p) + (uint32_t)D::SA * sub_b(p); }
    // offset of sub-bin s from the cell's first bin
    PCPX_HD static int sub_offset(uint32_t s)
    {
        return (int)(s % (uint32_t)D::SA) + (int)(s / (uint32_t)D::SA) * D::na;
    }
};

// One-pass layouts: the place phase.  Device: warp `warp` of `nwarps` takes the occupied cells
// warp, warp + nwarps, ...; lane i holds point i of the cell.  Host (one emulated thread at a
// time): thread tid takes the cells tid, tid + nthreads, ... serially; same staged order.
template <int S>
PCPX_HD void tile_phase_place_cells(const GridView& g, const TileParams& tp, const TileSmem& sm,
                                    int tid, int nthreads)
{
    using D = TileDims<S>;
    static_assert(D::one_pass && D::SA <= 4, "cells are rows of bins");
    TileGeom const tg = *sm.geom;
    TileCellBinner<S> const binner(g, tp, tg);
    if (tid == 0)
        for (uint32_t j = 0; j < kTilePad; ++j)
            sm.P[tg.n_points + j] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
#ifdef __CUDA_ARCH__
    int const lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    uint32_t const lt = (1u << lane) - 1u;
    constexpr int U   = 4;
    for (uint32_t j0 = (uint32_t)warp; j0 < tg.n_occ; j0 += (uint32_t)(nwarps * U))
    {
        float4 p[U];
        uint32_t start[U], count[U];
        int bin[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            uint32_t const j = j0 + (uint32_t)(u * nwarps);
            count[u]         = 0u;
            if (j < tg.n_occ)
            {
                int const cell = sm.occ[j];
                start[u] = sm.cstart[cell], count[u] = sm.ccount[cell];
                bin[u]   = binner.first_bin(cell);
                if ((uint32_t)lane < count[u])
                    p[u] = load_pt(g.pts + start[u] + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (count[u] == 0u)
                continue;
            // first chunk from the registers; longer cells: a counting sweep, then the placing sweep
            bool const v0     = (uint32_t)lane < count[u];
            uint32_t const s0 = v0 ? binner.sub_a(p[u]) : 0xFFu;
            uint32_t m[4]     = {0u, 0u, 0u, 0u}, tot[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int q = 0; q < D::SA; ++q)
            {
                m[q]   = __ballot_sync(0xFFFFFFFFu, s0 == (uint32_t)q);
                tot[q] = (uint32_t)__popc(m[q]);
            }
            for (uint32_t c = 32; c < count[u]; c += 32)
            {
                bool const v     = c + (uint32_t)lane < count[u];
                uint32_t const s = v ? binner.sub_a(load_pt(g.pts + start[u] + c + lane)) : 0xFFu;
#pragma unroll
                for (int q = 0; q < D::SA; ++q)
                    tot[q] += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, s == (uint32_t)q));
            }
            uint32_t off[4];
            off[0] = sm.F[bin[u]];
#pragma unroll
            for (int q = 1; q < 4; ++q)
                off[q] = off[q - 1] + tot[q - 1];
            if (lane >= 1 && lane < D::SA)
                sm.F[bin[u] + lane] = lane == 1 ? off[1] : (lane == 2 ? off[2] : off[3]);
            if (v0)
            {
                uint32_t const base = s0 == 0u ? off[0] : (s0 == 1u ? off[1] : (s0 == 2u ? off[2] : off[3]));
                uint32_t const mm   = s0 == 0u ? m[0] : (s0 == 1u ? m[1] : (s0 == 2u ? m[2] : m[3]));
                uint32_t const pos  = base + (uint32_t)__popc(mm & lt);
                sm.P[pos]           = p[u];
                sm.gpos[pos]        = start[u] + (uint32_t)lane;
            }
#pragma unroll
            for (int q = 0; q < D::SA; ++q)
                off[q] += (uint32_t)__popc(m[q]);
            for (uint32_t c = 32; c < count[u]; c += 32)
            {
                bool const v = c + (uint32_t)lane < count[u];
                float4 pp    = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v)
                    pp = load_pt(g.pts + start[u] + c + lane);
                uint32_t const s = v ? binner.sub_a(pp) : 0xFFu;
                uint32_t mc[4]   = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int q = 0; q < D::SA; ++q)
                    mc[q] = __ballot_sync(0xFFFFFFFFu, s == (uint32_t)q);
                if (v)
                {
                    uint32_t const base = s == 0u ? off[0] : (s == 1u ? off[1] : (s == 2u ? off[2] : off[3]));
                    uint32_t const mm   = s == 0u ? mc[0] : (s == 1u ? mc[1] : (s == 2u ? mc[2] : mc[3]));
                    uint32_t const pos  = base + (uint32_t)__popc(mm & lt);
                    sm.P[pos]           = pp;
                    sm.gpos[pos]        = start[u] + c + (uint32_t)lane;
                }
#pragma unroll
                for (int q = 0; q < D::SA; ++q)
                    off[q] += (uint32_t)__popc(mc[q]);
            }
        }
    }
#else
    for (uint32_t j = (uint32_t)tid; j < tg.n_occ; j += (uint32_t)nthreads)
    {
        int const cell       = sm.occ[j];
        uint32_t const start = sm.cstart[cell], count = sm.ccount[cell];
        int const bin        = binner.first_bin(cell);
        uint32_t off[4] = {sm.F[bin], 0u, 0u, 0u}, tot[4] = {0u, 0u, 0u, 0u};
        for (uint32_t i = 0; i < count; ++i)
            tot[binner.sub_a(load_pt(g.pts + start + i))]++;
        for (int q = 1; q < 4; ++q)
            off[q] = off[q - 1] + tot[q - 1];
        for (int q = 1; q < D::SA; ++q)
            sm.F[bin + q] = off[q];
        for (uint32_t i = 0; i < count; ++i)
        {
            float4 const p     = load_pt(g.pts + start + i);
            uint32_t const pos = off[binner.sub_a(p)]++;
            sm.P[pos]          = p;
            sm.gpos[pos]       = start + i;
        }
    }
#endif
}

// phase 2 (other layouts): bin counts (F[bin + 1] = count), one thread per occupied cell
template <int S>
PCPX_HD void tile_phase_count(const GridView& g, const TileParams& tp, const TileSmem& sm, int tid,
                              int nthreads)
{
    using D = TileDims<S>;
    TileGeom const tg = *sm.geom;
    TileCellBinner<S> const binner(g, tp, tg);
    for (uint32_t j = (uint32_t)tid; j < tg.n_occ; j += (uint32_t)nthreads)
    {
        int const cell       = sm.occ[j];
        uint32_t const start = sm.cstart[cell], count = sm.ccount[cell];
        int const bin        = binner.first_bin(cell);
        if (D::sub == 1)
        {
            sm.F[bin + 1] = count;
            continue;
        }
        uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        for (uint32_t i = 0; i < count; ++i)
        {
            uint32_t const s = binner.sub(load_pt(g.pts + start + i));
            n0 += s == 0u, n1 += s == 1u, n2 += s == 2u, n3 += s == 3u;
        }
        sm.F[bin + binner.sub_offset(0) + 1] = n0, sm.F[bin + binner.sub_offset(1) + 1] = n1;
        sm.F[bin + binner.sub_offset(2) + 1] = n2, sm.F[bin + binner.sub_offset(3) + 1] = n3;
    }
}

// phase 3 (other layouts): inclusive scan of the counts in place, after which F[i] is the start
// of bin i (F[0] = 0, F[bins] = the number of staged points).  Device: the first warp.
template <int S>
PCPX_HD void tile_phase_scan(const TileSmem& sm, int tid)
{
    constexpr int bins = TileDims<S>::bins;
#ifdef __CUDA_ARCH__
    constexpr int chunk = (bins + 31) / 32;
    if (tid >= 32)
        return;
    uint32_t s = 0;
    for (int i = tid * chunk; i < (tid + 1) * chunk && i < bins; ++i)
        s += sm.F[i + 1];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (tid >= o)
            incl += up;
    }
    uint32_t run = incl - s;
    for (int i = tid * chunk; i < (tid + 1) * chunk && i < bins; ++i)
    {
        run += sm.F[i + 1];
        sm.F[i + 1] = run;
    }
#else
    if (tid != 0)
        return;
    uint32_t run = 0;
    for (int i = 0; i < bins; ++i)
    {
        run += sm.F[i + 1];
        sm.F[i + 1] = run;
    }
#endif
}

// phase 4 (other layouts): placement, again one thread per occupied cell, its points in the
// order of the sorted array
template <int S>
PCPX_HD void tile_phase_place(const GridView& g, const TileParams& tp, const TileSmem& sm, int tid,
                              int nthreads)
{
    using D = TileDims<S>;
    TileGeom const tg = *sm.geom;
    TileCellBinner<S> const binner(g, tp, tg);
    for (uint32_t j = (uint32_t)tid; j < tg.n_occ; j += (uint32_t)nthreads)
    {
        int const cell       = sm.occ[j];
        uint32_t const start = sm.cstart[cell], count = sm.ccount[cell];
        int const bin        = binner.first_bin(cell);
        uint32_t w0 = sm.F[bin], w1 = 0, w2 = 0, w3 = 0;
        if (D::sub == 4)
            w1 = sm.F[bin + binner.sub_offset(1)], w2 = sm.F[bin + binner.sub_offset(2)],
            w3 = sm.F[bin + binner.sub_offset(3)];
        for (uint32_t i = 0; i < count; ++i)
        {
            float4 const p   = load_pt(g.pts + start + i);
            uint32_t const s = binner.sub(p);
            uint32_t const pos = s == 0u ? w0 : (s == 1u ? w1 : (s == 2u ? w2 : w3));
            w0 += s == 0u, w1 += s == 1u, w2 += s == 2u, w3 += s == 3u;
            sm.P[pos] = p;
            sm.gpos[pos] = start + i;
        }
    }
    if (tid == 0)
        for (uint32_t j = 0; j < kTilePad; ++j)
            sm.P[tg.n_points + j] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
}

// phase 5 (other layouts, after the bins are placed): which rows hold points, and the tile's
// row segments
template <int S>
PCPX_HD void tile_phase_rows(const TileSmem& sm, int tid, int nthreads)
{
    using D = TileDims<S>;
    for (int r = tid; r < D::nb * D::nc; r += nthreads)
    {
        int const ic = r / D::nb, ib = r - ic * D::nb;
        if (sm.F[(r + 1) * D::na] > sm.F[r * D::na])
        {
#ifdef __CUDA_ARCH__
            atomicOr(sm.rowmask + ic, 1u << ib);
#else
            sm.rowmask[ic] |= 1u << ib;
#endif
        }
    }
    if (tid == 0)
    {
        // the tile's own points: rows ic in [1, 1 + 4), ib in [SB, 5 SB), bins ia in [SA, 5 SA)
        uint32_t run = 0;
        int seg      = 0;
        for (int ic = 1; ic <= kTileCells; ++ic)
            for (int ib = D::SB; ib < D::SB + kTileCells * D::SB; ++ib, ++seg)
            {
                int const base  = (ic * D::nb + ib) * D::na;
                sm.seg_off[seg] = run;
                run += sm.F[base + D::SA + kTileCells * D::SA] - sm.F[base + D::SA];
            }
        sm.seg_off[seg] = run;
    }
}

// Staged position of the tile's i-th query (the tile's own points in row-major bin order):
// the segment by a branch-free binary search over the (non-decreasing) segment starts.
template <int S>
PCPX_HD uint32_t tile_query_pos(const TileSmem& sm, uint32_t i)
{
    using D = TileDims<S>;
    int seg = 0;
#pragma unroll
    for (int step = D::segs / 2; step >= 1; step >>= 1)
        seg += sm.seg_off[seg + step] <= i ? step : 0;
    int const ic = 1 + seg / (kTileCells * D::SB), ib = D::SB + seg % (kTileCells * D::SB);
    return sm.F[(ic * D::nb + ib) * D::na + D::SA] + (i - sm.seg_off[seg]);
}

// ---- the per-query search --------------------------------------------------------------------
template <int KL>
struct TileList
{
    uint32_t a[KL];

    PCPX_HD void reset()
    {
#pragma unroll
        for (int j = 0; j < KL; ++j)
            a[j] = kKeyEmpty;
    }
    PCPX_HD uint32_t get(uint32_t j) const // a[j] without dynamic register indexing
    {
        uint32_t r = a[0];
#pragma unroll
        for (int i = 1; i < KL; ++i)
            r = j == (uint32_t)i ? a[i] : r;
        return r;
    }
};

PCPX_HD int tile_zigzag(int i) { return (i & 1) ? (i + 1) >> 1 : -(i >> 1); } // 0, +1, -1, +2, -2, ...

// The search of one query is split in two:
//   Q1  tile_list_candidates: the geometry only — rows, bin ranges — and the staged positions of
//       the candidates go to a per-thread list in shared memory (a 4-instruction loop body);
//   Q2  tile_select: a REGULAR loop over that list, eight candidates at a time: keys, a 19-exchange
//       sorting network over the eight, and a bitonic merge into the sorted register list
//       (13.75 min/max per candidate for a 16-entry list instead of 24 for a sorted insert), no
//       exclusion test (the query's own staged position is left out when the list is made; a
//       foreign point inside the exclusion box shows up as a smallest key below 3 eps^2 and sends
//       the query to the retry queue).

PCPX_HD float tile_sqrt(float x) // only ever widens a bin range: an approximation is fine
{
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}

// Per-query state of the candidate listing: the listing stops when the list is full and resumes
// where it stopped, so a query's candidates can be worked off in several rounds (and the ball
// shrinks between rounds).
struct TileCursor
{
    float ta, tb, tc; // position of the query in bin / cell units, relative to the region
    float r2scan;     // squared radius of the ball that will have been scanned at the end
    float r2;         // current squared radius (shrinks once the list is full)
    int ib, ic;
    int jc, jb;          // next row to visit
    uint32_t p_resume;   // position inside that row, 0 = from its start
    bool done;
};

template <int S>
PCPX_HD TileCursor tile_cursor(const GridView& g, const TileParams& tp, const TileGeom& tg, float qx,
                               float qy, float qz)
{
    using D = TileDims<S>;
    int const A = tg.ax[0], B = tg.ax[1], C = tg.ax[2];
    TileCursor cu;
    cu.ta = (axis_of(qx, qy, qz, A) - axis_of(g.ox, g.oy, g.oz, A)) * tp.bins_per_len -
            (float)(tg.r0[0] * D::SA);
    cu.tb = (axis_of(qx, qy, qz, B) - axis_of(g.ox, g.oy, g.oz, B)) * tp.bins_per_len_b -
            (float)(tg.r0[1] * D::SB);
    cu.tc = (axis_of(qx, qy, qz, C) - axis_of(g.ox, g.oy, g.oz, C)) * tp.cells_per_len -
            (float)tg.r0[2];
    int ib = (int)floorf(cu.tb), ic = (int)floorf(cu.tc);
    cu.ib = ib < 0 ? 0 : (ib > D::nb - 1 ? D::nb - 1 : ib);
    cu.ic = ic < 0 ? 0 : (ic > D::nc - 1 ? D::nc - 1 : ic);
    // the scan ball must stay inside the staged region
    float const ga = fminf(cu.ta, (float)D::na - cu.ta) - tp.delta_bins;
    float const gb = fminf(cu.tb, (float)D::nb - cu.tb) - tp.delta_bins_b;
    float const gc = fminf(cu.tc, (float)D::nc - cu.tc) - tp.delta_cells;
    float rscan    = fminf(fminf(fminf(ga * tp.len_per_bin, gb * tp.len_per_bin_b), gc * tp.h),
                           tp.scan_cap * tp.h);
    rscan          = rscan > 0.f ? rscan : 0.f;
    cu.r2scan      = rscan * rscan * 0.999999f;
    cu.r2          = cu.r2scan;
    cu.jc = 0, cu.jb = 0, cu.p_resume = 0u;
    cu.done = false;
    return cu;
}

// Q1.  Writes the staged positions of the points in the bins the current ball touches (the
// query's own position `self` left out) to cl[j * stride], j = 0 .. count - 1, rows nearest
// first, until `cap` entries are listed or every row has been visited (cu.done).
template <int S>
PCPX_HD uint32_t tile_list_candidates(const TileParams& tp, const uint32_t* F,
                                      const uint32_t* rowmask, TileCursor& cu, uint32_t self,
                                      uint16_t* cl, int stride, uint32_t cap)
{
    using D        = TileDims<S>;
    float const r2 = cu.r2 * 1.00001f;
    uint32_t cnt   = 0;
    // where the previous round stopped: applies to the first row visited now and to no other
    uint32_t resume = cu.p_resume;
    cu.p_resume     = 0u;
    for (; cu.jc <= 2 * tp.rows_c; ++cu.jc, cu.jb = 0, resume = 0u)
    {
        int const dc = tile_zigzag(cu.jc), rc = cu.ic + dc;
        if ((unsigned)rc >= (unsigned)D::nc)
            continue;
        uint32_t const rm = rowmask[rc];
        if (rm == 0u)
            continue;
        float const gapc = dc > 0 ? (float)rc - cu.tc : (dc < 0 ? cu.tc - (float)(rc + 1) : 0.f);
        float lbc        = (gapc - tp.delta_cells) * tp.h;
        lbc              = lbc > 0.f ? lbc : 0.f;
        float const remc = r2 - lbc * lbc;
        if (remc < 0.f)
            continue;
        for (; cu.jb <= 2 * tp.rows_b; ++cu.jb)
        {
            uint32_t const pr = resume;
            resume            = 0u;
            int const db = tile_zigzag(cu.jb), rb = cu.ib + db;
            if ((unsigned)rb >= (unsigned)D::nb || !((rm >> rb) & 1u))
                continue;
            float const gapb =
                db > 0 ? (float)rb - cu.tb : (db < 0 ? cu.tb - (float)(rb + 1) : 0.f);
            float lbb       = (gapb - tp.delta_bins_b) * tp.len_per_bin_b;
            lbb             = lbb > 0.f ? lbb : 0.f;
            float const rem = remc - lbb * lbb;
            if (rem < 0.f)
                continue;
            float const reach = tile_sqrt(rem) * (tp.bins_per_len * 1.000002f) + tp.delta_bins;
            int alo = (int)floorf(cu.ta - reach), ahi = (int)floorf(cu.ta + reach);
            alo = alo < 0 ? 0 : alo;
            ahi = ahi > D::na - 1 ? D::na - 1 : ahi;
            if (alo > ahi)
                continue;
            int const base    = (rc * D::nb + rb) * D::na;
            uint32_t const hi = F[base + ahi + 1];
            uint32_t p        = F[base + alo];
            p                 = pr > p ? pr : p;
            if (p >= hi)
                continue;
            uint32_t const n = hi - p, room = cap - cnt;
            uint32_t take    = n < room ? n : room;
            for (uint32_t i = 0; i < take; ++i)
                cl[(cnt + i) * (uint32_t)stride] = (uint16_t)(p + i);
            uint32_t const next = p + take;
            if (self - p < take) // the query itself: the last entry listed now takes its slot
            {
                cl[(cnt + (self - p)) * (uint32_t)stride] = (uint16_t)(next - 1u);
                take -= 1u;
            }
            cnt += take;
            if (next < hi) // the list is full
            {
                cu.p_resume = next;
                return cnt;
            }
        }
    }
    cu.done = true;
    return cnt;
}

#define PCPX_CE(x, y)                                                                          \
    do                                                                                         \
    {                                                                                          \
        uint32_t const lo_ = (x) < (y) ? (x) : (y), hi_ = (x) < (y) ? (y) : (x);               \
        (x) = lo_, (y) = hi_;                                                                  \
    } while (0)

// 19 exchanges (optimal for 8 inputs)
PCPX_HD void tile_sort8(uint32_t* b)
{
    PCPX_CE(b[0], b[1]); PCPX_CE(b[2], b[3]); PCPX_CE(b[4], b[5]); PCPX_CE(b[6], b[7]);
    PCPX_CE(b[0], b[2]); PCPX_CE(b[1], b[3]); PCPX_CE(b[4], b[6]); PCPX_CE(b[5], b[7]);
    PCPX_CE(b[1], b[2]); PCPX_CE(b[5], b[6]); PCPX_CE(b[0], b[4]); PCPX_CE(b[3], b[7]);
    PCPX_CE(b[1], b[5]); PCPX_CE(b[2], b[6]);
    PCPX_CE(b[1], b[4]); PCPX_CE(b[3], b[6]);
    PCPX_CE(b[2], b[4]); PCPX_CE(b[3], b[5]);
    PCPX_CE(b[3], b[4]);
}

// The KL smallest of (sorted list a[KL]) + (sorted batch b[8]), sorted: the element-wise minimum
// of the ascending list and the descending (padded) batch is a bitonic sequence that holds them;
// a bitonic merge sorts it.  KL must be a power of two >= 8.
template <int KL>
PCPX_HD void tile_merge8(uint32_t* a, const uint32_t* b)
{
    static_assert((KL & (KL - 1)) == 0 && KL >= 8, "list size must be a power of two >= 8");
#pragma unroll
    for (int i = 0; i < 8; ++i)
        a[KL - 8 + i] = a[KL - 8 + i] < b[7 - i] ? a[KL - 8 + i] : b[7 - i];
#pragma unroll
    for (int s = KL / 2; s >= 1; s >>= 1)
#pragma unroll
        for (int i = 0; i < KL; ++i)
            if ((i & s) == 0)
                PCPX_CE(a[i], a[i + s]);
}

// Q2.  `top` must be reset by the caller; the list is padded to a multiple of 8 entries with the
// position of the staged array's pad entry (coordinates +inf: its key sorts behind every point).
template <int KL>
PCPX_HD void tile_select(const float4* P, const uint16_t* cl, int stride, uint32_t count,
                         float qx, float qy, float qz, uint32_t mask, TileList<KL>& top)
{
    for (uint32_t j0 = 0; j0 < count; j0 += 8)
    {
        uint32_t b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            uint32_t const pos = cl[(j0 + (uint32_t)i) * (uint32_t)stride];
            float4 const c     = P[pos];
            float const d2 = sqdist_x(fsub_x(c.x, qx), fsub_x(c.y, qy), fsub_x(c.z, qz));
            b[i]           = (f2u(d2) & ~mask) | pos;
        }
        tile_sort8(b);
        tile_merge8<KL>(top.a, b);
    }
}

// The whole search of one query: rounds of (list up to a cap, select), the ball shrinking to the
// k-th key's bucket between rounds.  The first round is short (the rows nearest to the query), so
// that the farther rows are already listed against a tight ball.
template <int KL, int S, int KS>
PCPX_HD void tile_search_batched(const TileParams& tp, const float4* P, const uint32_t* F,
                                 const uint32_t* rowmask, TileCursor& cu, uint32_t self,
                                 uint16_t* cl, int stride, float qx, float qy, float qz,
                                 uint32_t k, uint32_t first_cap, uint32_t pad_pos,
                                 TileList<KL>& top, uint32_t* n_cand)
{
    top.reset();
    uint32_t cap = first_cap, total = 0;
    do
    {
        uint32_t const cnt = tile_list_candidates<S>(tp, F, rowmask, cu, self, cl, stride, cap);
        for (uint32_t j = cnt; j < ((cnt + 7u) & ~7u); ++j)
            cl[j * (uint32_t)stride] = (uint16_t)pad_pos;
        tile_select<KL>(P, cl, stride, cnt, qx, qy, qz, tp.key_mask, top);
        // (NaN while the list is not full: fminf keeps r2)
        cu.r2 = fminf(cu.r2, u2f((KS > 0 ? top.a[KS > 0 ? KS - 1 : 0] : top.get(k - 1)) | tp.key_mask));
        cap   = (uint32_t)kTileCandCap;
        total += cnt;
    } while (!cu.done);
    if (n_cand)
        *n_cand = total;
}

// final: k eligible points found, the k-th inside the scanned ball, the (k+1)-th different from
// it in the kept bits, and no foreign point inside the exclusion box (its key would be the
// smallest; the query's own position was never listed)
template <int KL, int KS>
PCPX_HD bool tile_is_final(const TileList<KL>& top, uint32_t k, uint32_t mask, float r2scan,
                           float eps)
{
    uint32_t kk = top.a[KS > 0 ? KS - 1 : 0], kn = top.a[KS > 0 ? KS : 0];
    if (KS == 0)
    {
#pragma unroll
        for (int i = 0; i + 1 < KL; ++i)
            if (k == (uint32_t)i + 1u)
                kk = top.a[i], kn = top.a[i + 1];
    }
    bool ok = kk != kKeyEmpty && u2f(kk | mask) <= r2scan && ((kk ^ kn) & ~mask) != 0u;
    if (eps > 0.f)
        ok = ok && !(u2f(top.a[0] & ~mask) < 3.0001f * eps * eps);
    return ok;
}

// ---- epilogues over the k winners --------------------------------------------------------------

// The k winners' staged positions, out of the register list into the thread's column of `cl`, so
// that the epilogues are rolled loops instead of KL unrolled copies of their body.
template <int KL>
PCPX_HD void tile_store_winners(const TileList<KL>& top, uint32_t mask, uint16_t* cl, int stride)
{
#pragma unroll
    for (int j = 0; j < KL; ++j)
        cl[j * stride] = (uint16_t)(top.a[j] & mask);
}

PCPX_HD void tile_normal_rolled(const float4* P, const uint16_t* cl, int stride, uint32_t k,
                                float qx, float qy, float qz, float* n3, float* c3)
{
    float s1x = 0.f, s1y = 0.f, s1z = 0.f;
    Sym3 s2{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (uint32_t j = 0; j < k; ++j)
    {
        float4 const c = P[cl[j * (uint32_t)stride]];
        float const dx = c.x - qx, dy = c.y - qy, dz = c.z - qz;
        s1x += dx, s1y += dy, s1z += dz;
        s2.xx += dx * dx, s2.xy += dx * dy, s2.xz += dx * dz;
        s2.yy += dy * dy, s2.yz += dy * dz, s2.zz += dz * dz;
    }
    float const inv = 1.f / (float)k;
    float const mx = s1x * inv, my = s1y * inv, mz = s1z * inv;
    Sym3 m;
    m.xx = s2.xx - s1x * mx, m.xy = s2.xy - s1x * my, m.xz = s2.xz - s1x * mz;
    m.yy = s2.yy - s1y * my, m.yz = s2.yz - s1y * mz, m.zz = s2.zz - s1z * mz;
    smallest_eigenvector_fast(m, n3[0], n3[1], n3[2]);
    c3[0] = qx + mx, c3[1] = qy + my, c3[2] = qz + mz;
}

// tile_emit_sorted over the stored positions
template <class F>
PCPX_HD bool tile_emit_sorted_rolled(const float4* P, const uint16_t* cl, int stride, uint32_t k,
                                     float qx, float qy, float qz, F&& f)
{
    float hd = 0.f, ld = -1.f; // held back / last emitted
    uint32_t hid = 0, lid = 0, slot = 0;
    bool ok = true;
#pragma unroll 1
    for (uint32_t j = 0; j < k; ++j)
    {
        float4 const c    = P[cl[j * (uint32_t)stride]];
        float const d2    = sqdist_x(fsub_x(c.x, qx), fsub_x(c.y, qy), fsub_x(c.z, qz));
        uint32_t const id = f2u(c.w);
        if (j == 0)
        {
            hd = d2, hid = id;
            continue;
        }
        bool const before_held = d2 < hd || (d2 == hd && id < hid);
        float const ed         = before_held ? d2 : hd;
        uint32_t const eid     = before_held ? id : hid;
        ok = ok && !(ed < ld || (ed == ld && eid < lid && slot > 0));
        f(slot++, ed, eid);
        ld = ed, lid = eid;
        if (!before_held)
            hd = d2, hid = id;
    }
    ok = ok && !(hd < ld || (hd == ld && hid < lid && slot > 0));
    f(slot, hd, hid);
    return ok;
}

// Puts the k winners' positions (cl[0 .. k)) into exact ascending (d2, original index) order in
// place — same one-displacement rule as tile_emit_sorted — so that row outputs can be produced
// slot by slot afterwards.  Returns false when the order could not be established.
PCPX_HD bool tile_order_winners(const float4* P, uint16_t* cl, int stride, uint32_t k, float qx,
                                float qy, float qz)
{
    float hd = 0.f, ld = -1.f;
    uint32_t hid = 0, lid = 0, hpos = 0, slot = 0;
    bool ok = true;
#pragma unroll 1
    for (uint32_t j = 0; j < k; ++j)
    {
        uint32_t const pos = cl[j * (uint32_t)stride];
        float4 const c     = P[pos];
        float const d2     = sqdist_x(fsub_x(c.x, qx), fsub_x(c.y, qy), fsub_x(c.z, qz));
        uint32_t const id  = f2u(c.w);
        if (j == 0)
        {
            hd = d2, hid = id, hpos = pos;
            continue;
        }
        bool const before_held = d2 < hd || (d2 == hd && id < hid);
        float const ed         = before_held ? d2 : hd;
        uint32_t const eid     = before_held ? id : hid;
        ok = ok && !(ed < ld || (ed == ld && eid < lid && slot > 0));
        cl[slot * (uint32_t)stride] = (uint16_t)(before_held ? pos : hpos); // slot < j: already read
        ++slot;
        ld = ed, lid = eid;
        if (!before_held)
            hd = d2, hid = id, hpos = pos;
    }
    ok = ok && !(hd < ld || (hd == ld && hid < lid && slot > 0));
    cl[slot * (uint32_t)stride] = (uint16_t)hpos;
    return ok;
}

} // namespace pcpx
