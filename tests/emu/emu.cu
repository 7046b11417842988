// TEST HARNESS ONLY — not part of libpcpx.so and never a product path.
//
// Runs the SAME __host__ __device__ traversal code the CUDA kernels call (grid_core.cuh,
// knn_core.cuh, radius_core.cuh, normals_core.cuh, eig3.cuh) on the CPU, over an index built by
// a small host re-statement of the device build, so that `-m "not gpu"` tests can compare the
// search logic (pruning bounds, level walk, tie handling, exclusion box) with the oracle in a
// container that has no GPU.  The device build itself (radix sort, hash insertion kernels,
// scans) is only exercised by the `-m gpu` tests.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "big_k.cuh"
#include "plan.hpp"
#include "normals_core.cuh"
#include "radius_core.cuh"
#include "smoothing_core.cuh"
#include "tile_core.cuh"
#include "tree_core.cuh"

using namespace pcpx;

struct EmuIndex
{
    GridView g{};
    std::vector<float4> pts;
    std::vector<HashSlot> table;
    uint64_t n_input = 0;
    uint64_t cells_per_level[kMaxLevel + 2] = {};
};

// mirrors plan_for() in query_body.inc
static SearchPlan plan_for(const EmuIndex* ix, uint32_t k, double margin)
{
    PlanChoice const c = choose_plan(ix->g.n, ix->cells_per_level, ix->g.lfine, k, margin);
    return SearchPlan{c.level, c.rings};
}

static uint32_t host_claim(std::vector<HashSlot>& t, uint64_t key)
{
    uint32_t s = hash_slot(key, (uint32_t)t.size());
    for (;;)
    {
        if (t[s].key_hi == kEmptyKeyHi)
        {
            t[s].key_lo = (uint32_t)key, t[s].key_hi = (uint32_t)(key >> 32);
            return s;
        }
        if (t[s].key_lo == (uint32_t)key && t[s].key_hi == (uint32_t)(key >> 32))
            return s;
        s = s + 1 == t.size() ? 0u : s + 1;
    }
}

extern "C" {

void* emu_index_create(const float* xyz, size_t n, int use_box, const float* box6,
                       uint32_t max_level, uint32_t min_occ_in)
{
    EmuIndex* ix = new EmuIndex();
    ix->n_input  = n;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    float maxabs = 0.f;
    std::vector<uint8_t> inside(n, 1);
    size_t n_inside = 0;
    for (size_t i = 0; i < n; ++i)
    {
        for (int a = 0; a < 3; ++a)
        {
            float v = xyz[3 * i + a];
            mn[a] = std::min(mn[a], v), mx[a] = std::max(mx[a], v);
        }
        if (use_box)
            for (int a = 0; a < 3; ++a)
                inside[i] &= xyz[3 * i + a] >= box6[a] && xyz[3 * i + a] <= box6[3 + a];
        n_inside += inside[i];
    }
    if (use_box)
        for (int a = 0; a < 3; ++a)
            mn[a] = box6[a], mx[a] = box6[3 + a];
    if (n == 0)
        for (int a = 0; a < 3; ++a)
            mn[a] = mx[a] = 0.f;
    float extent = 0.f;
    for (int a = 0; a < 3; ++a)
    {
        extent = std::max(extent, mx[a] - mn[a]);
        maxabs = std::max(maxabs, std::max(std::fabs(mn[a]), std::fabs(mx[a])));
    }
    if (!(extent > 0.f) || !std::isfinite(extent))
        extent = maxabs > 0.f && std::isfinite(maxabs) ? maxabs * 1e-3f : 1.f;
    extent *= 1.0001f;
    GridView& g = ix->g;
    g.ox = mn[0], g.oy = mn[1], g.oz = mn[2];
    g.extent = extent;
    int lcap = (int)std::ceil(std::log2((double)std::max<size_t>(n_inside, 2)) / 2.0) + 1;
    if (max_level)
        lcap = (int)max_level;
    double const ulp = std::max((double)maxabs, (double)extent) * std::ldexp(1.0, -23);
    lcap = std::min(lcap, (int)std::floor(std::log2((double)extent / (64.0 * ulp))));
    lcap = std::max(1, std::min(lcap, kMaxLevel));
    g.lcap  = lcap;
    g.scale = std::ldexp(1.f, lcap) / extent;
    g.delta = 16.f * std::ldexp(std::max(extent, maxabs), -23);
    g.n     = (uint32_t)n_inside;

    std::vector<uint64_t> code(n);
    for (size_t i = 0; i < n; ++i)
    {
        QueryCell c = query_cell(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        code[i]     = inside[i] ? morton3(c.ux, c.uy, c.uz) : ~0ull;
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    ix->pts.assign(n + kPtsPad, make_float4(0.f, 0.f, 0.f, 0.f)); // padded like the device array
    for (size_t i = 0; i < n; ++i)
    {
        uint32_t o = order[i];
        ix->pts[i] = make_float4(xyz[3 * o], xyz[3 * o + 1], xyz[3 * o + 2], u2f(o));
    }
    g.pts = ix->pts.data();

    auto boundary = [&](uint32_t i) -> int {
        if (i == 0)
            return 0;
        QueryCell a = query_cell(g, ix->pts[i - 1].x, ix->pts[i - 1].y, ix->pts[i - 1].z);
        QueryCell b = query_cell(g, ix->pts[i].x, ix->pts[i].y, ix->pts[i].z);
        uint32_t diff = (a.ux ^ b.ux) | (a.uy ^ b.uy) | (a.uz ^ b.uz);
        if (!diff)
            return g.lcap + 1;
        int hb = 31 - __builtin_clz(diff);
        return g.lcap - hb;
    };
    std::vector<uint64_t> lh(kMaxLevel + 2, 0);
    for (uint32_t i = 0; i < g.n; ++i)
        lh[boundary(i)]++;
    double min_occ = min_occ_in ? (double)min_occ_in : 4.0;
    uint64_t cells = 0, total = 0;
    g.lfine = 0;
    for (int l = 0; l <= g.lcap; ++l)
    {
        cells += lh[l];
        if (l > 0 && (double)g.n / (double)std::max<uint64_t>(cells, 1) < min_occ)
            break;
        g.lfine = l;
        ix->cells_per_level[l] = cells;
        total += cells;
    }
    size_t slots = std::max<uint64_t>(64, (uint64_t)((double)total * 2.5) + 1);
    ix->table.assign(slots, HashSlot{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu});
    g.table_size = (uint32_t)slots;
    for (uint32_t i = 0; i < g.n; ++i)
    {
        int b = boundary(i);
        QueryCell c = query_cell(g, ix->pts[i].x, ix->pts[i].y, ix->pts[i].z);
        for (int l = b; l <= g.lfine; ++l)
        {
            int sh     = g.lcap - l;
            uint32_t s = host_claim(ix->table, cell_key(l, c.ux >> sh, c.uy >> sh, c.uz >> sh));
            ix->table[s].start = i;
        }
    }
    for (uint32_t i = 1; i <= g.n; ++i)
    {
        int b = i == g.n ? 0 : boundary(i);
        QueryCell c = query_cell(g, ix->pts[i - 1].x, ix->pts[i - 1].y, ix->pts[i - 1].z);
        for (int l = b; l <= g.lfine; ++l)
        {
            int sh     = g.lcap - l;
            uint32_t s = host_claim(ix->table, cell_key(l, c.ux >> sh, c.uy >> sh, c.uz >> sh));
            ix->table[s].count = i - ix->table[s].start;
        }
    }
    g.table = ix->table.data();
    return ix;
}

void emu_index_destroy(void* h) { delete static_cast<EmuIndex*>(h); }

void emu_plan(void* h, uint32_t k, double margin, int* level, int* rings, double* occupancy)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    PlanChoice const c =
        choose_plan(ix->g.n, ix->cells_per_level, ix->g.lfine, k, margin);
    *level = c.level, *rings = c.rings;
    *occupancy = ix->cells_per_level[c.level]
                     ? (double)ix->g.n / (double)ix->cells_per_level[c.level]
                     : 0.0;
}

// original index of every sorted position (the processing order of self-queries)
void emu_sorted_order(void* h, uint32_t* out)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    for (uint64_t i = 0; i < ix->g.n; ++i)
        out[i] = f2u(ix->g.pts[i].w);
}

void emu_index_info(void* h, uint64_t* n_indexed, int* lcap, int* lfine, uint64_t* slots)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    *n_indexed = ix->g.n, *lcap = ix->g.lcap, *lfine = ix->g.lfine, *slots = ix->table.size();
}

} // extern "C"

static void fetch(EmuIndex* ix, const float* q, size_t i, float& x, float& y, float& z,
                  size_t& row)
{
    if (!q)
    {
        float4 c = ix->pts[i];
        x = c.x, y = c.y, z = c.z, row = f2u(c.w);
    }
    else
        x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2], row = i;
}

static int list_size_for(uint32_t k)
{
    static const int sizes[] = {4, 8, 10, 12, 15, 16, 20, 24, 28, 30, 32};
    for (int s : sizes)
        if ((uint32_t)s >= k)
            return s;
    return 0;
}
constexpr int exact_k(int K) { return (K + 3) / 4 * 4; }

// mirrors knn_main_kernel / knn_retry_kernel in query_body.inc; mode 0 = two-pass with fallback (the
// product path), 1 = force the exact 64-bit search for every query
template <int K>
static void knn_impl(EmuIndex* ix, const float* q, size_t nq, uint32_t k, float eps,
                     SearchPlan plan, int mode, uint32_t* idx, float* d2, uint32_t* cnt,
                     uint64_t* st4, uint32_t* per_query_cand)
{
    for (size_t i = 0; i < nq; ++i)
    {
        float x, y, z;
        size_t row;
        fetch(ix, q, i, x, y, z, row);
        SearchStats st;
        bool done = false;
        if (mode == 0)
        {
            // main pass: one attempt of the plan; otherwise the retry pass walks coarser
            auto run = [&](auto rings_tag) {
                constexpr int R = decltype(rings_tag)::value;
                TopD<K> top;
                BlockGeom b;
                CellList cl;
                ShortListFor<K> sl;
                uint32_t* irow = idx + row * k;
                float* drow    = d2 ? d2 + row * k : nullptr;
                uint32_t* crow = cnt ? cnt + row : nullptr;
                if (knn_attempt_dist<K, R>(ix->g, query_cell(ix->g, x, y, z), plan.level, x, y, z,
                                           k, eps, top, b, cl, sl, &st))
                    return knn_two_pass_emit<K>(ix->g, BlockRegion<R>{b, plan.level}, sl, x, y, z,
                                                top, k, eps, irow, drow, crow);
                // the retry pass: one level coarser, then the octree descent
                if (plan.level > 0 &&
                    knn_attempt_dist<K, R>(ix->g, query_cell(ix->g, x, y, z), plan.level - 1, x, y,
                                           z, k, eps, top, b, cl, sl, &st))
                    return knn_two_pass_emit<K>(ix->g, BlockRegion<R>{b, plan.level - 1}, sl, x, y,
                                                z, top, k, eps, irow, drow, crow);
                knn_tree_dist<K>(ix->g, x, y, z, eps, top, sl, &st);
                return knn_two_pass_emit<K>(ix->g, TreeRegion{}, sl, x, y, z, top, k, eps, irow,
                                            drow, crow);
            };
            done = plan.rings >= 2 ? run(std::integral_constant<int, 2>{})
                                   : run(std::integral_constant<int, 1>{});
            if (!done && st4)
                st4[3] += 1; // retries
        }
        if (!done)
        {
            TopK<exact_k(K)> top;
            knn_search<exact_k(K), TIE_ORIGINAL_INDEX>(ix->g, x, y, z, k, eps, plan.level, top,
                                                       mode == 1 ? &st : nullptr);
            uint32_t n = 0;
            for (uint32_t j = 0; j < k; ++j)
            {
                bool valid       = top.a[j] != kEmptyEntry;
                idx[row * k + j] = valid ? (uint32_t)top.a[j] : 0xFFFFFFFFu;
                if (d2)
                    d2[row * k + j] = valid ? u2f((uint32_t)(top.a[j] >> 32)) : INFINITY;
                n += valid;
            }
            if (cnt)
                cnt[row] = n;
        }
        if (st4)
            st4[0] += st.candidates, st4[1] += st.lookups, st4[2] += st.attempts;
        if (per_query_cand)
            per_query_cand[i] = st.candidates; // processing order (Morton order for self-queries)
    }
}

#define EMU_DISPATCH(KR, CALL)                                                                 \
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
    default: return -5;                                                                        \
    }

extern "C" int emu_knn(void* h, const float* q, size_t nq, uint32_t k, double eps,
                       double level_factor, int mode, uint32_t* idx, float* d2, uint32_t* cnt,
                       uint64_t* st4, uint32_t* per_query_cand)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    if (k == 0 || nq == 0)
        return 0;
    SearchPlan const plan = plan_for(ix, k, level_factor);
    int const level       = plan.level;
    if (k > 32) // mirrors knn_big_kernel
    {
        if (k > kBigKMax)
            return -5;
        std::vector<uint64_t> keys(k);
        for (size_t i = 0; i < nq; ++i)
        {
            float x, y, z;
            size_t row;
            fetch(ix, q, i, x, y, z, row);
            BigHeap heap{keys.data(), 0u, k};
            knn_search_big(ix->g, x, y, z, (float)eps, level, heap);
            for (uint32_t j = 0; j < k; ++j)
            {
                bool valid       = j < heap.n;
                idx[row * k + j] = valid ? (uint32_t)keys[j] : 0xFFFFFFFFu;
                if (d2)
                    d2[row * k + j] = valid ? u2f((uint32_t)(keys[j] >> 32)) : INFINITY;
            }
            if (cnt)
                cnt[row] = heap.n;
        }
        return 0;
    }
    EMU_DISPATCH(list_size_for(k),
                 (knn_impl<KK>(ix, q, nq, k, (float)eps, plan, mode, idx, d2, cnt, st4,
                               per_query_cand)));
    return 0;
}

// mirrors normals_kernel / normal_exact and mean_distance_kernel
template <int K>
static void normals_impl(EmuIndex* ix, const float* q, size_t nq, uint32_t k, float eps,
                         SearchPlan plan, int mode, float* ctr, float* nrm, float* means,
                         uint32_t* ties)
{
    for (size_t i = 0; i < nq; ++i)
    {
        float x, y, z;
        size_t row;
        fetch(ix, q, i, x, y, z, row);
        float n3[3], c3[3];
        TopD<K> top;
        auto run = [&](auto rings_tag) {
            constexpr int R = decltype(rings_tag)::value;
            BlockGeom b;
            CellList cl;
            ShortListFor<K> sl;
            if (knn_attempt_dist<K, R>(ix->g, query_cell(ix->g, x, y, z), plan.level, x, y, z, k,
                                       eps, top, b, cl, sl, nullptr))
                return mode == 0 && normal_two_pass<K>(ix->g, BlockRegion<R>{b, plan.level}, sl, x,
                                                       y, z, top, k, eps, n3, c3, nullptr);
            if (plan.level > 0 &&
                knn_attempt_dist<K, R>(ix->g, query_cell(ix->g, x, y, z), plan.level - 1, x, y, z,
                                       k, eps, top, b, cl, sl, nullptr))
                return mode == 0 && normal_two_pass<K>(ix->g, BlockRegion<R>{b, plan.level - 1}, sl,
                                                       x, y, z, top, k, eps, n3, c3, nullptr);
            knn_tree_dist<K>(ix->g, x, y, z, eps, top, sl, nullptr);
            return mode == 0 &&
                   normal_two_pass<K>(ix->g, TreeRegion{}, sl, x, y, z, top, k, eps, n3, c3, nullptr);
        };
        bool ok = plan.rings >= 2 ? run(std::integral_constant<int, 2>{})
                                  : run(std::integral_constant<int, 1>{});
        if (!ok)
        {
            TopK<exact_k(K)> ids;
            int lv = knn_search<exact_k(K), TIE_ORIGINAL_INDEX>(ix->g, x, y, z, k, eps, plan.level,
                                                                ids, nullptr);
            normal_from_ids(ix->g, query_cell(ix->g, x, y, z), lv, ids, k, n3, c3, nullptr);
            if (ties)
                ++*ties;
        }
        for (int a = 0; a < 3; ++a)
        {
            nrm[3 * row + a] = n3[a];
            if (ctr)
                ctr[3 * row + a] = c3[a];
        }
        if (means)
            means[row] = mean_distance_d(top, k);
    }
}

extern "C" int emu_normals(void* h, const float* q, size_t nq, uint32_t k, double eps,
                           double level_factor, int mode, float* ctr, float* nrm, float* means,
                           uint32_t* ties)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    if (k == 0 || nq == 0)
        return 0;
    SearchPlan const plan = plan_for(ix, k, level_factor);
    EMU_DISPATCH(list_size_for(k),
                 (normals_impl<KK>(ix, q, nq, k, (float)eps, plan, mode, ctr, nrm, means, ties)));
    return 0;
}

extern "C" {

int emu_radius(void* h, const float* q, size_t nq, const float* radii, float r, uint32_t* cnt,
               const uint64_t* offsets, uint32_t* out_idx)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    for (size_t i = 0; i < nq; ++i)
    {
        float x, y, z;
        size_t row;
        fetch(ix, q, i, x, y, z, row);
        uint32_t c = 0;
        uint64_t w = offsets ? offsets[row] : 0;
        radius_visit(ix->g, x, y, z, radii ? radii[row] : r, [&](float4 const& p, uint32_t) {
            if (out_idx)
                out_idx[w++] = f2u(p.w);
            ++c;
            return false;
        });
        if (cnt)
            cnt[row] = c;
    }
    return 0;
}

int emu_density_keep(void* h, float r, uint32_t threshold, uint8_t* keep)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    for (size_t i = 0; i < ix->n_input; ++i)
    {
        float4 q = ix->pts[i];
        uint32_t c = 0;
        if (threshold)
            radius_visit_lazy(ix->g, q.x, q.y, q.z, r,
                              [&](float4 const&, uint32_t) { return ++c >= threshold; });
        keep[f2u(q.w)] = c >= threshold;
    }
    return 0;
}

void emu_smallest_eigenvector_fast(const float* cov6, size_t n, float* n3)
{
    for (size_t i = 0; i < n; ++i)
    {
        const float* c = cov6 + 6 * i;
        Sym3 m{c[0], c[1], c[2], c[3], c[4], c[5]};
        smallest_eigenvector_fast(m, n3[3 * i], n3[3 * i + 1], n3[3 * i + 2]);
    }
}

void emu_smallest_eigenvector(const float* cov6, float* n3, float* gap)
{
    Sym3 m{cov6[0], cov6[1], cov6[2], cov6[3], cov6[4], cov6[5]};
    smallest_eigenvector(m, n3[0], n3[1], n3[2], gap);
}

// ---- radius-search callers (smoothing_core.cuh); one step each, mirroring the kernels of
// smoothing.cu.  Attribute arrays come in ORIGINAL order and are gathered into the index's
// order exactly as gather3_kernel does.
static std::vector<float4> gather3(EmuIndex* ix, const float* in)
{
    std::vector<float4> out(ix->n_input);
    for (size_t t = 0; t < ix->n_input; ++t)
    {
        const float* r = in + 3 * (size_t)f2u(ix->pts[t].w);
        out[t]         = make_float4(r[0], r[1], r[2], 0.f);
    }
    return out;
}

void emu_bilateral_points_step(void* h, const float* normals, float sigmaf, float sigmag,
                               float* out_xyz)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    auto nrm     = gather3(ix, normals);
    for (size_t t = 0; t < ix->n_input; ++t)
    {
        float4 const s = ix->pts[t];
        bilateral_point(ix->g, nrm.data(), s.x, s.y, s.z, sigmaf, sigmag,
                        out_xyz + 3 * (size_t)f2u(s.w));
    }
}

void emu_bilateral_normals_step(void* h, const float* normals, float sigmaf, float sigmag,
                                float* out_normals)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    auto nrm     = gather3(ix, normals);
    for (size_t t = 0; t < ix->n_input; ++t)
    {
        float4 const s = ix->pts[t];
        float4 const n = nrm[t];
        bilateral_normal(ix->g, nrm.data(), s.x, s.y, s.z, n.x, n.y, n.z, sigmaf, sigmag,
                         out_normals + 3 * (size_t)f2u(s.w));
    }
}

// densities in the index's SORTED order, like wlop_density_kernel
void emu_wlop_density(void* h, float hh, float mu, float* out_sorted)
{
    EmuIndex* ix       = static_cast<EmuIndex*>(h);
    WlopParams const w = wlop_params(hh, mu);
    for (size_t t = 0; t < ix->n_input; ++t)
        out_sorted[t] = wlop_density(ix->g, w, ix->pts[t].x, ix->pts[t].y, ix->pts[t].z);
}

void emu_wlop_step(void* hp, const float* vj_sorted, void* hq, const float* wi_sorted, float hh,
                   float mu, float* out_xyz)
{
    EmuIndex* ip       = static_cast<EmuIndex*>(hp);
    EmuIndex* iq       = static_cast<EmuIndex*>(hq);
    WlopParams const w = wlop_params(hh, mu);
    for (size_t t = 0; t < iq->n_input; ++t)
    {
        float4 const q = iq->pts[t];
        wlop_step(ip->g, vj_sorted, iq->g, wi_sorted, w, q.x, q.y, q.z,
                  out_xyz + 3 * (size_t)f2u(q.w));
    }
}

} // extern "C"


// ---- the tile path (tile_core.cuh): the phases of tile_knn_kernel run sequentially over the
// threads of one emulated CTA per tile.  mode: 0 = kNN rows, 1 = mean distance, 2 = normals.
// `done[row]` = 1 when the tile pass gave the final answer (the product sends the rest to the
// retry queue).  stats: [0] tiles, [1] fallback tiles, [2] queries not final, [3] candidates,
// [4] largest staged region, [5] queries not final because the emitted order was ambiguous.
static uint32_t g_emu_first_cap = 32;
extern "C" void emu_tile_first_cap(uint32_t v) { g_emu_first_cap = v; }

template <int KL, int S>
static void tile_impl(EmuIndex* ix, uint32_t k, float eps, int mode, int level, uint32_t max_points,
                      float scan_cap, int nthreads, uint32_t* idx, float* d2, uint32_t* cnt,
                      float* nrm, float* ctr, float* means, uint8_t* done, uint64_t* stats)
{
    using D             = TileDims<S>;
    GridView const& g   = ix->g;
    TileParams tp = make_tile_params<S>(g, level, max_points, scan_cap);
    tp.first_cap  = std::min<uint32_t>(g_emu_first_cap, (uint32_t)kTileCandCap);
    std::vector<float4> P(max_points + kTilePad);
    tp.threads    = nthreads;
    std::vector<uint32_t> F(D::bins + 1), cstart(kRegionCellCount), ccount(kRegionCellCount);
    std::vector<uint8_t> occ(kRegionCellCount);
    TileGeom geom;
    TileSmem sm{};
    sm.P = P.data(), sm.F = F.data(), sm.cstart = cstart.data(), sm.ccount = ccount.data();
    sm.occ = occ.data(), sm.geom = &geom;
    constexpr int segs = D::segs;
    std::vector<uint32_t> rowmask(D::nc), seg_off(segs + 1), gpos(max_points);
    std::vector<uint16_t> clist((size_t)kTileCandCap * nthreads);
    sm.rowmask = rowmask.data(), sm.seg_off = seg_off.data(), sm.gpos = gpos.data();
    sm.cl = clist.data();
    int const sh = g.lcap - (level - kTileShift);
    auto tile_of = [&](uint32_t i) {
        QueryCell c = query_cell(g, ix->pts[i].x, ix->pts[i].y, ix->pts[i].z);
        return morton3(c.ux >> sh, c.uy >> sh, c.uz >> sh);
    };
    uint32_t s0 = 0;
    while (s0 < g.n)
    {
        uint32_t s1 = s0 + 1;
        uint64_t const code = tile_of(s0);
        while (s1 < g.n && tile_of(s1) == code)
            ++s1;
        stats[0]++;
        QueryCell const fc = query_cell(g, ix->pts[s0].x, ix->pts[s0].y, ix->pts[s0].z);
        uint64_t const txyz = tile_pack(fc.ux >> sh, fc.uy >> sh, fc.uz >> sh);
        // (stale contents must not matter)
        std::fill(F.begin(), F.end(), 0xDEADu);
        std::fill(rowmask.begin(), rowmask.end(), D::one_pass ? 0xFFFFFFFFu : 0u);
        for (int t = 0; t < nthreads; ++t)
            tile_phase_lookup<S>(g, tp, sm, txyz, t, nthreads);
        for (int t = 0; t < nthreads; ++t)
            tile_phase_plan<S>(tp, sm, txyz, t);
        stats[4] = std::max<uint64_t>(stats[4], geom.n_points);
        if (geom.fallback)
        {
            stats[1]++;
            stats[2] += s1 - s0;
            s0 = s1;
            continue;
        }
        // (thread order scrambled: nothing may depend on it)
        if constexpr (D::one_pass)
        {
            for (int t = nthreads - 1; t >= 0; --t)
                tile_phase_place_cells<S>(g, tp, sm, (t * 37) % nthreads, nthreads);
        }
        else
        {
            for (int t = nthreads - 1; t >= 0; --t)
                tile_phase_count<S>(g, tp, sm, (t * 37) % nthreads, nthreads);
            for (int t = 0; t < nthreads; ++t)
                tile_phase_scan<S>(sm, t);
            for (int t = nthreads - 1; t >= 0; --t)
                tile_phase_place<S>(g, tp, sm, (t * 37) % nthreads, nthreads);
            for (int t = 0; t < nthreads; ++t)
                tile_phase_rows<S>(sm, t, nthreads);
        }
        uint32_t nq = s1 - s0;
        if (sm.seg_off[segs] != nq || sm.F[D::bins] != geom.n_points)
            stats[1] += 1u << 20; // inconsistent query list: shows up as an absurd count
        for (int b = 0; b < D::bins; ++b)
            if (sm.F[b] > sm.F[b + 1])
                stats[1] += 1u << 20;
        for (uint32_t i = 0; i < nq; ++i)
        {
            uint32_t const pos = tile_query_pos<S>(sm, i);
            float4 const q     = sm.P[pos];
            uint32_t const row = f2u(q.w);
            if (sm.gpos[pos] < s0 || sm.gpos[pos] >= s1 || f2u(ix->pts[sm.gpos[pos]].w) != row)
                stats[1] += 1u << 20;
            TileList<KL> top;
            uint32_t cand = 0;
            int const tid = (int)(i % (uint32_t)nthreads);
            TileCursor cu = tile_cursor<S>(g, tp, geom, q.x, q.y, q.z);
            tile_search_batched<KL, S, 0>(tp, sm.P, sm.F, sm.rowmask, cu, eps > 0.f ? pos : kNoSelf,
                                          sm.cl + tid, nthreads, q.x, q.y, q.z, k, tp.first_cap,
                                          geom.n_points, top, &cand);
            bool ok = tile_is_final<KL, 0>(top, k, tp.key_mask, cu.r2scan, eps);
            if (ok)
                tile_store_winners<KL>(top, tp.key_mask, sm.cl + tid, nthreads);
            stats[3] += cand;
            auto emit = [&](auto&& f) {
                return tile_emit_sorted_rolled(sm.P, sm.cl + tid, nthreads, k, q.x, q.y, q.z, f);
            };
            if (ok && mode == 2)
            {
                float n3[3], c3[3];
                tile_normal_rolled(sm.P, sm.cl + tid, nthreads, k, q.x, q.y, q.z, n3, c3);
                for (int a = 0; a < 3; ++a)
                {
                    nrm[3 * (size_t)row + a] = n3[a];
                    if (ctr)
                        ctr[3 * (size_t)row + a] = c3[a];
                }
            }
            else if (ok && mode == 0 && k <= 16u)
            {
                // the device's row output: winners ordered in place, rows read off the positions
                ok = tile_order_winners(sm.P, sm.cl + tid, nthreads, k, q.x, q.y, q.z);
                for (uint32_t j = 0; ok && j < k; ++j)
                {
                    float4 const c = sm.P[sm.cl[tid + j * (uint32_t)nthreads]];
                    idx[(size_t)row * k + j] = f2u(c.w);
                    if (d2)
                        d2[(size_t)row * k + j] =
                            sqdist_x(fsub_x(c.x, q.x), fsub_x(c.y, q.y), fsub_x(c.z, q.z));
                }
                if (ok && cnt)
                    cnt[row] = k;
                if (!ok)
                    stats[5]++;
            }
            else if (ok && mode == 0)
            {
                ok = emit([&](uint32_t slot, float dd, uint32_t id) {
                    idx[(size_t)row * k + slot] = id;
                    if (d2)
                        d2[(size_t)row * k + slot] = dd;
                });
                if (ok && cnt)
                    cnt[row] = k;
                if (!ok)
                    stats[5]++;
            }
            else if (ok)
            {
                float sum = 0.f;
                ok = emit([&](uint32_t, float dd, uint32_t) { sum = sum + sqrtf(dd); });
                means[row] = sum / (float)k;
                if (!ok)
                    stats[5]++;
            }
            done[row] = ok;
            if (!ok)
                stats[2]++;
        }
        s0 = s1;
    }
}

extern "C" int emu_tile(void* h, uint32_t k, double eps, int mode, int sub, int alg, int level,
                        uint32_t max_points, double scan_cap, int nthreads, uint32_t* idx,
                        float* d2, uint32_t* cnt, float* nrm, float* ctr, float* means,
                        uint8_t* done, uint64_t* stats)
{
    EmuIndex* ix = static_cast<EmuIndex*>(h);
    (void)alg;
    if (level < kTileShift || level > ix->g.lfine)
        return -2;
#define EMU_TILE(KLV)                                                                          \
    case KLV:                                                                                  \
        if (sub == 1)                                                                          \
            tile_impl<KLV, 1>(ix, k, (float)eps, mode, level, max_points, (float)scan_cap,     \
                              nthreads, idx, d2, cnt, nrm, ctr, means, done, stats);           \
        else if (sub == 2)                                                                     \
            tile_impl<KLV, 2>(ix, k, (float)eps, mode, level, max_points, (float)scan_cap,     \
                              nthreads, idx, d2, cnt, nrm, ctr, means, done, stats);           \
        else                                                                                   \
            tile_impl<KLV, 4>(ix, k, (float)eps, mode, level, max_points, (float)scan_cap,     \
                              nthreads, idx, d2, cnt, nrm, ctr, means, done, stats);           \
        break;
    switch (k + 1 <= 8 ? 8 : (k + 1 <= 16 ? 16 : (k + 1 <= 32 ? 32 : 0)))
    {
        EMU_TILE(8)
        EMU_TILE(16)
        EMU_TILE(32)
    default: return -5;
    }
#undef EMU_TILE
    return 0;
}
