// kNN for k beyond the register-resident list (32 < k <= kBigKMax): one thread per query, an
// exact max-heap of 64-bit (bits(d2) << 32 | original index) keys in local memory, one pass, the
// same block walk / conservative bounds / level fallback as the fast path.  Not tuned — the
// reference's own uses are k <= 15 — but exact, so the API has no cliff at k = 33.
#pragma once
#include "knn_core.cuh"

namespace pcpx {

constexpr uint32_t kBigKMax = 256;

struct BigHeap
{
    uint64_t* a; // capacity >= k
    uint32_t n, k;

    PCPX_HD void reset() { n = 0; }
    PCPX_HD bool full() const { return n == k; }
    // fp32 view of the largest kept distance; NaN while not full ("cannot prune / not done")
    PCPX_HD float worst_d2() const { return full() ? u2f((uint32_t)(a[0] >> 32)) : u2f(0x7FC00000u); }
    PCPX_HD void sift_down(uint32_t i, uint32_t len)
    {
        uint64_t const v = a[i];
        for (;;)
        {
            uint32_t c = 2 * i + 1;
            if (c >= len)
                break;
            if (c + 1 < len && a[c + 1] > a[c])
                ++c;
            if (a[c] <= v)
                break;
            a[i] = a[c];
            i    = c;
        }
        a[i] = v;
    }
    PCPX_HD void offer(uint64_t key)
    {
        if (n < k)
        {
            uint32_t i = n++;
            while (i > 0)
            {
                uint32_t const p = (i - 1) / 2;
                if (a[p] >= key)
                    break;
                a[i] = a[p];
                i    = p;
            }
            a[i] = key;
        }
        else if (key < a[0])
        {
            a[0] = key;
            sift_down(0, n);
        }
    }
    // ascending order in place
    PCPX_HD void sort()
    {
        for (uint32_t len = n; len > 1; --len)
        {
            uint64_t const t = a[0];
            a[0]             = a[len - 1];
            a[len - 1]       = t;
            sift_down(0, len - 1);
        }
    }
};

// On return heap.a[0 .. heap.n) holds the answer ascending by (d2, original index).
PCPX_HD void knn_search_big(const GridView& g, float qx, float qy, float qz, float eps,
                            int start_level, BigHeap& heap)
{
    QueryCell const qc = query_cell(g, qx, qy, qz);
    for (int l = start_level;; --l)
    {
        heap.reset();
        BlockGeom const b   = block_geom(g, qc, l, qx, qy, qz);
        uint64_t const key0 = cell_key(l, b.cx, b.cy, b.cz);
        for (int i = 0; i < 27; ++i)
        {
            Offset3 const o = block27_offset(i);
            if (outside_block(b, o.dx, o.dy, o.dz))
                continue;
            if (cell_lb2(b, o.dx, o.dy, o.dz) > heap.worst_d2())
                continue;
            uint32_t start, count;
            if (!find_cell(g, key0 + key_delta(o.dx, o.dy, o.dz), start, count))
                continue;
            for (uint32_t p = start; p < start + count; ++p)
            {
                float4 const c = load_pt(g.pts + p);
                float const d2 = candidate_d2(c, qx, qy, qz, eps);
                if (d2 < INFINITY)
                    heap.offer(((uint64_t)f2u(d2) << 32) | f2u(c.w));
            }
        }
        if (l == 0 || heap.worst_d2() < b.block_lb2)
            break;
    }
    heap.sort();
}

} // namespace pcpx
