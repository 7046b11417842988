// Grid geometry, cell keys and the hash-table probe shared by the index build and every
// query kernel.  Everything here is __host__ __device__ so that tests/emu can run the SAME
// traversal code on the CPU against the oracle (a unit-test harness, not a product path:
// libpcpx.so only ever launches the __global__ kernels).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define PCPX_HD __host__ __device__ __forceinline__

namespace pcpx {

// ---- exact fp32 arithmetic (never contracted into FMA) -------------------------------------
// The parity contract is the reference's squared_distance evaluated WITHOUT fused multiply-add
// (common/norm.hpp:102-112 in a baseline x86-64 build).  On the device the _rn intrinsics are
// never fused by nvcc; host translation units are compiled with -ffp-contract=off.
PCPX_HD float fmul_x(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
PCPX_HD float fadd_x(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
PCPX_HD float fsub_x(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
PCPX_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
PCPX_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ((dx*dx) + (dy*dy)) + (dz*dz), d = p - t  — common/norm.hpp:108-111
PCPX_HD float sqdist_x(float dx, float dy, float dz)
{
    return fadd_x(fadd_x(fmul_x(dx, dx), fmul_x(dy, dy)), fmul_x(dz, dz));
}

// ---- the grid ------------------------------------------------------------------------------
// Root cube [o, o + extent)^3 covers the root voxel.  Level l splits it into 2^l cells per
// axis (a linear octree); fine coordinates are `lcap` bits per axis.  A cell at level l is
// addressed by its integer coordinates (fine >> (lcap - l)); cells that hold points at levels
// 0..lfine are stored in ONE open-addressing hash table keyed by (level, cx, cy, cz).
constexpr int kMaxLevel        = 19;          // 3 * 19 = 57 coordinate bits + 5 level bits
constexpr uint32_t kEmptyKeyHi = 0xFFFFFFFFu; // key_hi of an empty slot

struct __align__(16) HashSlot
{
    uint32_t key_lo, key_hi; // packed (level, cz, cy, cx)
    uint32_t start, count;   // the cell's span in the sorted point array
};

constexpr uint32_t kPtsPad = 4; // readable (zero) entries behind the sorted point array

struct GridView
{
    float ox, oy, oz; // root cube origin
    float extent;     // root cube side
    float scale;      // 2^lcap / extent: fine cells per unit length
    float delta;      // safety margin for every float-evaluated cell bound (see cell_assign note)
    int32_t lcap;     // bits per axis of the fine coordinates
    int32_t lfine;    // finest level with cells in the table
    uint32_t table_size;
    uint32_t n;       // indexed points (sorted array may hold more: un-indexed tail)
    const HashSlot* table;
    const float4* pts; // sorted by fine Morton code: x, y, z, bits(original index)
};

// cell_assign note.  A point's fine coordinate is u = min(2^lcap - 1, floor((x - o) * scale))
// evaluated in fp32.  Every step is monotone in x, so each cell owns an interval of floats whose
// ends lie within 2^-23 * max(extent, |x|max) (a few ulps) of the nominal o + c * h.  `delta`
// (16 such ulps) is subtracted from every lower bound derived from nominal cell faces, which
// makes all pruning and termination tests conservative in float arithmetic.
PCPX_HD uint32_t quantise(float x, float o, float scale, int lcap)
{
    float t = (x - o) * scale;
    t       = t > 0.f ? t : 0.f; // also NaN -> 0
    float m = (float)((1u << lcap) - 1u);
    t       = t < m ? t : m;
    return (uint32_t)t; // truncation == floor for t >= 0
}

struct QueryCell
{
    uint32_t ux, uy, uz; // fine coordinates
};

PCPX_HD QueryCell query_cell(const GridView& g, float qx, float qy, float qz)
{
    QueryCell c;
    c.ux = quantise(qx, g.ox, g.scale, g.lcap);
    c.uy = quantise(qy, g.oy, g.scale, g.lcap);
    c.uz = quantise(qz, g.oz, g.scale, g.lcap);
    return c;
}

PCPX_HD uint64_t cell_key(int level, uint32_t cx, uint32_t cy, uint32_t cz)
{
    return ((uint64_t)level << 57) | ((uint64_t)cz << 38) | ((uint64_t)cy << 19) | (uint64_t)cx;
}
// key(c + d) = key(c) + key_delta(d): valid while every coordinate stays inside the grid
PCPX_HD uint64_t key_delta(int dx, int dy, int dz)
{
    return (uint64_t)((int64_t)dx + ((int64_t)dy << 19) + ((int64_t)dz << 38));
}

PCPX_HD uint32_t hash_key(uint64_t key)
{
    uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    uint32_t h = lo ^ (hi * 0x85EBCA6Bu) ^ (lo >> 15);
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}
PCPX_HD uint32_t hash_slot(uint64_t key, uint32_t table_size)
{
#ifdef __CUDA_ARCH__
    return __umulhi(hash_key(key), table_size);
#else
    return (uint32_t)(((uint64_t)hash_key(key) * (uint64_t)table_size) >> 32);
#endif
}

PCPX_HD HashSlot load_slot(const HashSlot* p)
{
#ifdef __CUDA_ARCH__
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    HashSlot s;
    s.key_lo = v.x, s.key_hi = v.y, s.start = v.z, s.count = v.w;
    return s;
#else
    return *p;
#endif
}

// Linear probing.  Returns false when the cell is empty (absent from the table).
PCPX_HD bool find_cell(const GridView& g, uint64_t key, uint32_t& start, uint32_t& count)
{
    uint32_t s        = hash_slot(key, g.table_size);
    uint32_t const lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    for (;;)
    {
        HashSlot sl = load_slot(g.table + s);
        if (sl.key_hi == hi && sl.key_lo == lo)
        {
            start = sl.start, count = sl.count;
            return true;
        }
        if (sl.key_hi == kEmptyKeyHi)
            return false;
        s = s + 1 == g.table_size ? 0u : s + 1;
    }
}

PCPX_HD float4 load_pt(const float4* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// 3 x 21-bit Morton interleave (x lowest).  Only used as the SORT key; cells are looked up by
// their packed coordinates, which makes neighbour keys a single 64-bit add.
PCPX_HD uint64_t spread3(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFu;
    x          = (x | (x << 32)) & 0x1F00000000FFFFull;
    x          = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x          = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x          = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x          = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
PCPX_HD uint64_t morton3(uint32_t x, uint32_t y, uint32_t z)
{
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// Visiting order of a 3x3x3 block: own cell, 6 face, 12 edge, 8 corner neighbours (nearer
// cells first, so the list tightens before the far cells are tested).  A run-time lookup instead
// of an unrolled loop keeps the candidate/insert code in the kernel ONCE (instruction cache).
struct Offset3
{
    int dx, dy, dz;
};
PCPX_HD Offset3 block27_offset(int i)
{
    // 27 x 2 bits per axis, value + 1 (immediates: no table in local or constant memory)
    constexpr uint64_t wx = 0x22221562221561ull, wy = 0x28282215681615ull, wz = 0x2A802828156155ull;
    int const sh = 2 * i;
    Offset3 o;
    o.dx = (int)((wx >> sh) & 3u) - 1;
    o.dy = (int)((wy >> sh) & 3u) - 1;
    o.dz = (int)((wz >> sh) & 3u) - 1;
    return o;
}

} // namespace pcpx
