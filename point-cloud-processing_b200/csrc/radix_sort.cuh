// Hand-written LSD radix sort of (key, uint32 value) pairs, 8 bits per pass, stable.
//
// One `digit_histograms` kernel counts every pass's digits in a single read of the keys.  Each
// pass is then ONE kernel, `scatter_lookback`: warp-level multi-split ranking (match.any), the
// tile's digit counts published and the cross-tile offsets resolved by decoupled look-back,
// tile-local reorder in shared memory, coalesced runs out to global memory.
// Algorithmic traffic per element: K (histograms) + 2 P (K + 4) (P passes, K = key bytes).
// For n >= 2^30 (counts would collide with the flag bits) the three-kernel pass is used instead:
//   tile_histogram -> tile_offsets (exclusive scan per digit across tiles) -> scatter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcpx {
namespace rsort {

constexpr int kRadixBits   = 8;
constexpr int kRadix       = 1 << kRadixBits;
constexpr int kThreads     = 256;
constexpr int kWarps       = kThreads / 32;
constexpr int kItems       = 16;
constexpr int kTile        = kThreads * kItems; // 4096 pairs per CTA
constexpr int kMaxPasses   = 8;

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, int shift)
{
    return (uint32_t)(key >> shift) & (kRadix - 1);
}

// hist[pass][digit] over all keys, every pass at once.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) digit_histograms(
    const KeyT* __restrict__ keys, uint32_t n, int first_shift, int n_passes,
    uint32_t* __restrict__ hist /* [n_passes][256] */)
{
    __shared__ uint32_t sh[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < n_passes * kRadix; i += kThreads)
        sh[i] = 0;
    __syncthreads();
    uint32_t const stride = gridDim.x * kThreads;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    {
        KeyT const key = keys[i];
        // plain shared-memory atomics: lanes mostly hit different bins, and ptxas aggregates the
        // lanes that do share one — an order of magnitude cheaper than ranking with match.any
        for (int p = 0; p < n_passes; ++p)
            atomicAdd(&sh[p * kRadix + digit_of(key, first_shift + p * kRadixBits)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * kRadix; i += kThreads)
        if (sh[i])
            atomicAdd(&hist[i], sh[i]);
}

// counts[digit][tile]
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) tile_histogram(
    const KeyT* __restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
    uint32_t* __restrict__ counts)
{
    __shared__ uint32_t sh[kRadix];
    sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t const base = blockIdx.x * kTile;
#pragma unroll 4
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const i = base + r * kThreads + threadIdx.x;
        bool const valid = i < n;
        uint32_t const d = valid ? digit_of(keys[i], shift) : 0xFFFFFFFFu;
        uint32_t const peers = __match_any_sync(0xFFFFFFFFu, d);
        if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1)
            atomicAdd(&sh[d], __popc(peers));
    }
    __syncthreads();
    counts[threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

// One CTA per digit: exclusive scan of that digit's per-tile counts, offset by the digit's base
// (sum of the pass histogram over smaller digits).  In place.
__global__ void __launch_bounds__(kThreads) tile_offsets(
    uint32_t* __restrict__ counts, uint32_t n_tiles, const uint32_t* __restrict__ pass_hist)
{
    __shared__ uint32_t warp_sums[kWarps];
    __shared__ uint32_t carry;
    uint32_t const d = blockIdx.x;
    if (threadIdx.x == 0)
    {
        uint32_t b = 0;
        for (uint32_t j = 0; j < d; ++j)
            b += pass_hist[j];
        carry = b;
    }
    __syncthreads();
    uint32_t* row = counts + (size_t)d * n_tiles;
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += kThreads)
    {
        uint32_t const t = t0 + threadIdx.x;
        uint32_t const v = t < n_tiles ? row[t] : 0u;
        uint32_t incl    = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)(threadIdx.x & 31) >= o)
                incl += up;
        }
        if ((threadIdx.x & 31) == 31)
            warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w)
            wbase += warp_sums[w];
        uint32_t const c = carry;
        if (t < n_tiles)
            row[t] = c + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1)
            carry = c + wbase + incl;
        __syncthreads();
    }
}

template <typename KeyT>
struct ScatterSmem
{
    KeyT keys[kTile];
    uint32_t vals[kTile];
    uint32_t warp_hist[kWarps][kRadix];
    uint32_t digit_start[kRadix];  // start of each digit inside the tile-sorted order
    uint32_t global_base[kRadix];  // where this tile's run of each digit goes in the output
};

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) scatter(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
    uint32_t n_tiles, const uint32_t* __restrict__ offsets /* [digit][tile] */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem<KeyT>& sm = *reinterpret_cast<ScatterSmem<KeyT>*>(smem_raw);
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1u << lane) - 1u;

    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads)
        (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();

    // each warp owns a contiguous chunk of 32 * kItems pairs so that (warp, round, lane) order
    // is the input order — that is what keeps the sort stable
    uint32_t const tile_base = blockIdx.x * kTile;
    uint32_t const warp_base = tile_base + warp * (32 * kItems);
    uint32_t const tile_n    = min((uint32_t)kTile, n - tile_base);

    KeyT key[kItems];
    uint32_t rank[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const i = warp_base + r * 32 + lane;
        bool const valid = i < n;
        key[r]           = valid ? keys_in[i] : (KeyT)0;
        uint32_t const d = valid ? digit_of(key[r], shift) : 0xFFFFFFFFu;
        uint32_t const peers = __match_any_sync(0xFFFFFFFFu, d);
        int const leader     = __ffs(peers) - 1;
        uint32_t before      = 0;
        if (valid && lane == leader)
        {
            before               = sm.warp_hist[warp][d];
            sm.warp_hist[warp][d] = before + __popc(peers);
        }
        before  = __shfl_sync(0xFFFFFFFFu, before, leader);
        rank[r] = before + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive prefix over warps, digit totals -> exclusive scan over digits
    {
        uint32_t const d = threadIdx.x; // kThreads == kRadix
        uint32_t run     = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
        {
            uint32_t const c   = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = run;
            run += c;
        }
        // block-wide exclusive scan of `run` (digit totals)
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o)
                incl += up;
        }
        __shared__ uint32_t wsum[kWarps];
        if (lane == 31)
            wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w)
            wbase += wsum[w];
        sm.digit_start[d] = wbase + incl - run;
        sm.global_base[d] = offsets[(size_t)d * n_tiles + blockIdx.x];
    }
    __syncthreads();

    // tile-local reorder
#pragma unroll
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const i = warp_base + r * 32 + lane;
        if (i < n)
        {
            uint32_t const d   = digit_of(key[r], shift);
            uint32_t const pos = sm.digit_start[d] + sm.warp_hist[warp][d] + rank[r];
            sm.keys[pos]       = key[r];
            sm.vals[pos]       = vals_in[i];
        }
    }
    __syncthreads();

    // runs of equal digits go out to consecutive addresses
    for (uint32_t i = threadIdx.x; i < tile_n; i += kThreads)
    {
        KeyT const k       = sm.keys[i];
        uint32_t const d   = digit_of(k, shift);
        uint32_t const dst = sm.global_base[d] + (i - sm.digit_start[d]);
        keys_out[dst]      = k;
        vals_out[dst]      = sm.vals[i];
    }
}

// ---- single-kernel pass: the scatter computes its own tile histogram and resolves the
// cross-tile digit offsets by decoupled look-back, so a pass reads the keys once instead of
// twice and needs one launch instead of three.
//
// state[tile][digit] = count | flag.  A tile first publishes its local digit counts
// (kFlagLocal), then walks back over its predecessors adding their words until it meets an
// inclusive one (kFlagInclusive), and publishes its own inclusive prefix.  Tiles take their id
// from a ticket counter, so a tile only ever waits on tiles that started before it.
constexpr uint32_t kFlagLocal     = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kFlagMask      = 3u << 30;
constexpr uint32_t kLookbackMaxN  = 1u << 30; // counts share a word with two flag bits

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) scatter_lookback(
    const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
    const uint32_t* __restrict__ pass_hist /* [256] digit totals of this pass */,
    uint32_t* __restrict__ state /* [n_tiles][256], zeroed */, uint32_t* __restrict__ ticket)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem<KeyT>& sm = *reinterpret_cast<ScatterSmem<KeyT>*>(smem_raw);
    __shared__ uint32_t s_tile;
    __shared__ uint32_t wsum[kWarps];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t const lt_mask = (1u << lane) - 1u;

    if (threadIdx.x == 0)
        s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads)
        (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t const tile = s_tile;

    uint32_t const tile_base = tile * kTile;
    uint32_t const warp_base = tile_base + warp * (32 * kItems);
    uint32_t const tile_n    = min((uint32_t)kTile, n - tile_base);

    KeyT key[kItems];
    uint32_t rank[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const i = warp_base + r * 32 + lane;
        bool const valid = i < n;
        key[r]           = valid ? keys_in[i] : (KeyT)0;
        uint32_t const d = valid ? digit_of(key[r], shift) : 0xFFFFFFFFu;
        uint32_t const peers = __match_any_sync(0xFFFFFFFFu, d);
        int const leader     = __ffs(peers) - 1;
        uint32_t before      = 0;
        if (valid && lane == leader)
        {
            before                = sm.warp_hist[warp][d];
            sm.warp_hist[warp][d] = before + __popc(peers);
        }
        before  = __shfl_sync(0xFFFFFFFFu, before, leader);
        rank[r] = before + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    {
        uint32_t const d = threadIdx.x; // kThreads == kRadix
        uint32_t run     = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
        {
            uint32_t const c   = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = run;
            run += c;
        }
        // publish the local count at once so that successors can make progress
        volatile uint32_t* my = state + (size_t)tile * kRadix + d;
        if (tile > 0)
            *my = run | kFlagLocal;
        else
            *my = run | kFlagInclusive;

        // block-wide exclusive scans over digits: position inside the tile, global digit base
        uint32_t const total = pass_hist[d];
        uint32_t incl = run, incl_g = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const up  = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            uint32_t const upg = __shfl_up_sync(0xFFFFFFFFu, incl_g, o);
            if (lane >= o)
                incl += up, incl_g += upg;
        }
        __shared__ uint32_t wsum_g[kWarps];
        if (lane == 31)
            wsum[warp] = incl, wsum_g[warp] = incl_g;
        __syncthreads();
        uint32_t wbase = 0, wbase_g = 0;
        for (int w = 0; w < warp; ++w)
            wbase += wsum[w], wbase_g += wsum_g[w];
        sm.digit_start[d]         = wbase + incl - run;
        uint32_t const digit_base = wbase_g + incl_g - total;

        // decoupled look-back over the predecessor tiles
        uint32_t excl = 0;
        if (tile > 0)
        {
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t)
            {
                volatile uint32_t* p = state + (size_t)t * kRadix + d;
                uint32_t v           = *p;
                while ((v & kFlagMask) == 0u)
                    v = *p;
                excl += v & ~kFlagMask;
                if (v & kFlagInclusive)
                    break;
            }
            *my = (excl + run) | kFlagInclusive;
        }
        sm.global_base[d] = digit_base + excl;
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < kItems; ++r)
    {
        uint32_t const i = warp_base + r * 32 + lane;
        if (i < n)
        {
            uint32_t const d   = digit_of(key[r], shift);
            uint32_t const pos = sm.digit_start[d] + sm.warp_hist[warp][d] + rank[r];
            sm.keys[pos]       = key[r];
            sm.vals[pos]       = vals_in[i];
        }
    }
    __syncthreads();

    for (uint32_t i = threadIdx.x; i < tile_n; i += kThreads)
    {
        KeyT const k       = sm.keys[i];
        uint32_t const d   = digit_of(k, shift);
        uint32_t const dst = sm.global_base[d] + (i - sm.digit_start[d]);
        keys_out[dst]      = k;
        vals_out[dst]      = sm.vals[i];
    }
}

} // namespace rsort
} // namespace pcpx
