// Normal orientation by breadth-first propagation over the directed kNN graph (SURVEY.md §8f
// rank 4; reference: algorithm/estimate_normals.hpp:187-302, graph/knn_adjacency_list.hpp:117-156,
// graph/search.hpp:41-85).
//
// The reference's search is a sequential FIFO walk, and which edge first reaches a vertex decides
// the vertex's sign, so the result is defined by the queue order.  That order is reproduced
// exactly, level by level, without a queue:
//   * the queue content of one BFS level is an array `frontier` in queue order;
//   * edge e of frontier entry i has the code i * k + e; a vertex not reached in earlier levels
//     is reached by the edge of MINIMAL code pointing at it (atomicMin on a 64-bit key);
//   * the next level's queue order is the order of the winning codes, i.e. an ordered compaction
//     (per-entry count -> exclusive scan -> emit), not a sort.
// All sizes stay on the device; the host queues levels in batches and only reads back the
// frontier size once per batch to see whether the search has ended.  No CPU path.
#include "api_util.hpp"

using namespace pcpx;

namespace {

constexpr int kB         = 256;
constexpr int kGrid      = 148;
constexpr int kChunk     = 8;  // edges whose loads are in flight together
constexpr uint32_t kPad  = 0xFFFFFFFFu;

struct BfsState
{
    uint32_t size[2];    // frontier sizes, ping-pong by level parity
    uint32_t levels;     // levels that reached at least one new vertex
    uint32_t done;       // blocks of the current accept kernel that have finished
    unsigned long long reached; // vertices reached, root included
    unsigned long long root_key;
};

__device__ __forceinline__ uint32_t orderable(float z)
{
    uint32_t const b = __float_as_uint(z + 0.f); // -0 -> +0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// root = FIRST point of maximal z in input order (std::max_element, estimate_normals.hpp:223-230)
__global__ void __launch_bounds__(kB) root_kernel(const float4* __restrict__ pts, uint32_t n,
                                                  BfsState* st)
{
    unsigned long long best = 0;
    for (uint32_t t = blockIdx.x * kB + threadIdx.x; t < n; t += gridDim.x * kB)
    {
        float4 const p = pts[t];
        unsigned long long const key =
            ((unsigned long long)orderable(p.z) << 32) | (0xFFFFFFFFu - __float_as_uint(p.w));
        best = key > best ? key : best;
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long const other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best)
        atomicMax(&st->root_key, best);
}

// relabel: original index -> vertex id of the search (nullptr: the ids are the original indices)
__global__ void start_kernel(BfsState* st, uint32_t* frontier, uint8_t* visited, float* normals,
                             const uint32_t* relabel)
{
    uint32_t root = 0xFFFFFFFFu - (uint32_t)(st->root_key & 0xFFFFFFFFu);
    if (relabel)
        root = relabel[root];
    frontier[0]         = root;
    visited[root]       = 1;
    normals[3 * (size_t)root] = 0.f, normals[3 * (size_t)root + 1] = 0.f,
                    normals[3 * (size_t)root + 2] = 1.f; // :233-237
    st->size[0] = 1, st->size[1] = 0;
    st->levels = 0, st->reached = 1, st->done = 0;
}

__device__ __forceinline__ uint32_t edge_target(const uint32_t* __restrict__ nbr, uint32_t u,
                                                uint32_t k, uint32_t e, int reverse)
{
    return nbr[(size_t)u * k + (reverse ? k - 1 - e : e)];
}

// ---- the four phases of one BFS level ---------------------------------------------------------
// Three small kernels per level.  (A single cooperative kernel with grid barriers between the
// phases was measured and is not faster: 49 - 62 ms against 48 - 55 ms for 996 levels over 5 M
// points; the level time is the chain of dependent random loads, not the kernel boundaries.)
// Arrays other phases write (frontier, visited, key, won, normals, the state) are read with
// ld.global.cg and are deliberately not const __restrict__.

__device__ __forceinline__ void propose_phase(const BfsState* st, int parity,
                                              const uint32_t* frontier,
                                              const uint32_t* __restrict__ nbr, uint32_t k,
                                              int reverse, const uint8_t* visited,
                                              unsigned long long* key, uint32_t tid,
                                              uint32_t nthreads)
{
    unsigned long long const total = (unsigned long long)__ldcg(&st->size[parity]) * k;
    if (total <= 0xFFFFFFFFull) // the usual case: 32-bit index arithmetic (k is not a power of 2)
    {
        for (uint32_t c = tid; c < (uint32_t)total; c += nthreads)
        {
            uint32_t const i = c / k, e = c - i * k;
            uint32_t const v = edge_target(nbr, __ldcg(frontier + i), k, e, reverse);
            if (v != kPad && !__ldcg(visited + v))
                atomicMin(&key[v], (unsigned long long)c);
        }
        return;
    }
    for (unsigned long long c = tid; c < total; c += nthreads)
    {
        uint32_t const i = (uint32_t)(c / k), e = (uint32_t)(c % k);
        uint32_t const v = edge_target(nbr, __ldcg(frontier + i), k, e, reverse);
        if (v != kPad && !__ldcg(visited + v))
            atomicMin(&key[v], c);
    }
}

// per frontier entry: flip the children it wins (search.hpp:72-78 + estimate_normals.hpp:289-300)
// and count them
__device__ __forceinline__ void accept_phase(const BfsState* st, int parity,
                                             const uint32_t* frontier,
                                             const uint32_t* __restrict__ nbr, uint32_t k,
                                             int reverse, const uint8_t* visited,
                                             const unsigned long long* key, float* normals,
                                             uint32_t* won, uint32_t tid, uint32_t nthreads)
{
    uint32_t const m = __ldcg(&st->size[parity]);
    for (uint32_t i = tid; i < m; i += nthreads)
    {
        uint32_t const u = __ldcg(frontier + i);
        float const ax = __ldcg(normals + 3 * (size_t)u), ay = __ldcg(normals + 3 * (size_t)u + 1),
                    az = __ldcg(normals + 3 * (size_t)u + 2);
        uint32_t cnt = 0;
        for (uint32_t e0 = 0; e0 < k; e0 += kChunk)
        {
            // the chunk's loads are issued together (three dependent rounds per chunk instead of
            // three per edge); the flips below still happen in edge order
            uint32_t v[kChunk];
            bool mine[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
            {
                uint32_t const t = edge_target(nbr, u, k, e0 + c < k ? e0 + c : 0u, reverse);
                v[c]             = e0 + c < k ? t : kPad;
            }
            // every load of the chunk is issued before any is used: no short-circuit, absent
            // edges read entry 0 and are masked afterwards
            uint8_t seen[kChunk];
            unsigned long long owner[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
            {
                uint32_t const vv = v[c] != kPad ? v[c] : 0u;
                seen[c]  = __ldcg(visited + vv);
                owner[c] = __ldcg(key + vv);
            }
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
                mine[c] = (v[c] != kPad) & (seen[c] == 0) &
                          (owner[c] == (unsigned long long)i * k + e0 + c);
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
            {
                if (!mine[c])
                    continue;
                float* nv      = normals + 3 * (size_t)v[c];
                float const bx = __ldcg(nv), by = __ldcg(nv + 1), bz = __ldcg(nv + 2);
                // common::inner_product (common/norm.hpp:34-45): v2 * v1 per axis, left to right
                float const prod =
                    __fadd_rn(__fadd_rn(__fmul_rn(bx, ax), __fmul_rn(by, ay)), __fmul_rn(bz, az));
                if (prod < 0.f && !(fabsf(prod - 0.f) < 1e-5f))
                    nv[0] = -bx, nv[1] = -by, nv[2] = -bz;
                ++cnt;
            }
        }
        won[i] = cnt;
    }
}

// exclusive scan of won[0 .. m) in place by ONE block of up to 1024 threads (the sum of all
// frontier sizes over a whole search is n, so this never does more than n / blockDim rounds in
// total); publishes the next frontier size
__device__ __forceinline__ void scan_phase(BfsState* st, int parity, uint32_t* won)
{
    constexpr int kScanItems = 8; // consecutive entries per thread and round
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    uint32_t const m      = __ldcg(&st->size[parity]);
    uint32_t const nwarps = blockDim.x >> 5;
    uint32_t const span   = blockDim.x * kScanItems;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += span)
    {
        uint32_t const i0 = base + threadIdx.x * kScanItems;
        uint32_t x[kScanItems];
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < kScanItems; ++r)
        {
            x[r] = i0 + r < m ? __ldcg(won + i0 + r) : 0u;
            mine += x[r];
        }
        uint32_t incl = mine;
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o)
                incl += y;
        }
        if ((threadIdx.x & 31) == 31)
            warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            uint32_t w = threadIdx.x < nwarps ? warp_sums[threadIdx.x] : 0u;
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t const y = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (threadIdx.x >= o)
                    w += y;
            }
            warp_sums[threadIdx.x] = w; // inclusive over warps
        }
        __syncthreads();
        uint32_t run = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0u) +
                       incl - mine;
#pragma unroll
        for (int r = 0; r < kScanItems; ++r)
        {
            if (i0 + r < m)
                won[i0 + r] = run;
            run += x[r];
        }
        __syncthreads();
        if (threadIdx.x == 0)
            carry += warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        st->size[parity ^ 1] = carry;
        st->reached += carry;
        st->levels += carry ? 1u : 0u;
    }
}

__device__ __forceinline__ void emit_phase(const BfsState* st, int parity, const uint32_t* frontier,
                                           const uint32_t* __restrict__ nbr, uint32_t k,
                                           int reverse, uint8_t* visited,
                                           const unsigned long long* key, const uint32_t* offset,
                                           uint32_t* next, uint32_t tid, uint32_t nthreads)
{
    uint32_t const m = __ldcg(&st->size[parity]);
    for (uint32_t i = tid; i < m; i += nthreads)
    {
        uint32_t const u = __ldcg(frontier + i);
        uint32_t w       = __ldcg(offset + i);
        for (uint32_t e0 = 0; e0 < k; e0 += kChunk)
        {
            uint32_t v[kChunk];
            bool mine[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
            {
                uint32_t const t = edge_target(nbr, u, k, e0 + c < k ? e0 + c : 0u, reverse);
                v[c]             = e0 + c < k ? t : kPad;
            }
            // the owner of v is unique, so nobody else tests visited[v] with a matching key
            uint8_t seen[kChunk];
            unsigned long long owner[kChunk];
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
            {
                uint32_t const vv = v[c] != kPad ? v[c] : 0u;
                seen[c]  = __ldcg(visited + vv);
                owner[c] = __ldcg(key + vv);
            }
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
                mine[c] = (v[c] != kPad) & (seen[c] == 0) &
                          (owner[c] == (unsigned long long)i * k + e0 + c);
#pragma unroll
            for (int c = 0; c < kChunk; ++c)
                if (mine[c])
                {
                    next[w++]     = v[c];
                    visited[v[c]] = 1;
                }
        }
    }
}

__global__ void __launch_bounds__(kB) propose_kernel(const BfsState* st, int parity,
                                                     const uint32_t* frontier, const uint32_t* nbr,
                                                     uint32_t k, int reverse,
                                                     const uint8_t* visited,
                                                     unsigned long long* key)
{
    propose_phase(st, parity, frontier, nbr, k, reverse, visited, key,
                  blockIdx.x * kB + threadIdx.x, gridDim.x * kB);
}
// accept + scan in one launch: the block that finishes last (a ticket counter, no waiting) scans
// the counts of the whole frontier.  One launch less per level is worth more than the three
// warps-worth of scan parallelism it gives up: a level is bound by the ~9 us each dependent
// launch costs, not by the work.
__global__ void __launch_bounds__(kB) accept_scan_kernel(BfsState* st, int parity,
                                                         const uint32_t* frontier,
                                                         const uint32_t* nbr, uint32_t k,
                                                         int reverse, const uint8_t* visited,
                                                         const unsigned long long* key,
                                                         float* normals, uint32_t* won)
{
    accept_phase(st, parity, frontier, nbr, k, reverse, visited, key, normals, won,
                 blockIdx.x * kB + threadIdx.x, gridDim.x * kB);
    __shared__ bool is_last;
    __threadfence(); // this block's counts are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0)
        is_last = atomicAdd(&st->done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!is_last)
        return;
    __threadfence();
    scan_phase(st, parity, won);
    if (threadIdx.x == 0)
        st->done = 0;
}
__global__ void __launch_bounds__(kB) emit_kernel(const BfsState* st, int parity,
                                                  const uint32_t* frontier, const uint32_t* nbr,
                                                  uint32_t k, int reverse, uint8_t* visited,
                                                  const unsigned long long* key,
                                                  const uint32_t* offset, uint32_t* next)
{
    emit_phase(st, parity, frontier, nbr, k, reverse, visited, key, offset, next,
               blockIdx.x * kB + threadIdx.x, gridDim.x * kB);
}

// ---- the search in the index's own order -----------------------------------------------------
// The queue order of the search depends on the graph and the edge order only, not on how the
// vertices are numbered.  With the original indices as ids every access of a level — neighbour
// row, visited flag, key, normal — lands on a random page; numbered by SORTED position the
// neighbours of a vertex are a few thousand ids away at most, and a level stays inside a few
// pages (TLB) and L2 lines.
__global__ void __launch_bounds__(kB) inverse_kernel(const float4* __restrict__ pts, uint32_t n,
                                                     uint32_t* __restrict__ inv)
{
    uint32_t const t = blockIdx.x * kB + threadIdx.x;
    if (t < n)
        inv[__float_as_uint(pts[t].w)] = t;
}

__global__ void __launch_bounds__(kB) relabel_graph_kernel(const float4* __restrict__ pts,
                                                           uint32_t n, uint32_t k,
                                                           const uint32_t* __restrict__ nbr,
                                                           const uint32_t* __restrict__ inv,
                                                           uint32_t* __restrict__ nbr_sorted)
{
    unsigned long long const c = blockIdx.x * (unsigned long long)kB + threadIdx.x;
    if (c >= (unsigned long long)n * k)
        return;
    uint32_t const pos = (uint32_t)(c / k), j = (uint32_t)(c % k);
    uint32_t const v   = nbr[(size_t)__float_as_uint(pts[pos].w) * k + j];
    nbr_sorted[c]      = v == kPad ? kPad : inv[v];
}

// forward: rows in input order -> rows in sorted order; !forward: back
__global__ void __launch_bounds__(kB) permute_rows_kernel(const float4* __restrict__ pts,
                                                          uint32_t n, const float* __restrict__ in,
                                                          float* __restrict__ out, int forward)
{
    uint32_t const t = blockIdx.x * kB + threadIdx.x;
    if (t >= n)
        return;
    size_t const o   = __float_as_uint(pts[t].w);
    const float* src = in + 3 * (forward ? o : (size_t)t);
    float* dst       = out + 3 * (forward ? (size_t)t : o);
    dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
}

// same for a caller-supplied graph: points in input order, packed or strided
__global__ void __launch_bounds__(kB) root_from_rows_kernel(const float* __restrict__ xyz,
                                                            uint32_t stride_f, uint32_t n,
                                                            BfsState* st)
{
    unsigned long long best = 0;
    for (uint32_t t = blockIdx.x * kB + threadIdx.x; t < n; t += gridDim.x * kB)
    {
        unsigned long long const key =
            ((unsigned long long)orderable(xyz[(size_t)t * stride_f + 2]) << 32) |
            (0xFFFFFFFFu - t);
        best = key > best ? key : best;
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long const other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best)
        atomicMax(&st->root_key, best);
}

uint32_t grid_of(size_t n) { return (uint32_t)std::min<size_t>(std::max<size_t>((n + kB - 1) / kB, 1), kGrid); }

// The search proper.  `st` holds the root key; nbr = n rows of k targets (kPad = no edge).
BfsState run_search(cudaStream_t s, size_t n, uint32_t k, const uint32_t* nbr, int reverse,
                    float* d_nrm, BfsState* st, uint32_t& launches,
                    const uint32_t* relabel = nullptr)
{
    DevBuf<uint32_t> fa(n), fb(n), won(n);
    DevBuf<uint8_t> visited(n);
    DevBuf<unsigned long long> key(n);
    PCPX_CUDA(cudaMemsetAsync(visited.get(), 0, n, s));
    PCPX_CUDA(cudaMemsetAsync(key.get(), 0xFF, n * 8, s));
    start_kernel<<<1, 1, 0, s>>>(st, fa.get(), visited.get(), d_nrm, relabel);
    PCPX_CHECK_LAUNCH();
    ++launches;
    BfsState h{};
    uint32_t const gn = grid_of(n);
    if (k == 0)
    {
        PCPX_CUDA(cudaMemcpyAsync(&h, st, sizeof h, cudaMemcpyDeviceToHost, s));
        PCPX_CUDA(cudaStreamSynchronize(s));
        return h;
    }
    // kLevelsPerBatch levels (3 kernels each) are captured ONCE into a CUDA graph and the graph is
    // launched per batch: the search is otherwise bound by the host's launch rate (4 100 launches
    // for 996 levels took as long as the kernels).  The batch is an even number of levels, so the
    // two frontier buffers and the parity are back where they started and the graph can be
    // replayed as is.  Levels past the end of the search see an empty frontier and do nothing;
    // the host looks at the frontier size once per batch.
    constexpr int kLevelsPerBatch = 32;
    struct GraphGuard
    {
        cudaGraph_t graph   = nullptr;
        cudaGraphExec_t exe = nullptr;
        ~GraphGuard()
        {
            if (exe)
                cudaGraphExecDestroy(exe);
            if (graph)
                cudaGraphDestroy(graph);
        }
    } gg;
    PCPX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    {
        uint32_t *cur = fa.get(), *nxt = fb.get();
        int parity = 0;
        for (int b = 0; b < kLevelsPerBatch; ++b)
        {
            propose_kernel<<<gn, kB, 0, s>>>(st, parity, cur, nbr, k, reverse, visited.get(),
                                             key.get());
            accept_scan_kernel<<<gn, kB, 0, s>>>(st, parity, cur, nbr, k, reverse, visited.get(),
                                                 key.get(), d_nrm, won.get());
            emit_kernel<<<gn, kB, 0, s>>>(st, parity, cur, nbr, k, reverse, visited.get(),
                                          key.get(), won.get(), nxt);
            std::swap(cur, nxt);
            parity ^= 1;
        }
    }
    cudaError_t const cap = cudaStreamEndCapture(s, &gg.graph);
    if (cap != cudaSuccess)
        fail(PCPX_ERR_CUDA, "stream capture of the orientation search failed: %s",
             cudaGetErrorString(cap));
    PCPX_CUDA(cudaGraphInstantiate(&gg.exe, gg.graph, 0));
    for (uint64_t level = 0; level < n; level += kLevelsPerBatch)
    {
        PCPX_CUDA(cudaGraphLaunch(gg.exe, s));
        launches += 3 * kLevelsPerBatch;
        PCPX_CUDA(cudaMemcpyAsync(&h, st, sizeof h, cudaMemcpyDeviceToHost, s));
        PCPX_CUDA(cudaStreamSynchronize(s));
        if (h.size[0] == 0) // an even number of levels per batch: the live frontier is parity 0
            return h;
    }
    PCPX_CUDA(cudaMemcpyAsync(&h, st, sizeof h, cudaMemcpyDeviceToHost, s));
    PCPX_CUDA(cudaStreamSynchronize(s)); // also keeps the buffers above alive until done
    return h;
}

void check_edge_order(int edge_order)
{
    if (edge_order != PCPX_EDGES_FURTHEST_FIRST && edge_order != PCPX_EDGES_NEAREST_FIRST)
        fail(PCPX_ERR_INVALID_ARG,
             "edge_order must be PCPX_EDGES_FURTHEST_FIRST or PCPX_EDGES_NEAREST_FIRST");
}

} // namespace

extern "C" {

int pcpx_orient_normals(const pcpx_index* index, uint32_t k, double eps, int edge_order,
                        float* normals, uint32_t* out_levels, uint64_t* out_reached)
{
    return guarded([&] {
        pcpx_index& ix = checked(index);
        size_t const n = ix.n_input;
        if (out_levels)
            *out_levels = 0;
        if (out_reached)
            *out_reached = 0;
        if (n == 0)
            return;
        if (!normals)
            fail(PCPX_ERR_INVALID_ARG, "normals is NULL");
        check_edge_order(edge_order);
        std::lock_guard<std::mutex> lock(ix.mtx);
        ScopedDevice guard(ix.device);
        CallTimer timer(ix);
        cudaStream_t const s = ix.stream;

        // normals in place on the device
        bool const direct = is_device_pointer(normals);
        DevBuf<float> staged;
        float* d_nrm = normals;
        if (!direct)
        {
            staged.alloc(3 * n);
            PCPX_CUDA(cudaMemcpyAsync(staged.get(), normals, 12 * n, cudaMemcpyHostToDevice, s));
            d_nrm = staged.get();
        }

        // the directed kNN graph (graph/knn_adjacency_list.hpp:139-152), rows in input order
        uint32_t const kk = std::max(k, 1u);
        DevBuf<uint32_t> nbr(n * (size_t)kk);
        timer.kernel_begin();
        if (k > 0)
        {
            DevBuf<uint32_t> retries(1);
            PCPX_CUDA(cudaMemsetAsync(retries.get(), 0, 4, s));
            QueryBatch const qb{nullptr, 3u, nullptr, (uint32_t)n};
            launch_knn(ix, qb, k, (float)eps, nbr.get(), nullptr, nullptr, retries.get());
            PCPX_CUDA(cudaStreamSynchronize(s)); // `retries` is released here
        }
        // renumber the graph and the normals by sorted position (see inverse_kernel)
        uint32_t const n32 = (uint32_t)n;
        uint32_t const nb  = (uint32_t)((n + kB - 1) / kB);
        DevBuf<uint32_t> inv(n), nbr_sorted(n * (size_t)kk);
        DevBuf<float> nrm_sorted(3 * n);
        inverse_kernel<<<nb, kB, 0, s>>>(ix.grid.pts, n32, inv.get());
        if (k > 0)
            relabel_graph_kernel<<<(uint32_t)((n * (size_t)k + kB - 1) / kB), kB, 0, s>>>(
                ix.grid.pts, n32, k, nbr.get(), inv.get(), nbr_sorted.get());
        permute_rows_kernel<<<nb, kB, 0, s>>>(ix.grid.pts, n32, d_nrm, nrm_sorted.get(), 1);
        PCPX_CHECK_LAUNCH();
        DevBuf<BfsState> st(1);
        PCPX_CUDA(cudaMemsetAsync(st.get(), 0, sizeof(BfsState), s));
        root_kernel<<<grid_of(n), kB, 0, s>>>(ix.grid.pts, n32, st.get());
        uint32_t launches = 6;
        BfsState const h =
            run_search(s, n, k, nbr_sorted.get(), edge_order == PCPX_EDGES_FURTHEST_FIRST,
                       nrm_sorted.get(), st.get(), launches, inv.get());
        permute_rows_kernel<<<nb, kB, 0, s>>>(ix.grid.pts, n32, nrm_sorted.get(), d_nrm, 0);
        PCPX_CHECK_LAUNCH();
        ++launches;
        timer.kernel_end();
        if (!direct)
            PCPX_CUDA(cudaMemcpyAsync(normals, d_nrm, 12 * n, cudaMemcpyDeviceToHost, s));
        timer.done();
        ix.timings.kernel_launches = launches;
        if (out_levels)
            *out_levels = h.levels;
        if (out_reached)
            *out_reached = h.reached;
    });
}

int pcpx_orient_normals_graph(const float* xyz, size_t n, size_t stride_bytes,
                              const uint32_t* neighbours, uint32_t k, int edge_order, int device,
                              float* normals, uint32_t* out_levels, uint64_t* out_reached)
{
    return guarded([&] {
        if (out_levels)
            *out_levels = 0;
        if (out_reached)
            *out_reached = 0;
        if (n == 0)
            return;
        if (n >= 0xFFFFFFFFull)
            fail(PCPX_ERR_UNSUPPORTED, "graphs of 2^32 - 1 vertices or more need 64-bit indices");
        if (!xyz || !normals || (k && !neighbours))
            fail(PCPX_ERR_INVALID_ARG, "xyz / neighbours / normals is NULL");
        check_edge_order(edge_order);
        if (stride_bytes == 0)
            stride_bytes = 12;
        if (stride_bytes < 12 || stride_bytes % 4)
            fail(PCPX_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and >= 12");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            fail(PCPX_ERR_NO_DEVICE, "no CUDA device: libpcpx has no CPU path");
        if (device < 0)
            PCPX_CUDA(cudaGetDevice(&device));
        if (device >= ndev)
            fail(PCPX_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, ndev);
        ScopedDevice guard(device);
        cudaStream_t s = nullptr;
        PCPX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        struct StreamGuard
        {
            cudaStream_t s;
            ~StreamGuard() { cudaStreamDestroy(s); }
        } sg{s};

        InBuf pts;
        pts.stage(xyz, n, stride_bytes, 3, s);
        DevBuf<uint32_t> nbr_staged;
        const uint32_t* d_nbr = neighbours;
        if (k && !is_device_pointer(neighbours))
        {
            for (size_t i = 0; i < n * (size_t)k; ++i)
                if (neighbours[i] != kPad && neighbours[i] >= n)
                    fail(PCPX_ERR_INVALID_ARG, "neighbours[%llu] = %u is not a vertex",
                         (unsigned long long)i, neighbours[i]);
            nbr_staged.alloc(n * (size_t)k);
            PCPX_CUDA(cudaMemcpyAsync(nbr_staged.get(), neighbours, n * (size_t)k * 4,
                                      cudaMemcpyHostToDevice, s));
            d_nbr = nbr_staged.get();
        }
        else if (k)
        {
            // device-resident rows: checked on the device before the search dereferences them
            uint32_t const bad = count_bad_indices(s, neighbours, n * (size_t)k, (uint32_t)n, kPad);
            if (bad)
                fail(PCPX_ERR_INVALID_ARG, "%u entries of neighbours are not vertices (< n) nor "
                     "PCPX_NO_NEIGHBOUR", bad);
        }
        bool const direct = is_device_pointer(normals);
        DevBuf<float> staged;
        float* d_nrm = normals;
        if (!direct)
        {
            staged.alloc(3 * n);
            PCPX_CUDA(cudaMemcpyAsync(staged.get(), normals, 12 * n, cudaMemcpyHostToDevice, s));
            d_nrm = staged.get();
        }
        DevBuf<BfsState> st(1);
        PCPX_CUDA(cudaMemsetAsync(st.get(), 0, sizeof(BfsState), s));
        root_from_rows_kernel<<<grid_of(n), kB, 0, s>>>(pts.d, pts.stride_f, (uint32_t)n, st.get());
        uint32_t launches = 1;
        BfsState const h  = run_search(s, n, k, d_nbr, edge_order == PCPX_EDGES_FURTHEST_FIRST,
                                       d_nrm, st.get(), launches);
        if (!direct)
            PCPX_CUDA(cudaMemcpyAsync(normals, d_nrm, 12 * n, cudaMemcpyDeviceToHost, s));
        PCPX_CUDA(cudaStreamSynchronize(s));
        if (out_levels)
            *out_levels = h.levels;
        if (out_reached)
            *out_reached = h.reached;
    });
}

} // extern "C"
