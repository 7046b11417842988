// The GPU-resident index object behind the opaque pcpx_index handle.
#pragma once
#include <mutex>
#include <vector>

#include "grid_core.cuh"
#include "host_util.hpp"

struct pcpx_index
{
    int device          = 0;
    cudaStream_t stream = nullptr;
    uint64_t n_input    = 0; // points handed in (sorted array length)
    uint64_t n_indexed  = 0; // points inside the root voxel
    float bbox_min[3]{}, bbox_max[3]{};
    uint32_t code_bits = 0;
    uint64_t n_cells   = 0;
    uint64_t cells_per_level[pcpx::kMaxLevel + 2] = {}; // occupied cells at each stored level
    pcpx::GridView grid{}; // device pointers into the buffers below
    pcpx::DevBuf<float4> pts;          // n_input entries, Morton order; w = original index
    pcpx::DevBuf<pcpx::HashSlot> table; // all levels
    pcpx::DevBuf<uint8_t> bnd;          // n_indexed entries: coarsest level at which sorted point i opens a new cell
    pcpx_timings timings{-1.f, -1.f, -1.f, -1.f, -1.f, 0u, 0u, 0u, 0u};
    std::mutex mtx; // one call at a time per index (calls serialise on `stream`)
    // tile list of one level (query.cu: ensure_tile_list), built on first use and kept:
    // tile_starts[i] = first sorted position of tile i, tile_starts[n_tiles] = n_indexed
    mutable pcpx::DevBuf<uint32_t> tile_starts, tile_count, tile_scratch;
    mutable pcpx::DevBuf<uint64_t> tile_xyz; // tile coordinates (tile_core.cuh: tile_pack)
    mutable uint32_t query_launches = 0; // kernels launched by the last kNN-shaped call
    mutable uint32_t deferred_queries = 0; // ... queries its tile pass handed to the per-thread path
    mutable uint32_t expanded_queries = 0; // ... queries that needed the retry kernel (coarser levels)
    mutable int tile_level         = -1;
    mutable uint32_t tile_capacity = 0;
    mutable uint32_t tile_region   = 0; // staged-region capacity that fits 97 % of the queries' tiles
    // the same index on further devices (pcpx_index_params.devices[1..]); owned
    std::vector<pcpx_index*> replicas;

    ~pcpx_index()
    {
        for (pcpx_index* r : replicas)
        {
            cudaSetDevice(r->device);
            delete r;
        }
        cudaSetDevice(device);
        if (stream)
            cudaStreamDestroy(stream);
    }
    size_t device_bytes() const { return pts.bytes() + table.bytes() + bnd.bytes(); }
};

namespace pcpx {

// index.cu
pcpx_index* build_index(const float* xyz, size_t n, size_t stride_bytes,
                        const pcpx_index_params* params);

// Sort helper exported by index.cu for query batches: fills `order` (device, nq entries) with
// the query permutation sorted by fine Morton code of the query positions.
void sort_queries_by_cell(const pcpx_index& ix, const float* d_queries, uint32_t stride_floats,
                          uint32_t nq, uint32_t* d_order);

} // namespace pcpx
