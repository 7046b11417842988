// Helpers shared by the translation units that implement the C ABI (api.cu, smoothing.cu):
// host <-> device staging of user buffers, per-call timing, exception -> status mapping.
#pragma once
#include <algorithm>
#include <memory>
#include <vector>

#include "query.hpp"

namespace pcpx {

// A user buffer that kernels can read: the pointer itself when it is device memory, otherwise a
// packed device copy.
struct InBuf
{
    const float* d = nullptr;
    uint32_t stride_f = 3;
    DevBuf<float> staged;
    size_t h2d_bytes = 0;

    // n rows of `width` floats, `stride_bytes` apart
    void stage(const float* p, size_t n, size_t stride_bytes, int width, cudaStream_t s)
    {
        if (!p || n == 0)
            return;
        if (is_device_pointer(p))
        {
            d = p, stride_f = (uint32_t)(stride_bytes / 4);
            return;
        }
        staged.alloc(n * (size_t)width);
        if (stride_bytes == (size_t)width * 4) // packed rows: one DMA, not one per row
            PCPX_CUDA(cudaMemcpyAsync(staged.get(), p, n * (size_t)width * 4,
                                      cudaMemcpyHostToDevice, s));
        else
            PCPX_CUDA(cudaMemcpy2DAsync(staged.get(), (size_t)width * 4, p, stride_bytes,
                                        (size_t)width * 4, n, cudaMemcpyHostToDevice, s));
        d = staged.get(), stride_f = (uint32_t)width;
        h2d_bytes = n * (size_t)width * 4;
    }
};

// A user result buffer kernels can write: in place when it is device memory, otherwise a device
// temporary copied back by finish().
template <typename T>
struct OutBuf
{
    T* user = nullptr;
    T* d    = nullptr;
    size_t n = 0;
    DevBuf<T> tmp;
    bool direct = false;

    void prepare(T* p, size_t count)
    {
        user = p, n = count;
        if (!p || count == 0)
            return;
        direct = is_device_pointer(p);
        if (direct)
            d = p;
        else
        {
            tmp.alloc(count);
            d = tmp.get();
        }
    }
    void finish(cudaStream_t s, size_t count = (size_t)-1)
    {
        if (user && !direct && n)
            PCPX_CUDA(cudaMemcpyAsync(user, d, std::min(n, count) * sizeof(T),
                                      cudaMemcpyDeviceToHost, s));
    }
};

struct CallTimer
{
    pcpx_index& ix;
    Event t0, k0, k1, t1;
    explicit CallTimer(pcpx_index& i) : ix(i)
    {
        ix.timings.query_sort_ms   = 0.f;
        ix.timings.kernel_launches = 0;
        ix.timings.retry_queries   = 0;
        t0.record(ix.qstream());
    }
    void kernel_begin() { k0.record(ix.qstream()); }
    void kernel_end() { k1.record(ix.qstream()); }
    void done()
    {
        t1.record(ix.qstream());
        PCPX_CUDA(cudaStreamSynchronize(ix.qstream()));
        ix.timings.kernel_ms = elapsed_ms(k0, k1);
        ix.timings.total_ms  = elapsed_ms(t0, t1);
    }
};

// queries + their processing order
struct Batch
{
    InBuf in;
    DevBuf<uint32_t> order;
    QueryBatch qb{nullptr, 3u, nullptr, 0u};

    void prepare(pcpx_index& ix, const float* queries, size_t nq, size_t stride_bytes)
    {
        if (nq >= 0xFFFFFFFFull)
            fail(PCPX_ERR_UNSUPPORTED, "more than 2^32 - 2 queries in one call");
        if (!queries)
        {
            if (nq != ix.n_input)
                fail(PCPX_ERR_INVALID_ARG,
                     "queries == NULL means the indexed cloud itself: nq must be %llu, got %llu",
                     (unsigned long long)ix.n_input, (unsigned long long)nq);
            qb = QueryBatch{nullptr, 3u, nullptr, (uint32_t)nq};
            return;
        }
        if (stride_bytes == 0)
            stride_bytes = 12;
        if (stride_bytes < 12 || stride_bytes % 4)
            fail(PCPX_ERR_INVALID_ARG, "query_stride_bytes must be a multiple of 4 and >= 12");
        in.stage(queries, nq, stride_bytes, 3, ix.qstream());
        qb = QueryBatch{in.d, in.stride_f, nullptr, (uint32_t)nq};
        if (nq > 1)
        {
            Event s0, s1;
            s0.record(ix.qstream());
            order.alloc(nq);
            sort_queries_by_cell(ix, in.d, in.stride_f, (uint32_t)nq, order.get());
            s1.record(ix.qstream());
            PCPX_CUDA(cudaStreamSynchronize(ix.qstream()));
            ix.timings.query_sort_ms = elapsed_ms(s0, s1);
            qb.order                 = order.get();
        }
    }
};

template <class F>
int guarded(F&& f)
{
    try
    {
        f();
        return PCPX_OK;
    }
    catch (Error const& e)
    {
        // the call's buffers went back to the pool while unwinding: nothing that was already
        // submitted may still be using them when another thread takes them
        cudaDeviceSynchronize();
        last_error_storage() = e.msg;
        return e.code;
    }
    catch (std::bad_alloc const&)
    {
        last_error_storage() = "host allocation failed";
        return PCPX_ERR_OUT_OF_MEMORY;
    }
    catch (std::exception const& e)
    {
        last_error_storage() = e.what();
        return PCPX_ERR_CUDA;
    }
}

inline pcpx_index& checked(const pcpx_index* ix)
{
    if (!ix)
        fail(PCPX_ERR_INVALID_ARG, "index is NULL");
    return *const_cast<pcpx_index*>(ix);
}


} // namespace pcpx
