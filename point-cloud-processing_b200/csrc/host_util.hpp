// Host-side plumbing shared by the translation units of libpcpx.so: error transport across the
// C ABI (exceptions never leave the library), RAII device buffers, pointer classification.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <utility>

#include "pcpx.h"

namespace pcpx {

struct Error
{
    int code;
    std::string msg;
};

inline std::string& last_error_storage()
{
    static thread_local std::string s;
    return s;
}

[[noreturn]] inline void fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error{code, buf};
}

#define PCPX_CUDA(call)                                                                        \
    do                                                                                         \
    {                                                                                          \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            ::pcpx::fail(                                                                      \
                e_ == cudaErrorMemoryAllocation ? PCPX_ERR_OUT_OF_MEMORY : PCPX_ERR_CUDA,      \
                "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

#define PCPX_CHECK_LAUNCH() PCPX_CUDA(cudaGetLastError())

// Owning device allocation (stream-ordered free is not needed: the index outlives its calls).
template <typename T>
class DevBuf
{
  public:
    DevBuf() = default;
    explicit DevBuf(size_t n) { alloc(n); }
    DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr, o.n_ = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        if (this != &o)
        {
            release();
            p_ = o.p_, n_ = o.n_;
            o.p_ = nullptr, o.n_ = 0;
        }
        return *this;
    }
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }

    void alloc(size_t n)
    {
        release();
        n_ = n;
        if (n)
            PCPX_CUDA(cudaMalloc(reinterpret_cast<void**>(&p_), n * sizeof(T)));
    }
    void release()
    {
        if (p_)
            cudaFree(p_);
        p_ = nullptr, n_ = 0;
    }
    T* get() const { return p_; }
    size_t size() const { return n_; }
    size_t bytes() const { return n_ * sizeof(T); }
    T* detach()
    {
        T* p = p_;
        p_ = nullptr, n_ = 0;
        return p;
    }

  private:
    T* p_     = nullptr;
    size_t n_ = 0;
};

// true when `p` is memory a kernel on `device` can dereference at full speed (device or managed)
inline bool is_device_pointer(const void* p)
{
    if (!p)
        return false;
    cudaPointerAttributes a{};
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess)
    {
        cudaGetLastError(); // plain malloc memory on old drivers: clear the sticky error
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct ScopedDevice
{
    int prev = -1;
    explicit ScopedDevice(int dev)
    {
        PCPX_CUDA(cudaGetDevice(&prev));
        if (dev != prev)
            PCPX_CUDA(cudaSetDevice(dev));
        else
            prev = -1;
    }
    ~ScopedDevice()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

struct Event
{
    cudaEvent_t e = nullptr;
    Event() { PCPX_CUDA(cudaEventCreate(&e)); }
    ~Event()
    {
        if (e)
            cudaEventDestroy(e);
    }
    Event(const Event&)            = delete;
    Event& operator=(const Event&) = delete;
    void record(cudaStream_t s) { PCPX_CUDA(cudaEventRecord(e, s)); }
};

inline float elapsed_ms(const Event& a, const Event& b)
{
    float ms = 0.f;
    PCPX_CUDA(cudaEventElapsedTime(&ms, a.e, b.e));
    return ms;
}

} // namespace pcpx
