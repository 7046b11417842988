// Host-side plumbing shared by the translation units of libpcpx.so: error transport across the
// C ABI (exceptions never leave the library), RAII device buffers, pointer classification.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

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

// Process-wide caching allocator for device memory.  cudaMalloc / cudaFree cost tens of
// microseconds to milliseconds each and synchronise the device; an index build needs a dozen
// temporaries and a serving loop rebuilds indices continuously, so freed blocks are kept per
// device and handed out again (best fit, at most 25 % slack).  Blocks are only recycled after
// the stream work that used them has been synchronised by the caller (every C-ABI call ends
// with a stream synchronise before its buffers go out of scope).
class DevicePool
{
  public:
    static DevicePool& instance()
    {
        static DevicePool p;
        return p;
    }
    void* acquire(size_t bytes)
    {
        if (bytes == 0)
            return nullptr;
        int dev = 0;
        PCPX_CUDA(cudaGetDevice(&dev));
        size_t const want = (bytes + 511) & ~size_t(511);
        {
            std::lock_guard<std::mutex> lock(m_);
            auto& fl  = free_[dev];
            auto best = fl.end();
            for (auto it = fl.lower_bound(want); it != fl.end(); ++it)
            {
                if (it->first <= want + want / 4 + 4096)
                    best = it;
                break;
            }
            if (best != fl.end())
            {
                void* p = best->second;
                size_[p] = best->first;
                cached_[dev] -= best->first;
                fl.erase(best);
                return p;
            }
        }
        void* p       = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaErrorMemoryAllocation)
        {
            cudaGetLastError();
            trim(dev); // give cached blocks back and retry once
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess)
            fail(e == cudaErrorMemoryAllocation ? PCPX_ERR_OUT_OF_MEMORY : PCPX_ERR_CUDA,
                 "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        std::lock_guard<std::mutex> lock(m_);
        size_[p] = want;
        dev_[p]  = dev;
        return p;
    }
    // A freed block is kept for reuse only while the device's cache stays under its cap
    // (default: one eighth of the device's memory, at least 1 GiB — a step over a 100 M-point cloud
    // re-uses ~6 GB of blocks, and a cap below the working set turns every step into
    // cudaFree + cudaMalloc, both device-synchronising; PCPX_POOL_CAP_MB /
    // pcpx_set_tuning("pool_cap_mb") set it explicitly): the smallest cached blocks are returned
    // to the driver until the newcomer fits, and a block larger than the cap goes straight back.
    void release(void* p)
    {
        if (!p)
            return;
        std::multimap<size_t, void*> drop;
        {
            std::lock_guard<std::mutex> lock(m_);
            auto it = size_.find(p);
            if (it == size_.end())
            {
                drop.emplace(0, p);
            }
            else
            {
                int const dev      = dev_[p];
                size_t const bytes = it->second;
                auto& fl           = free_[dev];
                size_t& cached     = cached_[dev];
                size_t const cap_  = cap_for(dev);
                if (bytes > cap_)
                {
                    size_.erase(p), dev_.erase(p);
                    drop.emplace(bytes, p);
                }
                else
                {
                    while (cached + bytes > cap_ && !fl.empty())
                    {
                        auto victim = fl.begin(); // smallest first: large blocks are the costly ones to re-allocate
                        cached -= victim->first;
                        size_.erase(victim->second), dev_.erase(victim->second);
                        drop.insert(*victim);
                        fl.erase(victim);
                    }
                    fl.emplace(bytes, p);
                    cached += bytes;
                }
            }
        }
        for (auto& b : drop)
            cudaFree(b.second);
    }
    void set_cap(size_t bytes)
    {
        std::lock_guard<std::mutex> lock(m_);
        cap_ = bytes;
    }
    size_t cached_bytes(int dev)
    {
        std::lock_guard<std::mutex> lock(m_);
        return cached_[dev];
    }
    // hand a block over to the caller (it will be freed with cudaFree by pcpx_free)
    void forget(void* p)
    {
        std::lock_guard<std::mutex> lock(m_);
        size_.erase(p);
        dev_.erase(p);
    }
    void trim(int dev)
    {
        std::multimap<size_t, void*> blocks;
        {
            std::lock_guard<std::mutex> lock(m_);
            blocks.swap(free_[dev]);
            cached_[dev] = 0;
            for (auto& b : blocks)
                size_.erase(b.second), dev_.erase(b.second);
        }
        for (auto& b : blocks)
            cudaFree(b.second);
    }

  private:
    DevicePool()
    {
        if (const char* e = std::getenv("PCPX_POOL_CAP_MB"))
            cap_ = (size_t)std::strtoull(e, nullptr, 10) << 20;
    }
    // (m_ held) explicit cap, or one eighth of the device's memory
    size_t cap_for(int dev)
    {
        if (cap_ != 0)
            return cap_;
        auto it = auto_cap_.find(dev);
        if (it != auto_cap_.end())
            return it->second;
        size_t free_b = 0, total_b = 0;
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != dev)
            cudaSetDevice(dev);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
            total_b = 0, cudaGetLastError();
        if (cur != dev)
            cudaSetDevice(cur);
        size_t const cap = std::max<size_t>((size_t)1 << 30, total_b / 8);
        auto_cap_[dev]   = cap;
        return cap;
    }
    std::mutex m_;
    size_t cap_ = 0; // 0: automatic
    std::map<int, size_t> auto_cap_;
    std::map<int, size_t> cached_;
    std::map<int, std::multimap<size_t, void*>> free_;
    std::map<void*, size_t> size_;
    std::map<void*, int> dev_;
};

// Owning device allocation out of the pool.
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
            p_ = static_cast<T*>(DevicePool::instance().acquire(n * sizeof(T)));
    }
    void release()
    {
        if (p_)
            DevicePool::instance().release(p_);
        p_ = nullptr, n_ = 0;
    }
    T* get() const { return p_; }
    size_t size() const { return n_; }
    size_t bytes() const { return n_ * sizeof(T); }
    // ownership leaves the pool: the caller frees it with cudaFree
    T* detach()
    {
        T* p = p_;
        if (p)
            DevicePool::instance().forget(p);
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

struct Stream
{
    cudaStream_t s = nullptr;
    Stream() { PCPX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
    ~Stream()
    {
        if (s)
            cudaStreamDestroy(s);
    }
    Stream(const Stream&)            = delete;
    Stream& operator=(const Stream&) = delete;
};

// ---- small device -> host read-backs ---------------------------------------------------------
// Counters, histograms and the bounding-box partials the host needs in the middle of a call do
// NOT go through the copy engines: a small DMA queued behind ANOTHER stream's bulk transfer
// waits for all of it (measured: two clouds streamed through the C ABI from two host threads
// overlapped almost nothing, 6.7 ms per cloud against 8.0 serial, because each of the five
// read-backs of a step could sit behind the other cloud's 2.3 ms result copy).  A one-block
// kernel stores the words into mapped pinned host memory instead; the stream is synchronised
// and the host reads them there.
#ifdef __CUDACC__
static __global__ void mailbox_store_kernel(const uint32_t* __restrict__ src,
                                            uint32_t* __restrict__ dst, uint32_t nwords)
{
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x)
        dst[i] = src[i];
}

// Mailboxes are pinned allocations (cudaHostAlloc maps them into every context: milliseconds, and
// serialised across devices), so they are pooled for the life of the process: a call borrows one
// and gives it back — the replica threads of a multi-device call are short-lived, a
// thread_local mailbox cost 24 ms per call on eight devices.
class HostMailbox
{
  public:
    static constexpr size_t kBytes = 64 * 1024;
    HostMailbox()
    {
        {
            std::lock_guard<std::mutex> lock(mutex());
            if (!spare().empty())
            {
                p_ = spare().back();
                spare().pop_back();
            }
        }
        if (!p_)
            PCPX_CUDA(cudaHostAlloc(&p_, kBytes, cudaHostAllocMapped | cudaHostAllocPortable));
    }
    ~HostMailbox()
    {
        std::lock_guard<std::mutex> lock(mutex());
        spare().push_back(p_);
    }
    HostMailbox(const HostMailbox&)            = delete;
    HostMailbox& operator=(const HostMailbox&) = delete;
    uint32_t* words() { return static_cast<uint32_t*>(p_); }

  private:
    static std::mutex& mutex()
    {
        static std::mutex m;
        return m;
    }
    static std::vector<void*>& spare()
    {
        static std::vector<void*>* v = new std::vector<void*>(); // (never destroyed: no teardown order issues)
        return *v;
    }
    void* p_ = nullptr;
};

// Device-resident index arrays handed in by a caller (neighbour rows, start sets) are checked on
// the device before any kernel dereferences them: *bad counts the entries that are neither `pad`
// nor < limit.
static __global__ void count_bad_indices_kernel(const uint32_t* __restrict__ idx, size_t count,
                                                uint32_t limit, uint32_t pad,
                                                uint32_t* __restrict__ bad)
{
    uint32_t mine = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x)
    {
        uint32_t const v = idx[i];
        mine += (v != pad && v >= limit) ? 1u : 0u;
    }
    if (mine)
        atomicAdd(bad, mine);
}

// dst (host) <- src (device), `bytes` a multiple of 4; returns with the stream synchronised
inline void read_back(cudaStream_t s, void* dst, const void* src_dev, size_t bytes)
{
    if (bytes == 0)
    {
        PCPX_CUDA(cudaStreamSynchronize(s));
        return;
    }
    if (bytes > HostMailbox::kBytes || bytes % 4)
    {
        PCPX_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, s));
        PCPX_CUDA(cudaStreamSynchronize(s));
        return;
    }
    HostMailbox mailbox;
    uint32_t* box = mailbox.words();
    mailbox_store_kernel<<<1, 256, 0, s>>>(static_cast<const uint32_t*>(src_dev), box,
                                           (uint32_t)(bytes / 4));
    PCPX_CHECK_LAUNCH();
    PCPX_CUDA(cudaStreamSynchronize(s));
    std::memcpy(dst, box, bytes);
}
#endif

#ifdef __CUDACC__
// number of entries of the device array idx[0, count) that are neither `pad` nor < limit
inline uint32_t count_bad_indices(cudaStream_t s, const uint32_t* d_idx, size_t count, uint32_t limit,
                                  uint32_t pad)
{
    if (count == 0)
        return 0u;
    DevBuf<uint32_t> bad(1);
    PCPX_CUDA(cudaMemsetAsync(bad.get(), 0, 4, s));
    unsigned const blocks = (unsigned)std::min<size_t>((count + 255) / 256, 148u * 8u);
    count_bad_indices_kernel<<<blocks, 256, 0, s>>>(d_idx, count, limit, pad, bad.get());
    PCPX_CHECK_LAUNCH();
    uint32_t h = 0;
    read_back(s, &h, bad.get(), 4);
    return h;
}
#endif

inline float elapsed_ms(const Event& a, const Event& b)
{
    float ms = 0.f;
    PCPX_CUDA(cudaEventElapsedTime(&ms, a.e, b.e));
    return ms;
}

} // namespace pcpx
