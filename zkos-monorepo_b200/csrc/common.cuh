// Shared host-side plumbing for libzkgpu: error type, CUDA checks, RAII device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <atomic>

namespace zk {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
enum { ZK_OK = 0, ZK_ERR_CUDA = -1, ZK_ERR_ARG = -2, ZK_ERR_STATE = -3, ZK_ERR_INTERNAL = -4 };

#define ZK_CUDA(x)                                                                                   \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess)                                                                       \
            throw zk::Error(zk::ZK_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " @" + \
                                                 __FILE__ + ":" + std::to_string(__LINE__));         \
    } while (0)
#define ZK_REQUIRE(c, msg) \
    do { if (!(c)) throw zk::Error(zk::ZK_ERR_ARG, std::string(msg)); } while (0)

extern std::atomic<uint64_t> g_launches;  // kernels launched by this library (zkgpu_launch_count)
#define ZK_LAUNCH(kern, grid, block, smem, st, ...)        \
    do {                                                   \
        kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__); \
        ++zk::g_launches;                                  \
        ZK_CUDA(cudaGetLastError());                       \
    } while (0)

// Optional per-kernel-class device timing (zkgpu_kernel_timing): CUDA events recorded on the launching stream
// around the launches of one class; read back with zkgpu_kernel_times.  Off by default (no events recorded).
enum { KT_MSM_BUCKETS = 0, KT_MSM_SORT = 1, KT_MSM_REDUCE = 2, KT_NTT = 3, KT_EVAL_H = 4, KT_PERM = 5, KT_POLY = 6, KT_LOOKUP = 7, KT_MISC = 8, KT_HOSTGAP = 9, KT_SLOTS = 10 };
extern bool g_ktime_on;
void ktime_begin(int slot, cudaStream_t st);
void ktime_end(int slot, cudaStream_t st);
struct KtScope {
    int slot; cudaStream_t st;
    KtScope(int s, cudaStream_t t) : slot(s), st(t) { if (g_ktime_on) ktime_begin(slot, st); }
    ~KtScope() { if (g_ktime_on) ktime_end(slot, st); }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        if (count) ZK_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};

static inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Kernel attributes (cudaFuncSetAttribute) belong to one device: run `fn` once for every CUDA device of this process a
// kernel is launched on.  The current device is the launching thread's.
struct DeviceOnce {
    std::atomic<uint64_t> done{0};
    template <class F> void run(F fn) {
        int dev = 0;
        ZK_CUDA(cudaGetDevice(&dev));
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return;
        fn();   // idempotent: two threads racing here both set the same attribute
        done.fetch_or(bit, std::memory_order_release);
    }
};

}  // namespace zk
