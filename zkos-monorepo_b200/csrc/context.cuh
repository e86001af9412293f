// Library state.  One host process drives every CUDA device selected by zkgpu_init(device_mask)
// ("host dispatcher assigns proof i -> GPU; per-GPU host thread + streams"): a
// `Context` per device holds that device's library stream, its replica of every registered SRS and the
// scratch buffers of the primitive entry points (MSM / NTT / G1-FFT calls that a patched halo2 makes one at
// a time).  There is NO process-wide lock: primitive calls serialise on the mutex of the device they run
// on (they share its scratch buffers), the prover serialises per proving-key replica (prover.cu), and
// concurrent single-proof callers are coalesced into batches instead of queueing on a mutex.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include "common.cuh"
#include "ntt.cuh"
#include "msm.cuh"
#include "g1fft.cuh"

namespace zk {

struct SrsEntry {
    unsigned k = 0;
    size_t n = 0;
    MsmPlan plan;                   // fixed-base plan (window c, W windows)
    DevBuf<g1_affine_t> table[2];   // [0] = g, [1] = g_lagrange; W * n points each: T[w][i] = 2^(c w) * base[i]
    // latency plan: a second, narrow-window table for launch groups of a few MSMs (single proofs), where the bucket
    // reduction's dependent chain — not the number of additions — is what takes the time
    MsmPlan lat_plan;
    DevBuf<g1_affine_t> lat_tables;   // [2][W_lat][n]: g, then g_lagrange
    size_t lat_stride() const { return n * lat_plan.W; }
    // direct plan: every multiple d * 2^(8 w) * G_i (d <= 128) of both bases, so that a few MSMs are plain sums of table entries
    // (msm.cu "Direct path"; 2.1 GB per basis at n = 2^13)
    DevBuf<g1_affine_t> direct_tables;   // [2][W][n][128]
    size_t direct_stride = 0;            // points per basis
};

struct Context {
    std::recursive_mutex mu;   // serialises the primitive entry points on this device (shared scratch below)
    int device = -1;           // CUDA device index
    int slot = 0;              // position in the runtime's device list
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::map<uint64_t, std::unique_ptr<SrsEntry>> srs;   // same handle values on every device
    MsmWorkspace ws;
    DevBuf<fr_t> fr_buf, fr_scratch;
    DevBuf<g1_affine_t> pt_buf, aff_buf;
    DevBuf<g1_xyzz_t> xyzz_buf;

    void bind() const { ZK_CUDA(cudaSetDevice(device)); }
    void require() { bind(); }   // historical name: make this device current for the calling thread
    SrsEntry& get_srs(uint64_t h);
    void msm_srs_dev(SrsEntry& S, int basis, const fr_t* d_scalars, size_t n, size_t m, g1_affine_t* d_out_affine, cudaStream_t st);
    void release_all();
};

struct Runtime {
    std::mutex init_mu;
    bool inited = false;
    std::vector<std::unique_ptr<Context>> devs;   // selected devices, ascending CUDA index; devs[0] = primary
    std::shared_mutex tab_mu;                     // handle tables (SRS here, proving keys in prover.cu)
    uint64_t next_handle = 1;

    void init(int device_mask);   // idempotent for the same mask; throws ZK_ERR_CUDA without a usable GPU
    void require();               // lazily selects device 0
    void shutdown();
    Context& primary() { require(); return *devs[0]; }
    size_t count() { require(); return devs.size(); }
    Context* by_cuda_index(int dev);   // nullptr if that device was not selected
    Context& of_pointer(const void* device_ptr);   // the selected device that owns a device allocation
};
Runtime& rt();
// the primary device, made current for the calling thread (what the single-device entry points run on)
Context& ctx();
extern thread_local std::string g_last_error;

// RAII: make a device current and hold its primitive-call mutex
struct DeviceScope {
    Context& C;
    std::lock_guard<std::recursive_mutex> lk;
    explicit DeviceScope(Context& c) : C(c), lk(c.mu) { C.bind(); }
};

}  // namespace zk
