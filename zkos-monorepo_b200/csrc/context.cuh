// Process-wide library state: one CUDA device per process (one process per GPU), one library-owned
// stream, registered SRS tables and reusable device workspaces.  API calls serialise on `mu`.
#pragma once
#include <map>
#include <memory>
#include <mutex>
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
};

struct Context {
    std::recursive_mutex mu;
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::map<uint64_t, std::unique_ptr<SrsEntry>> srs;
    uint64_t next_handle = 1;
    MsmWorkspace ws;
    DevBuf<fr_t> fr_buf, fr_scratch;
    DevBuf<g1_affine_t> pt_buf, aff_buf;
    DevBuf<g1_xyzz_t> xyzz_buf;

    void init(int dev);
    void require();   // lazily binds device 0; throws ZK_ERR_CUDA when there is no usable GPU
    void shutdown();
    SrsEntry& get_srs(uint64_t h);
    void msm_srs_dev(SrsEntry& S, int basis, const fr_t* d_scalars, size_t n, size_t m, g1_affine_t* d_out_affine, cudaStream_t st);
};
Context& ctx();
extern thread_local std::string g_last_error;

}  // namespace zk
