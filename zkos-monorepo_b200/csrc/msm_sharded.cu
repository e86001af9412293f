// Point-sharded large MSM (BASELINE.json north_star: "only the standalone large-MSM sweep splits points
// across GPUs and sums partial results over NVLink").
//
// halo2curves `best_multiexp(coeffs, bases)` over n points with the bases RESIDENT: zkgpu_bases_register splits the points
// into contiguous shards, one per selected device, and uploads each shard once (a device never holds more than its shard).
// zkgpu_msm_g1_bases sends every device its slice of the scalars, runs a full Pippenger there (msm.cu, plain mode) down to
// ONE XYZZ point, moves the G partial points (128 bytes each) to the primary device with peer copies over NVLink and adds
// them there.  That gather is the only exchange step of the whole path; nothing else crosses a device boundary.
// In the one-process-per-GPU deployment (torchrun) every rank registers only ITS shard and the 64-byte partial results are
// all-gathered by the host framework (zkgpu/multi.py) — the same per-device code, G = 1.
#include "api_util.hpp"
#include "host_util.hpp"
#include <thread>
#include <chrono>

namespace zk {

struct BasesShard {
    Context* C = nullptr;
    size_t lo = 0, hi = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr, k0 = nullptr, k1 = nullptr;
    DevBuf<g1_affine_t> pts;
    DevBuf<fr_t> sc;
    DevBuf<g1_xyzz_t> res;
    MsmWorkspace ws;
    MsmPlan plan;
    bool scalars_resident = false;
    ~BasesShard() {
        if (st) cudaStreamDestroy(st);
        for (cudaEvent_t e : {done, k0, k1}) if (e) cudaEventDestroy(e);
    }
};
struct BasesEntry {
    size_t n = 0;
    std::mutex mu;   // one MSM at a time per handle (shard workspaces)
    std::vector<std::unique_ptr<BasesShard>> shards;
    DevBuf<g1_xyzz_t> gathered;   // primary device: one partial point per shard
    DevBuf<g1_affine_t> out;
    double last_kernel_ms = 0;
};
static std::map<uint64_t, std::shared_ptr<BasesEntry>> g_bases;   // guarded by rt().tab_mu
void bases_release_all() {
    std::map<uint64_t, std::shared_ptr<BasesEntry>> drop;
    { std::unique_lock<std::shared_mutex> tl(rt().tab_mu); drop.swap(g_bases); }
}

// out[0] = sum_i in[i] (XYZZ), normalised to affine — G - 1 additions on one thread
__global__ void k_sum_partials(const g1_xyzz_t* __restrict__ in, unsigned count, g1_affine_t* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    g1_xyzz_t acc = g1_xyzz_t::identity();
    for (unsigned i = 0; i < count; ++i) {
        g1_xyzz_t p;
        p.x = fe_load(&in[i].x); p.y = fe_load(&in[i].y); p.zz = fe_load(&in[i].zz); p.zzz = fe_load(&in[i].zzz);
        acc = xyzz_add(acc, p);
    }
    g1_affine_t a = xyzz_to_affine(acc);
    fe_store(&out->x, a.x); fe_store(&out->y, a.y);
}

}  // namespace zk

using namespace zk;

extern "C" {

int zkgpu_bases_register(const uint64_t* bases_affine, size_t n, uint64_t* handle_out) {
    API_TRY
    ZK_REQUIRE(bases_affine && handle_out && n >= 1, "null pointer / empty base vector");
    ZK_REQUIRE(n <= ((size_t)1 << 27), "bases: at most 2^27 points");
    Runtime& R = rt(); R.require();
    std::shared_ptr<BasesEntry> E(new BasesEntry);
    E->n = n;
    const size_t G = std::min<size_t>(R.devs.size(), n);
    for (size_t g = 0; g < G; ++g) {
        std::unique_ptr<BasesShard> S(new BasesShard);
        S->C = R.devs[g].get();
        S->lo = n * g / G; S->hi = n * (g + 1) / G;
        S->C->bind();
        ZK_CUDA(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
        ZK_CUDA(cudaEventCreateWithFlags(&S->done, cudaEventDisableTiming));
        ZK_CUDA(cudaEventCreate(&S->k0)); ZK_CUDA(cudaEventCreate(&S->k1));
        const size_t cnt = S->hi - S->lo;
        S->pts.alloc(cnt); S->sc.alloc(cnt); S->res.alloc(1);
        ZK_CUDA(cudaMemcpyAsync(S->pts.p, bases_affine + 8 * S->lo, cnt * sizeof(g1_affine_t), cudaMemcpyHostToDevice, S->st));
        S->plan = msm_plan(cnt, false);
        E->shards.push_back(std::move(S));
    }
    for (auto& S : E->shards) { S->C->bind(); ZK_CUDA(cudaStreamSynchronize(S->st)); }
    R.devs[0]->bind();
    E->gathered.alloc(G); E->out.alloc(1);
    std::unique_lock<std::shared_mutex> tl(R.tab_mu);
    uint64_t h = R.next_handle++;
    g_bases[h] = std::move(E);
    *handle_out = h;
    API_END
}

int zkgpu_bases_release(uint64_t h) {
    API_TRY
    std::shared_ptr<BasesEntry> E;
    {
        std::unique_lock<std::shared_mutex> tl(rt().tab_mu);
        auto it = g_bases.find(h);
        ZK_REQUIRE(it != g_bases.end(), "unknown bases handle");
        E = it->second; g_bases.erase(it);
    }
    API_END
}

int zkgpu_msm_g1_bases(uint64_t h, const uint64_t* scalars, size_t n, uint64_t out_jacobian[12], double* kernel_ms) {
    API_TRY
    ZK_REQUIRE(out_jacobian, "null pointer");
    std::shared_ptr<BasesEntry> E;
    {
        std::shared_lock<std::shared_mutex> tl(rt().tab_mu);
        auto it = g_bases.find(h);
        ZK_REQUIRE(it != g_bases.end(), "unknown bases handle");
        E = it->second;
    }
    ZK_REQUIRE(n == E->n, "msm: coeffs.len() != bases.len()");   // upstream assert_eq!
    std::lock_guard<std::mutex> lk(E->mu);
    const size_t G = E->shards.size();
    Context& C0 = *rt().devs[0];
    std::vector<std::exception_ptr> err(G);
    auto one = [&](size_t g) {
        try {
            BasesShard& S = *E->shards[g];
            S.C->bind();
            const size_t cnt = S.hi - S.lo;
            if (scalars) {
                ZK_CUDA(cudaMemcpyAsync(S.sc.p, scalars + 4 * S.lo, cnt * sizeof(fr_t), cudaMemcpyHostToDevice, S.st));
                S.scalars_resident = true;
            }
            ZK_REQUIRE(S.scalars_resident, "msm: scalars == NULL but no earlier call left scalars resident");
            ZK_CUDA(cudaEventRecord(S.k0, S.st));
            msm_run(S.plan, S.sc.p, S.pts.p, 1, S.res.p, S.ws, S.st);
            ZK_CUDA(cudaEventRecord(S.k1, S.st));
            // the partial point goes to the primary device over the peer link (a plain device copy when g == 0)
            ZK_CUDA(cudaMemcpyPeerAsync(E->gathered.p + g, C0.device, S.res.p, S.C->device, sizeof(g1_xyzz_t), S.st));
            ZK_CUDA(cudaEventRecord(S.done, S.st));
        } catch (...) { err[g] = std::current_exception(); }
    };
    if (G == 1) one(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(one, g);
        for (auto& t : th) t.join();
    }
    for (auto& e : err) if (e) std::rethrow_exception(e);
    DeviceScope scope(C0);
    cudaStream_t st = C0.stream;
    for (auto& S : E->shards) ZK_CUDA(cudaStreamWaitEvent(st, S->done, 0));
    ZK_LAUNCH(k_sum_partials, 1, 32, 0, st, E->gathered.p, (unsigned)G, E->out.p);
    g1_affine_t res;
    ZK_CUDA(cudaMemcpyAsync(&res, E->out.p, sizeof res, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    double kms = 0;
    for (auto& S : E->shards) {
        S->C->bind();
        float ms = 0;
        if (cudaEventElapsedTime(&ms, S->k0, S->k1) == cudaSuccess) kms = std::max<double>(kms, ms);
    }
    C0.bind();
    E->last_kernel_ms = kms;
    if (kernel_ms) *kernel_ms = kms;
    fq_t one_q = fe_one<FqTag>();
    if (res.is_identity()) { memset(out_jacobian, 0, 96); memcpy(out_jacobian + 4, one_q.l, 32); }
    else { memcpy(out_jacobian, res.x.l, 32); memcpy(out_jacobian + 4, res.y.l, 32); memcpy(out_jacobian + 8, one_q.l, 32); }
    API_END
}

}  // extern "C"
