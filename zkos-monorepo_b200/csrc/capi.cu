// libzkgpu C ABI (include/zkgpu.h): host-buffer entry points that a patched halo2curves /
// halo2_proofs would call in place of best_multiexp / best_fft / EvaluationDomain transforms, plus
// device-resident variants for the prover pipeline and the benchmark.  No CPU fallback anywhere: if
// CUDA is unavailable every compute call returns ZKGPU_ERR_CUDA.
#include "api_util.hpp"
#include "prover_kernels.cuh"
#include "host_util.hpp"

namespace zk {
void prover_release_all();  // prover.cu
void bases_release_all();   // msm_sharded.cu
std::atomic<uint64_t> g_launches{0};
Runtime& rt() { static Runtime r; return r; }
Context& ctx() { Context& c = rt().primary(); c.bind(); return c; }
thread_local std::string g_last_error;

// ---- per-kernel-class timing ------------------------------------------------------------------
// Event pools are per CUDA device (an event may only be recorded on a stream of the device it was created on);
// zkgpu_kernel_times sums a class over the devices.
bool g_ktime_on = false;
namespace {
struct KtSlot { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; std::vector<cudaEvent_t> free_ev; double ms = 0; uint64_t launches = 0; };
const int KT_MAX_DEV = 16;
KtSlot g_kt[KT_MAX_DEV][KT_SLOTS];
std::mutex g_kt_mu;
int kt_dev() { int d = 0; cudaGetDevice(&d); return d < KT_MAX_DEV ? d : KT_MAX_DEV - 1; }
cudaEvent_t kt_event(KtSlot& s) {
    if (!s.free_ev.empty()) { cudaEvent_t e = s.free_ev.back(); s.free_ev.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void kt_collect(KtSlot& s) {
    for (auto& pr : s.pending) {
        if (cudaEventSynchronize(pr.second) == cudaSuccess) {
            float ms = 0; if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { s.ms += ms; s.launches++; }
        }
        s.free_ev.push_back(pr.first); s.free_ev.push_back(pr.second);
    }
    s.pending.clear();
}
}  // namespace
void ktime_begin(int slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    KtSlot& s = g_kt[kt_dev()][slot];
    cudaEvent_t a = kt_event(s), b = kt_event(s);
    cudaEventRecord(a, st);
    s.pending.push_back({a, b});
}
void ktime_end(int slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    KtSlot& s = g_kt[kt_dev()][slot];
    if (!s.pending.empty()) cudaEventRecord(s.pending.back().second, st);
    if (s.pending.size() > 4096) kt_collect(s);
}

// zkgpu_init(device_mask): bit i selects CUDA device i; 0 selects every visible device.
void Runtime::init(int device_mask) {
    std::lock_guard<std::mutex> lk(init_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(ZK_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libzkgpu has no CPU fallback)");
    if (count > 30) count = 30;
    unsigned mask = device_mask == 0 ? ((1u << count) - 1) : (unsigned)device_mask;
    ZK_REQUIRE(device_mask >= 0 && (mask >> count) == 0, "zkgpu_init: device_mask selects a device that does not exist");
    if (inited) {
        unsigned cur = 0;
        for (auto& d : devs) cur |= 1u << d->device;
        ZK_REQUIRE(cur == mask, "zkgpu_init: already initialised with another device mask (zkgpu_shutdown first)");
        return;
    }
    std::vector<std::unique_ptr<Context>> sel;
    for (int dev = 0; dev < count; ++dev) {
        if (!((mask >> dev) & 1)) continue;
        ZK_CUDA(cudaSetDevice(dev));
        cudaDeviceProp prop;
        ZK_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            throw Error(ZK_ERR_CUDA, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                         "); libzkgpu is built for sm_100a only");
        std::unique_ptr<Context> c(new Context);
        c->device = dev; c->slot = (int)sel.size(); c->sm_count = prop.multiProcessorCount;
        ZK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        sel.push_back(std::move(c));
    }
    // peer access between the selected devices: partial results of a point-sharded MSM are summed over NVLink
    for (auto& a : sel)
        for (auto& b : sel) {
            if (a->device == b->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a->device, b->device) == cudaSuccess && can) {
                cudaSetDevice(a->device);
                cudaError_t pe = cudaDeviceEnablePeerAccess(b->device, 0);
                if (pe != cudaSuccess) cudaGetLastError();   // already enabled by the host application: fine
            }
        }
    ZK_CUDA(cudaSetDevice(sel[0]->device));
    devs = std::move(sel);
    inited = true;
}
void Runtime::require() {
    if (!inited) init(1);
}
Context* Runtime::by_cuda_index(int dev) {
    require();
    for (auto& d : devs) if (d->device == dev) return d.get();
    return nullptr;
}
Context& Runtime::of_pointer(const void* device_ptr) {
    require();
    cudaPointerAttributes at;
    ZK_CUDA(cudaPointerGetAttributes(&at, device_ptr));
    ZK_REQUIRE(at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged, "expected a device pointer");
    Context* c = by_cuda_index(at.device);
    ZK_REQUIRE(c != nullptr, "device pointer belongs to a CUDA device that zkgpu_init did not select");
    return *c;
}
void Context::release_all() {
    std::lock_guard<std::recursive_mutex> lk(mu);
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    srs.clear();
    ws = MsmWorkspace();
    fr_buf.release(); fr_scratch.release(); pt_buf.release(); xyzz_buf.release(); aff_buf.release();
    cudaStreamDestroy(stream);
    stream = nullptr;
}
void Runtime::shutdown() {
    std::lock_guard<std::mutex> lk(init_mu);
    if (!inited) return;
    prover_release_all();
    bases_release_all();
    {
        std::unique_lock<std::shared_mutex> tl(tab_mu);
        for (auto& d : devs) d->release_all();
    }
    ntt_clear_cache();
    devs.clear();
    inited = false;
}

SrsEntry& Context::get_srs(uint64_t h) {
    auto it = srs.find(h);
    ZK_REQUIRE(it != srs.end(), "unknown SRS handle");
    return *it->second;
}

// out[m] affine = MSM over the SRS basis for m scalar vectors already on the device
void Context::msm_srs_dev(SrsEntry& S, int basis, const fr_t* d_scalars, size_t n, size_t m, g1_affine_t* d_out_affine, cudaStream_t st) {
    ZK_REQUIRE(basis == 0 || basis == 1, "basis must be 0 (g) or 1 (g_lagrange)");
    ZK_REQUIRE(n >= 1 && n <= S.n, "msm: n exceeds the registered SRS size");
    if (S.lat_tables.p && m <= ZK_LAT_MAX_M) {
        // one or a few commitments per call (what `ParamsKZG::commit{,_lagrange}` issues one at a time): the latency path
        MsmPlan lp = S.lat_plan;
        lp.n = n; lp.tstride = S.n;
        const fr_t* sc[ZK_LAT_MAX_M];
        for (size_t i = 0; i < m; ++i) sc[i] = d_scalars + i * n;
        if (S.direct_tables.p) msm_direct_run(sc, basis ? ~0u : 0u, S.direct_stride, S.direct_tables.p, n, S.n, m, d_out_affine, ws, st);
        else msm_lat_run(lp, sc, basis ? ~0u : 0u, S.lat_stride(), S.lat_tables.p, m, d_out_affine, ws, st);
        return;
    }
    MsmPlan plan = S.plan;
    plan.n = n; plan.tstride = S.n;
    const size_t chunk = 1024;
    xyzz_buf.ensure(m < chunk ? m : chunk);
    for (size_t off = 0; off < m; off += chunk) {
        size_t cnt = m - off < chunk ? m - off : chunk;
        msm_run(plan, d_scalars + off * n, S.table[basis].p, cnt, xyzz_buf.p, ws, st);
        g1_normalize(xyzz_buf.p, d_out_affine + off, cnt, st);
    }
}

static void jacobian_out(const g1_affine_t& a, uint64_t out[12]) {
    fq_t one = fe_one<FqTag>();
    if (a.is_identity()) {
        memset(out, 0, 96);
        memcpy(out + 4, one.l, 32);
        return;
    }
    memcpy(out, a.x.l, 32); memcpy(out + 4, a.y.l, 32); memcpy(out + 8, one.l, 32);
}

}  // namespace zk

using namespace zk;


extern "C" {

int zkgpu_abi_version(void) { return 3; }
/* Keccak-256 (EVM padding): the hash of the Keccak256 transcript and of the contract-side `commitment` public inputs; host only */
int zkgpu_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    if ((!data && len) || !out) { g_last_error = "null pointer"; return ZKGPU_ERR_ARG; }
    keccak256(data ? data : reinterpret_cast<const uint8_t*>(""), len, out);
    return ZKGPU_OK;
}
const char* zkgpu_last_error(void) { return g_last_error.c_str(); }
uint64_t zkgpu_launch_count(void) { return g_launches.load(); }

void zkgpu_kernel_timing(int enable) { g_ktime_on = enable != 0; }
int zkgpu_msm_additions(uint64_t* total, int reset) {
    API_TRY
    Runtime& R = rt(); R.require();
    uint64_t sum = 0;
    for (auto& dp : R.devs) { DeviceScope scope(*dp); ZK_CUDA(cudaDeviceSynchronize()); sum += msm_entries_counter(reset != 0); }
    if (total) *total = sum;
    API_END
}
int zkgpu_kernel_times(int slot, double* total_ms, uint64_t* launches, int reset) {
    if (slot < 0 || slot >= KT_SLOTS) return ZKGPU_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_kt_mu);
    double ms = 0; uint64_t cnt = 0;
    int cur = 0; cudaGetDevice(&cur);
    for (int d = 0; d < KT_MAX_DEV; ++d) {
        KtSlot& s = g_kt[d][slot];
        if (!s.pending.empty()) { cudaSetDevice(d); kt_collect(s); }
        ms += s.ms; cnt += s.launches;
        if (reset) { s.ms = 0; s.launches = 0; }
    }
    cudaSetDevice(cur);
    if (total_ms) *total_ms = ms;
    if (launches) *launches = cnt;
    return ZKGPU_OK;
}
void* zkgpu_stream(void) { return rt().inited ? (void*)rt().devs[0]->stream : nullptr; }

int zkgpu_init(int device_mask) {
    API_TRY
    rt().init(device_mask);
    API_END
}
void zkgpu_shutdown(void) {
    try { rt().shutdown(); } catch (...) {}
}
int zkgpu_device_count(void) { return rt().inited ? (int)rt().devs.size() : 0; }
int zkgpu_device_index(int slot) { return rt().inited && slot >= 0 && slot < (int)rt().devs.size() ? rt().devs[slot]->device : -1; }

int zkgpu_msm_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_jacobian[12]) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(scalars && bases && out_jacobian, "null pointer");
    g1_affine_t res;
    if (n == 0) { res.x = fq_t::zero(); res.y = fq_t::zero(); jacobian_out(res, out_jacobian); return ZKGPU_OK; }
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(n); C.pt_buf.ensure(n); C.xyzz_buf.ensure(1); C.aff_buf.ensure(1);
    ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, bases, n * 64, cudaMemcpyHostToDevice, st));
    MsmPlan plan = msm_plan(n, false);
    msm_run(plan, C.fr_buf.p, C.pt_buf.p, 1, C.xyzz_buf.p, C.ws, st);
    g1_normalize(C.xyzz_buf.p, C.aff_buf.p, 1, st);
    ZK_CUDA(cudaMemcpyAsync(&res, C.aff_buf.p, 64, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    jacobian_out(res, out_jacobian);
    API_END
}

int zkgpu_srs_register(const uint64_t* g, const uint64_t* g_lagrange, uint32_t k, uint64_t* handle_out) {
    API_TRY
    ZK_REQUIRE(g && g_lagrange && handle_out, "null pointer");
    ZK_REQUIRE(k >= 1 && k <= 20, "srs: k out of range");
    Runtime& R = rt(); R.require();
    uint64_t h;
    { std::unique_lock<std::shared_mutex> tl(R.tab_mu); h = R.next_handle++; }
    // one replica per selected device (the SRS is <= 1 MiB at k = 13, its window tables ~10 MiB per basis)
    for (auto& dp : R.devs) {
        DeviceScope scope(*dp);
        Context& C = *dp;
        std::unique_ptr<SrsEntry> S(new SrsEntry);
        S->k = k; S->n = (size_t)1 << k;
        S->plan = msm_plan(S->n, true);
        S->plan.tstride = S->n;
        cudaStream_t st = C.stream;
        C.pt_buf.ensure(S->n);
        const uint64_t* src[2] = {g, g_lagrange};
        // second half of each window table: the same windows over the prefix sums of the basis (MsmPlan::diff_offset)
        const bool with_diff = msm_diff_enabled() && k <= 16 && 2 * S->n * S->plan.W < (1ull << 31);
        const size_t half = S->n * S->plan.W;
        DevBuf<g1_affine_t> prefix;
        if (with_diff) prefix.alloc(S->n);
        for (int b = 0; b < 2; ++b) {
            S->table[b].alloc(with_diff ? 2 * half : half);
            ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, src[b], S->n * 64, cudaMemcpyHostToDevice, st));
            msm_precompute_table(S->plan, C.pt_buf.p, S->table[b].p, st);
            if (with_diff) {
                msm_prefix_bases(C.pt_buf.p, S->n, prefix.p, st);
                msm_precompute_table(S->plan, prefix.p, S->table[b].p + half, st);
            }
        }
        if (with_diff) { ZK_CUDA(cudaStreamSynchronize(st)); S->plan.diff_offset = half; }
        // narrow-window tables of both bases, back to back, for the latency path (a few MSMs per launch group: single proofs)
        if (msm_lat_window()) {
            S->lat_plan = msm_plan(S->n, true, msm_lat_window());
            S->lat_plan.tstride = S->n;
            S->lat_tables.alloc(2 * S->n * S->lat_plan.W);
            for (int b = 0; b < 2; ++b) {
                ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, src[b], S->n * 64, cudaMemcpyHostToDevice, st));
                msm_precompute_table(S->lat_plan, C.pt_buf.p, S->lat_tables.p + (size_t)b * S->n * S->lat_plan.W, st);
            }
            ZK_CUDA(cudaStreamSynchronize(st));
            // all multiples for the direct path, if they fit comfortably (a quarter of the free HBM at most)
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const size_t stride = msm_direct_points_per_basis(S->n);
            if (msm_direct_enabled() && k <= 15 && 2 * stride * sizeof(g1_affine_t) <= free_b / 4) {
                S->direct_tables.alloc(2 * stride);
                S->direct_stride = stride;
                for (int b = 0; b < 2; ++b) msm_direct_build(S->lat_tables.p + (size_t)b * S->lat_stride(), S->n, S->direct_tables.p + (size_t)b * stride, st);
            }
        }
        std::unique_lock<std::shared_mutex> tl(R.tab_mu);
        C.srs[h] = std::move(S);
    }
    *handle_out = h;
    API_END
}
int zkgpu_srs_release(uint64_t h) {
    API_TRY
    Runtime& R = rt(); R.require();
    size_t erased = 0;
    for (auto& dp : R.devs) { DeviceScope scope(*dp); std::unique_lock<std::shared_mutex> tl(R.tab_mu); erased += dp->srs.erase(h); }
    ZK_REQUIRE(erased > 0, "unknown SRS handle");
    API_END
}

int zkgpu_msm_g1_srs_batch(uint64_t srs, int basis, const uint64_t* scalars, size_t n, size_t m, uint64_t* out_affine) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(scalars && out_affine, "null pointer");
    SrsEntry& S = C.get_srs(srs);
    if (m == 0) return ZKGPU_OK;
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(n * m); C.aff_buf.ensure(m);
    ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, scalars, n * m * 32, cudaMemcpyHostToDevice, st));
    C.msm_srs_dev(S, basis, C.fr_buf.p, n, m, C.aff_buf.p, st);
    ZK_CUDA(cudaMemcpyAsync(out_affine, C.aff_buf.p, m * 64, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}
int zkgpu_msm_g1_srs(uint64_t srs, int basis, const uint64_t* scalars, size_t n, uint64_t out_jacobian[12]) {
    g1_affine_t res;
    int rc = zkgpu_msm_g1_srs_batch(srs, basis, scalars, n, 1, reinterpret_cast<uint64_t*>(&res));
    if (rc) return rc;
    jacobian_out(res, out_jacobian);
    return ZKGPU_OK;
}
int zkgpu_msm_g1_srs_batch_dev(uint64_t srs, int basis, const void* d_scalars, size_t n, size_t m, void* d_out_affine, void* stream) {
    API_TRY
    ZK_REQUIRE(d_scalars && d_out_affine, "null pointer");
    DeviceScope scope(rt().of_pointer(d_scalars));
    Context& C = scope.C;
    SrsEntry& S = C.get_srs(srs);
    cudaStream_t st = stream ? (cudaStream_t)stream : C.stream;
    C.msm_srs_dev(S, basis, (const fr_t*)d_scalars, n, m, (g1_affine_t*)d_out_affine, st);
    // the device's shared MSM workspace is reused by the next call (possibly on another stream): finish before the lock drops
    ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}

static void ntt_host(uint64_t* a, const fr_t& omega, uint32_t log_n, size_t m, bool scale, const fr_t& factor) {
    Context& C = ctx(); C.require();
    ZK_REQUIRE(a, "null pointer");
    ZK_REQUIRE(log_n <= 24, "ntt: log_n > 24 unsupported");
    size_t n = (size_t)1 << log_n;
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(n * m);
    C.fr_scratch.ensure(ntt_scratch_elems(log_n, m));
    ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, a, n * m * 32, cudaMemcpyHostToDevice, st));
    NttJob J;
    J.in = C.fr_buf.p; J.out = C.fr_buf.p; J.scratch = C.fr_scratch.p; J.batch = m; J.log_n = log_n; J.omega = omega;
    J.has_scale = scale; J.scale = factor;
    ntt_run(J, st);
    ZK_CUDA(cudaMemcpyAsync(a, C.fr_buf.p, n * m * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
}

int zkgpu_ntt_fr_batch(uint64_t* a, const uint64_t omega[4], uint32_t log_n, size_t m) {
    API_BEGIN
    ZK_REQUIRE(omega, "null pointer");
    fr_t w; memcpy(w.l, omega, 32);
    if (m) ntt_host(a, w, log_n, m, false, w);
    API_END
}
int zkgpu_ntt_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n) { return zkgpu_ntt_fr_batch(a, omega, log_n, 1); }

int zkgpu_ntt_fr_batch_dev(void* d_a, const uint64_t omega[4], uint32_t log_n, size_t m, void* d_scratch, void* stream) {
    API_TRY
    ZK_REQUIRE(d_a && omega, "null pointer");
    DeviceScope scope(rt().of_pointer(d_a));
    Context& C = scope.C;
    fr_t w; memcpy(w.l, omega, 32);
    cudaStream_t st = stream ? (cudaStream_t)stream : C.stream;
    NttJob J;
    J.in = (fr_t*)d_a; J.out = (fr_t*)d_a; J.scratch = (fr_t*)d_scratch; J.batch = m; J.log_n = log_n; J.omega = w;
    if (!J.scratch && ntt_scratch_elems(log_n, m)) {
        C.fr_scratch.ensure(ntt_scratch_elems(log_n, m));
        J.scratch = C.fr_scratch.p;
    }
    ntt_run(J, st);
    // same reason as above when the library's scratch was used; a caller-supplied scratch keeps the call asynchronous
    if (!stream || !d_scratch) ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}

int zkgpu_domain_ntt_fr(uint64_t* a, uint32_t k, int inverse, size_t m) {
    API_BEGIN
    ZK_REQUIRE(k <= 24, "k out of range");
    if (m) {
        if (inverse) ntt_host(a, fr_omega_inv(k), k, m, true, fr_pow2_inv(k));
        else ntt_host(a, fr_omega(k), k, m, false, fr_t::zero());
    }
    API_END
}

int zkgpu_coset_ntt_fr(const uint64_t* coeffs, uint32_t k, uint32_t ext_k, uint64_t* out) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(coeffs && out, "null pointer");
    ZK_REQUIRE(ext_k >= k && ext_k <= 24, "coset ntt: need k <= ext_k <= 24");
    size_t n = (size_t)1 << k, en = (size_t)1 << ext_k;
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(n + en);
    C.fr_scratch.ensure(ntt_scratch_elems(ext_k, 1));
    fr_t* d_in = C.fr_buf.p; fr_t* d_out = C.fr_buf.p + n;
    ZK_CUDA(cudaMemcpyAsync(d_in, coeffs, n * 32, cudaMemcpyHostToDevice, st));
    NttJob J;
    J.in = d_in; J.out = d_out; J.scratch = C.fr_scratch.p; J.batch = 1; J.log_n = ext_k; J.omega = fr_omega(ext_k);
    J.in_stride = n; J.in_valid = n;
    J.pre_coset = 1; J.cs1 = fr_from_limbs(fr_consts::ZETA); J.cs2 = fr_from_limbs(fr_consts::ZETA_INV);
    ntt_run(J, st);
    ZK_CUDA(cudaMemcpyAsync(out, d_out, en * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}

int zkgpu_coset_intt_fr(uint64_t* evals, uint32_t k, uint32_t ext_k, uint32_t quotient_degree) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(evals, "null pointer");
    ZK_REQUIRE(ext_k >= k && ext_k <= 24, "coset intt: need k <= ext_k <= 24");
    size_t n = (size_t)1 << k, en = (size_t)1 << ext_k;
    ZK_REQUIRE((size_t)quotient_degree * n <= en, "coset intt: quotient_degree * 2^k exceeds the extended domain");
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(en);
    C.fr_scratch.ensure(ntt_scratch_elems(ext_k, 1));
    ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, evals, en * 32, cudaMemcpyHostToDevice, st));
    NttJob J;
    J.in = C.fr_buf.p; J.out = C.fr_buf.p; J.scratch = C.fr_scratch.p; J.batch = 1; J.log_n = ext_k; J.omega = fr_omega_inv(ext_k);
    J.post_coset = 1; J.cs1 = fr_from_limbs(fr_consts::ZETA_INV); J.cs2 = fr_from_limbs(fr_consts::ZETA);
    J.has_scale = 1; J.scale = fr_pow2_inv(ext_k);
    ntt_run(J, st);
    size_t keep = (size_t)quotient_degree * n;
    ZK_CUDA(cudaMemcpyAsync(evals, C.fr_buf.p, keep * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    memset(evals + 4 * keep, 0, (en - keep) * 32);
    API_END
}

/* ---- vectorised Fr helpers (synthetic circuit / witness construction in bench.py and tools) ------ */
int zkgpu_fr_vec_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(op >= 0 && op <= 4 && a && out && (b || op >= 3), "fr_vec_op: bad arguments");
    if (n == 0) return ZKGPU_OK;
    cudaStream_t st = C.stream;
    C.fr_buf.ensure(3 * n);
    ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, a, n * 32, cudaMemcpyHostToDevice, st));
    if (op < 3) ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p + n, b, n * 32, cudaMemcpyHostToDevice, st));
    launch_vec_op(op, C.fr_buf.p, C.fr_buf.p + n, C.fr_buf.p + 2 * n, n, st);
    ZK_CUDA(cudaMemcpyAsync(out, C.fr_buf.p + 2 * n, n * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}
int zkgpu_fr_to_mont(const uint64_t* canonical, uint64_t* out, size_t n) { return zkgpu_fr_vec_op(3, canonical, nullptr, out, n); }
int zkgpu_fr_from_mont(const uint64_t* mont, uint64_t* out, size_t n) { return zkgpu_fr_vec_op(4, mont, nullptr, out, n); }
/* out[i] = Fr::random(rng) (eight next_u64 each, 512-bit reduction) */
static void fr_random_from(SmallRng& rng, uint64_t* out, size_t n) {
    Context& C = ctx();
    cudaStream_t st = C.stream;
    std::vector<uint64_t> raw(8 * n);
    for (size_t i = 0; i < 8 * n; ++i) raw[i] = rng.next_u64();
    C.fr_buf.ensure(3 * n);
    uint64_t* d_raw = reinterpret_cast<uint64_t*>(C.fr_buf.p + n);
    ZK_CUDA(cudaMemcpyAsync(d_raw, raw.data(), raw.size() * 8, cudaMemcpyHostToDevice, st));
    launch_reduce_wide(d_raw, C.fr_buf.p, n, st);
    ZK_CUDA(cudaMemcpyAsync(out, C.fr_buf.p, n * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
}
/* rng = SmallRng::seed_from_u64(seed) */
int zkgpu_fr_random(uint64_t seed, uint64_t* out, size_t n) {
    API_BEGIN
    ZK_REQUIRE(out || n == 0, "null pointer");
    if (n == 0) return ZKGPU_OK;
    SmallRng rng(seed);
    fr_random_from(rng, out, n);
    API_END
}
/* rng = the caller's running SmallRng (xoshiro256++ state, advanced in place) */
int zkgpu_fr_random_rng(uint64_t rng_state[4], uint64_t* out, size_t n) {
    API_BEGIN
    ZK_REQUIRE(rng_state && (out || n == 0), "null pointer");
    if (n == 0) return ZKGPU_OK;
    SmallRng rng(rng_state);
    fr_random_from(rng, out, n);
    memcpy(rng_state, rng.s, 32);
    API_END
}

/* halo2_proofs::arithmetic::eval_polynomial(poly, point) (reference call site
 * /root/reference/crates/powers-of-tau/lib.rs:151): Horner evaluation of n coefficients at x. */
int zkgpu_eval_polynomial(const uint64_t* coeffs, size_t n, const uint64_t x[4], uint64_t out[4]) {
    API_BEGIN
    Context& C = ctx();
    ZK_REQUIRE((coeffs || n == 0) && x && out, "null pointer");
    fr_t xv; memcpy(xv.l, x, 32);
    fr_t acc = fr_t::zero();
    if (n) {
        // chunks of 2^log_c coefficients (zero-padded at the top), one CTA each; the chunk values are folded on the host:
        // p(x) = sum_j x^(j 2^log_c) p_j(x)
        unsigned log_c = 0;
        while (log_c < 16 && ((size_t)1 << log_c) < n) ++log_c;
        const size_t chunk = (size_t)1 << log_c, jobs_n = (n + chunk - 1) / chunk;
        cudaStream_t st = C.stream;
        C.fr_buf.ensure(jobs_n * chunk + jobs_n);
        fr_t* d_vals = C.fr_buf.p + jobs_n * chunk;
        if (jobs_n * chunk > n) ZK_CUDA(cudaMemsetAsync(C.fr_buf.p + n, 0, (jobs_n * chunk - n) * sizeof(fr_t), st));
        ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, coeffs, n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        std::vector<EvalJob> jobs(jobs_n);
        for (size_t j = 0; j < jobs_n; ++j) { jobs[j].poly = C.fr_buf.p + j * chunk; jobs[j].x = xv; }
        DevBuf<EvalJob> d_jobs(jobs_n);
        ZK_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs_n * sizeof(EvalJob), cudaMemcpyHostToDevice, st));
        launch_poly_eval(d_jobs.p, d_vals, jobs_n, log_c, st);
        std::vector<fr_t> vals(jobs_n);
        ZK_CUDA(cudaMemcpyAsync(vals.data(), d_vals, jobs_n * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        fr_t xc = xv;
        for (unsigned i = 0; i < log_c; ++i) xc = sqr(xc);
        for (size_t j = jobs_n; j-- > 0;) acc = acc * xc + vals[j];
    }
    memcpy(out, acc.l, 32);
    API_END
}

int zkgpu_fft_g1(uint64_t* points_jacobian, const uint64_t omega[4], uint32_t log_n) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(points_jacobian && omega, "null pointer");
    ZK_REQUIRE(log_n <= 20, "g1 fft: log_n out of range");
    size_t n = (size_t)1 << log_n;
    fr_t w; memcpy(w.l, omega, 32);
    cudaStream_t st = C.stream;
    // Jacobian (X,Y,Z) -> XYZZ (X, Y, Z^2, Z^3) on the host (two products per point)
    std::vector<g1_xyzz_t> h(n);
    for (size_t i = 0; i < n; ++i) {
        fq_t X, Y, Z;
        memcpy(X.l, points_jacobian + 12 * i, 32); memcpy(Y.l, points_jacobian + 12 * i + 4, 32); memcpy(Z.l, points_jacobian + 12 * i + 8, 32);
        h[i].x = X; h[i].y = Y; h[i].zz = sqr(Z); h[i].zzz = h[i].zz * Z;
    }
    C.xyzz_buf.ensure(n); C.aff_buf.ensure(n);
    ZK_CUDA(cudaMemcpyAsync(C.xyzz_buf.p, h.data(), n * sizeof(g1_xyzz_t), cudaMemcpyHostToDevice, st));
    g1_fft(C.xyzz_buf.p, log_n, w, st);
    g1_normalize(C.xyzz_buf.p, C.aff_buf.p, n, st);
    std::vector<g1_affine_t> a(n);
    ZK_CUDA(cudaMemcpyAsync(a.data(), C.aff_buf.p, n * 64, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n; ++i) jacobian_out(a[i], points_jacobian + 12 * i);
    API_END
}

int zkgpu_g_to_lagrange(const uint64_t* g_affine, uint32_t k, uint64_t* out_affine) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(g_affine && out_affine, "null pointer");
    ZK_REQUIRE(k <= 20, "g_to_lagrange: k out of range");
    size_t n = (size_t)1 << k;
    cudaStream_t st = C.stream;
    C.pt_buf.ensure(n); C.xyzz_buf.ensure(n); C.aff_buf.ensure(n);
    ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, g_affine, n * 64, cudaMemcpyHostToDevice, st));
    g1_from_affine(C.pt_buf.p, C.xyzz_buf.p, n, st);
    g1_fft(C.xyzz_buf.p, k, fr_omega_inv(k), st);
    g1_scale(C.xyzz_buf.p, n, fr_pow2_inv(k), st);
    g1_normalize(C.xyzz_buf.p, C.aff_buf.p, n, st);
    ZK_CUDA(cudaMemcpyAsync(out_affine, C.aff_buf.p, n * 64, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}

}  // extern "C"
