// libzkgpu C ABI (include/zkgpu.h): host-buffer entry points that a patched halo2curves /
// halo2_proofs would call in place of best_multiexp / best_fft / EvaluationDomain transforms, plus
// device-resident variants for the prover pipeline and the benchmark.  No CPU fallback anywhere: if
// CUDA is unavailable every compute call returns ZKGPU_ERR_CUDA.
#include "../../include/zkgpu.h"
#include "context.cuh"
#include "prover_kernels.cuh"
#include "host_util.hpp"

namespace zk {
void prover_release_all();  // prover.cu
std::atomic<uint64_t> g_launches{0};
Context& ctx() { static Context c; return c; }
thread_local std::string g_last_error;

// ---- per-kernel-class timing ------------------------------------------------------------------
bool g_ktime_on = false;
namespace {
struct KtSlot { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; std::vector<cudaEvent_t> free_ev; double ms = 0; uint64_t launches = 0; };
KtSlot g_kt[KT_SLOTS];
std::mutex g_kt_mu;
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
    KtSlot& s = g_kt[slot];
    cudaEvent_t a = kt_event(s), b = kt_event(s);
    cudaEventRecord(a, st);
    s.pending.push_back({a, b});
}
void ktime_end(int slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    KtSlot& s = g_kt[slot];
    if (!s.pending.empty()) cudaEventRecord(s.pending.back().second, st);
    if (s.pending.size() > 4096) kt_collect(s);
}

void Context::init(int dev) {
    std::lock_guard<std::recursive_mutex> lk(mu);
    if (inited) {
        ZK_REQUIRE(dev == device, "zkgpu_init: already bound to another device (one process per GPU)");
        return;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(ZK_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libzkgpu has no CPU fallback)");
    ZK_REQUIRE(dev >= 0 && dev < count, "zkgpu_init: bad device index");
    ZK_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        throw Error(ZK_ERR_CUDA, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                     "); libzkgpu is built for sm_100a only");
    ZK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    device = dev; sm_count = prop.multiProcessorCount; inited = true;
}
void Context::require() {
    if (!inited) init(0);
    ZK_CUDA(cudaSetDevice(device));
}
void Context::shutdown() {
    std::lock_guard<std::recursive_mutex> lk(mu);
    if (!inited) return;
    cudaStreamSynchronize(stream);
    prover_release_all();
    srs.clear();
    ws = MsmWorkspace();
    fr_buf.release(); fr_scratch.release(); pt_buf.release(); xyzz_buf.release(); aff_buf.release();
    ntt_clear_cache();
    cudaStreamDestroy(stream);
    inited = false; device = -1;
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

#define API_BEGIN try { std::lock_guard<std::recursive_mutex> lk_(ctx().mu);
#define API_END                                                           \
    return ZKGPU_OK; }                                                    \
    catch (const zk::Error& e) { g_last_error = e.what(); return e.code; } \
    catch (const std::exception& e) { g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }

extern "C" {

int zkgpu_abi_version(void) { return 2; }
const char* zkgpu_last_error(void) { return g_last_error.c_str(); }
uint64_t zkgpu_launch_count(void) { return g_launches.load(); }

void zkgpu_kernel_timing(int enable) { g_ktime_on = enable != 0; }
int zkgpu_kernel_times(int slot, double* total_ms, uint64_t* launches, int reset) {
    if (slot < 0 || slot >= KT_SLOTS) return ZKGPU_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_kt_mu);
    kt_collect(g_kt[slot]);
    if (total_ms) *total_ms = g_kt[slot].ms;
    if (launches) *launches = g_kt[slot].launches;
    if (reset) { g_kt[slot].ms = 0; g_kt[slot].launches = 0; }
    return ZKGPU_OK;
}
void* zkgpu_stream(void) { return ctx().inited ? (void*)ctx().stream : nullptr; }

int zkgpu_init(int device) {
    API_BEGIN
    ctx().init(device);
    API_END
}
void zkgpu_shutdown(void) {
    try { ctx().shutdown(); } catch (...) {}
}

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
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(g && g_lagrange && handle_out, "null pointer");
    ZK_REQUIRE(k >= 1 && k <= 20, "srs: k out of range");
    std::unique_ptr<SrsEntry> S(new SrsEntry);
    S->k = k; S->n = (size_t)1 << k;
    S->plan = msm_plan(S->n, true);
    S->plan.tstride = S->n;
    cudaStream_t st = C.stream;
    C.pt_buf.ensure(S->n);
    const uint64_t* src[2] = {g, g_lagrange};
    for (int b = 0; b < 2; ++b) {
        S->table[b].alloc(S->n * S->plan.W);
        ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, src[b], S->n * 64, cudaMemcpyHostToDevice, st));
        msm_precompute_table(S->plan, C.pt_buf.p, S->table[b].p, st);
    }
    ZK_CUDA(cudaStreamSynchronize(st));
    uint64_t h = C.next_handle++;
    C.srs[h] = std::move(S);
    *handle_out = h;
    API_END
}
int zkgpu_srs_release(uint64_t h) {
    API_BEGIN
    Context& C = ctx();
    ZK_REQUIRE(C.srs.erase(h) == 1, "unknown SRS handle");
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
    API_BEGIN
    Context& C = ctx(); C.require();
    SrsEntry& S = C.get_srs(srs);
    cudaStream_t st = stream ? (cudaStream_t)stream : C.stream;
    C.msm_srs_dev(S, basis, (const fr_t*)d_scalars, n, m, (g1_affine_t*)d_out_affine, st);
    if (!stream) ZK_CUDA(cudaStreamSynchronize(st));
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
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(d_a && omega, "null pointer");
    fr_t w; memcpy(w.l, omega, 32);
    cudaStream_t st = stream ? (cudaStream_t)stream : C.stream;
    NttJob J;
    J.in = (fr_t*)d_a; J.out = (fr_t*)d_a; J.scratch = (fr_t*)d_scratch; J.batch = m; J.log_n = log_n; J.omega = w;
    if (!J.scratch && ntt_scratch_elems(log_n, m)) {
        C.fr_scratch.ensure(ntt_scratch_elems(log_n, m));
        J.scratch = C.fr_scratch.p;
    }
    ntt_run(J, st);
    if (!stream) ZK_CUDA(cudaStreamSynchronize(st));
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
/* out[i] = Fr::random(rng) for rng = SmallRng::seed_from_u64(seed) (eight next_u64 each, 512-bit reduction) */
int zkgpu_fr_random(uint64_t seed, uint64_t* out, size_t n) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(out || n == 0, "null pointer");
    if (n == 0) return ZKGPU_OK;
    cudaStream_t st = C.stream;
    std::vector<uint64_t> raw(8 * n);
    SmallRng rng(seed);
    for (size_t i = 0; i < 8 * n; ++i) raw[i] = rng.next_u64();
    C.fr_buf.ensure(3 * n);
    uint64_t* d_raw = reinterpret_cast<uint64_t*>(C.fr_buf.p + n);
    ZK_CUDA(cudaMemcpyAsync(d_raw, raw.data(), raw.size() * 8, cudaMemcpyHostToDevice, st));
    launch_reduce_wide(d_raw, C.fr_buf.p, n, st);
    ZK_CUDA(cudaMemcpyAsync(out, C.fr_buf.p, n * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
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
