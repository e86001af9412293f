// `ParamsKZG::setup(k, rng)` on the device — the seeded test SRS the reference's own prove/verify tests build
// (/root/reference/crates/halo2-verifier/src/generator.rs:118-119, rng = SmallRng::seed_from_u64(42),
// /root/reference/crates/shielder-setup/lib.rs:29-40): s = Fr::random(rng), g[i] = G * s^i,
// g_lagrange = g_to_lagrange(g) (K6, g1fft.cu).  The production SRS comes from the ppot ceremony files
// through crates/powers-of-tau instead; k = 13 of that ceremony is not in the reference tree
// (.MISSING_LARGE_BLOBS), which is why the benchmark uses this setup.
#include "api_util.hpp"
#include "host_util.hpp"

namespace zk {

__global__ void __launch_bounds__(64) k_fixed_base_mul(const fr_t* __restrict__ scalars, g1_affine_t base, g1_xyzz_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t s = from_mont(fe_load(scalars + i));
    g1_xyzz_t r = xyzz_mul(g1_xyzz_t::from_affine(base), s.l);
    fe_store(&out[i].x, r.x); fe_store(&out[i].y, r.y); fe_store(&out[i].zz, r.zz); fe_store(&out[i].zzz, r.zzz);
}

// y^2 == x^3 + 3 (or the identity (0,0)) for every point: `G1Affine::from_xy(x, y).unwrap()` of the ptau reader
// (/root/reference/crates/powers-of-tau/lib.rs:206-224)
__global__ void k_g1_on_curve(const g1_affine_t* __restrict__ pts, size_t n, unsigned long long* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p;
    p.x = fe_load(&pts[i].x); p.y = fe_load(&pts[i].y);
    if (p.is_identity()) return;
    fq_t three = fq_t::zero(); three.l[0] = 3; three = to_mont(three);
    if (!(sqr(p.y) == sqr(p.x) * p.x + three)) atomicAdd(bad, 1ull);
}

}  // namespace zk

using namespace zk;

extern "C" int zkgpu_g1_on_curve(const uint64_t* points_affine, size_t n, uint64_t* bad_count) {
    try {
        DeviceScope api_scope_(rt().primary());
        Context& C = api_scope_.C;
        ZK_REQUIRE((points_affine || n == 0) && bad_count, "null pointer");
        *bad_count = 0;
        if (!n) return ZKGPU_OK;
        cudaStream_t st = C.stream;
        C.pt_buf.ensure(n);
        DevBuf<unsigned long long> d_bad(1);
        ZK_CUDA(cudaMemsetAsync(d_bad.p, 0, 8, st));
        ZK_CUDA(cudaMemcpyAsync(C.pt_buf.p, points_affine, n * 64, cudaMemcpyHostToDevice, st));
        ZK_LAUNCH(k_g1_on_curve, ceil_div(n, 128), 128, 0, st, C.pt_buf.p, n, d_bad.p);
        unsigned long long bad = 0;
        ZK_CUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        *bad_count = bad;
        return ZKGPU_OK;
    } catch (const zk::Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }
}

static int params_setup_impl(uint32_t k, SmallRng& rng, uint64_t* g_out, uint64_t* g_lagrange_out) {
    try {
        DeviceScope api_scope_(rt().primary());
        Context& C = api_scope_.C;
        ZK_REQUIRE(g_out, "null pointer");
        ZK_REQUIRE(k >= 1 && k <= 24 && (k <= 20 || !g_lagrange_out), "params_setup: k out of range (g_lagrange up to 2^20)");
        const size_t n = (size_t)1 << k;
        cudaStream_t st = C.stream;
        // s = Fr::random(rng) = from_u512 of eight next_u64
        uint64_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = rng.next_u64();
        fr_t lo, hi;
        for (int i = 0; i < 4; ++i) {
            lo.l[2 * i] = (uint32_t)w[i]; lo.l[2 * i + 1] = (uint32_t)(w[i] >> 32);
            hi.l[2 * i] = (uint32_t)w[4 + i]; hi.l[2 * i + 1] = (uint32_t)(w[4 + i] >> 32);
        }
        fr_reduce_raw(lo); fr_reduce_raw(hi);
        fr_t r2 = fe_r2<FrTag>();
        fr_t s = lo * r2 + hi * (r2 * r2);
        std::vector<fr_t> pows(n);
        fr_t cur = fe_one<FrTag>();
        for (size_t i = 0; i < n; ++i) { pows[i] = cur; cur = cur * s; }
        g1_affine_t G;
        G.x = fq_t::zero(); G.y = fq_t::zero();
        G.x.l[0] = 1; G.y.l[0] = 2;
        G.x = to_mont(G.x); G.y = to_mont(G.y);
        C.fr_buf.ensure(n); C.xyzz_buf.ensure(n); C.aff_buf.ensure(n); C.pt_buf.ensure(n);
        ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, pows.data(), n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        ZK_LAUNCH(k_fixed_base_mul, ceil_div(n, 64), 64, 0, st, C.fr_buf.p, G, C.xyzz_buf.p, n);
        g1_normalize(C.xyzz_buf.p, C.pt_buf.p, n, st);
        ZK_CUDA(cudaMemcpyAsync(g_out, C.pt_buf.p, n * 64, cudaMemcpyDeviceToHost, st));
        if (!g_lagrange_out) { ZK_CUDA(cudaStreamSynchronize(st)); return ZKGPU_OK; }
        // g_lagrange = n^-1 * FFT_{omega^-1}(g)
        g1_from_affine(C.pt_buf.p, C.xyzz_buf.p, n, st);
        g1_fft(C.xyzz_buf.p, k, fr_omega_inv(k), st);
        g1_scale(C.xyzz_buf.p, n, fr_pow2_inv(k), st);
        g1_normalize(C.xyzz_buf.p, C.aff_buf.p, n, st);
        ZK_CUDA(cudaMemcpyAsync(g_lagrange_out, C.aff_buf.p, n * 64, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        return ZKGPU_OK;
    } catch (const zk::Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }
}

// g_out[i] = s^(start + i) * G for i < count, s = the toxic scalar `ParamsKZG::setup` draws from SmallRng::seed_from_u64(seed):
// a slice of the setup SRS, so a rank of a point-sharded MSM can build only its shard of the bases.
extern "C" int zkgpu_setup_powers(uint64_t seed, uint64_t start, size_t count, uint64_t* g_out) {
    try {
        DeviceScope api_scope_(rt().primary());
        Context& C = api_scope_.C;
        ZK_REQUIRE(g_out || count == 0, "null pointer");
        ZK_REQUIRE(count <= ((size_t)1 << 26), "setup_powers: at most 2^26 points per call");
        if (!count) return ZKGPU_OK;
        cudaStream_t st = C.stream;
        SmallRng rng(seed);
        uint64_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = rng.next_u64();
        fr_t lo, hi;
        for (int i = 0; i < 4; ++i) {
            lo.l[2 * i] = (uint32_t)w[i]; lo.l[2 * i + 1] = (uint32_t)(w[i] >> 32);
            hi.l[2 * i] = (uint32_t)w[4 + i]; hi.l[2 * i + 1] = (uint32_t)(w[4 + i] >> 32);
        }
        fr_reduce_raw(lo); fr_reduce_raw(hi);
        fr_t r2 = fe_r2<FrTag>();
        fr_t s = lo * r2 + hi * (r2 * r2);
        std::vector<fr_t> pows(count);
        fr_t cur = fr_pow_u64(s, start);
        for (size_t i = 0; i < count; ++i) { pows[i] = cur; cur = cur * s; }
        g1_affine_t G;
        G.x = fq_t::zero(); G.y = fq_t::zero();
        G.x.l[0] = 1; G.y.l[0] = 2;
        G.x = to_mont(G.x); G.y = to_mont(G.y);
        C.fr_buf.ensure(count); C.xyzz_buf.ensure(count); C.pt_buf.ensure(count);
        ZK_CUDA(cudaMemcpyAsync(C.fr_buf.p, pows.data(), count * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        ZK_LAUNCH(k_fixed_base_mul, ceil_div(count, 64), 64, 0, st, C.fr_buf.p, G, C.xyzz_buf.p, count);
        g1_normalize(C.xyzz_buf.p, C.pt_buf.p, count, st);
        ZK_CUDA(cudaMemcpyAsync(g_out, C.pt_buf.p, count * 64, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        return ZKGPU_OK;
    } catch (const zk::Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }
}

extern "C" int zkgpu_params_setup(uint32_t k, uint64_t seed, uint64_t* g_out, uint64_t* g_lagrange_out) {
    SmallRng rng(seed);
    return params_setup_impl(k, rng, g_out, g_lagrange_out);
}
// The same from a RUNNING SmallRng: `rng_state` is the caller's xoshiro256++ state; it is advanced by the one `Fr::random`
// that `ParamsKZG::setup` draws, so the caller's stream continues exactly as the reference's does when one rng produces the
// SRS, the witness and the proof (/root/reference/crates/halo2-verifier/src/generator.rs:117-130).
extern "C" int zkgpu_params_setup_rng(uint32_t k, uint64_t rng_state[4], uint64_t* g_out, uint64_t* g_lagrange_out) {
    if (!rng_state) { g_last_error = "null pointer"; return ZKGPU_ERR_ARG; }
    SmallRng rng(rng_state);
    int rc = params_setup_impl(k, rng, g_out, g_lagrange_out);
    if (rc == ZKGPU_OK) memcpy(rng_state, rng.s, 32);
    return rc;
}

// Sum of n affine points on the HOST (no GPU needed): the combine step of a point-sharded MSM, where every GPU
// contributes one partial result (gather of <= 8 points, then G-1 additions).
extern "C" int zkgpu_g1_sum_affine(const uint64_t* points_affine, size_t n, uint64_t out_affine[8]) {
    try {
        ZK_REQUIRE((points_affine || n == 0) && out_affine, "null pointer");
        g1_xyzz_t acc = g1_xyzz_t::identity();
        for (size_t i = 0; i < n; ++i) {
            g1_affine_t p;
            memcpy(&p, points_affine + 8 * i, 64);
            xyzz_madd(acc, p, false);
        }
        g1_affine_t r = xyzz_to_affine(acc);
        memcpy(out_affine, &r, 64);
        return ZKGPU_OK;
    } catch (const zk::Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::exception& e) { g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }
}
