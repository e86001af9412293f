// K6 — group-valued FFT over BN254 G1: halo2curves `best_fft::<Fr, G1>` as used by
// `g_to_lagrange` in ParamsKZG::{from_parts, downsize} (reference call site
// /root/reference/crates/powers-of-tau/lib.rs:71).  Butterflies are point add/sub, the twiddle
// multiply is a 254-bit scalar multiplication.  Runs once per params load, so it is a plain
// global-memory radix-2 DIT: bit-reversal permutation, then log_n stage kernels of n/2 threads.
#include "g1fft.cuh"
#include "ntt.cuh"

namespace zk {

__device__ __forceinline__ g1_xyzz_t ld_xyzz(const g1_xyzz_t* p) {
    g1_xyzz_t r;
    r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz);
    return r;
}
__device__ __forceinline__ void st_xyzz(g1_xyzz_t* p, const g1_xyzz_t& v) {
    fe_store(&p->x, v.x); fe_store(&p->y, v.y); fe_store(&p->zz, v.zz); fe_store(&p->zzz, v.zzz);
}

__global__ void k_g1_bitrev(g1_xyzz_t* a, unsigned log_n) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ((size_t)1 << log_n)) return;
    size_t rk = __brev((unsigned)k) >> (32 - log_n);
    if (k < rk) {
        g1_xyzz_t x = ld_xyzz(a + k), y = ld_xyzz(a + rk);
        st_xyzz(a + k, y); st_xyzz(a + rk, x);
    }
}

// stage s: pairs (i, i + 2^s), twiddle w^(jj * n / 2^(s+1))
__global__ void __launch_bounds__(64) k_g1_stage(g1_xyzz_t* a, const fr_t* tw, unsigned log_n, unsigned s) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ((size_t)1 << (log_n - 1))) return;
    size_t half = (size_t)1 << s;
    size_t jj = b & (half - 1);
    size_t i = ((b >> s) << (s + 1)) | jj;
    g1_xyzz_t lo = ld_xyzz(a + i), hi = ld_xyzz(a + i + half);
    if (jj) {
        fr_t w = from_mont(fe_ldg(tw + (jj << (log_n - s - 1))));
        hi = xyzz_mul(hi, w.l);
    }
    st_xyzz(a + i, xyzz_add(lo, hi));
    st_xyzz(a + i + half, xyzz_add(lo, xyzz_neg(hi)));
}

__global__ void __launch_bounds__(64) k_g1_scale(g1_xyzz_t* a, size_t n, fr_t scalar_canonical) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_xyzz(a + i, xyzz_mul(ld_xyzz(a + i), scalar_canonical.l));
}

__global__ void k_g1_from_affine(const g1_affine_t* in, g1_xyzz_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p;
    p.x = fe_load(&in[i].x); p.y = fe_load(&in[i].y);
    st_xyzz(out + i, g1_xyzz_t::from_affine(p));
}

void g1_fft(g1_xyzz_t* d_a, unsigned log_n, const fr_t& omega, cudaStream_t st) {
    if (log_n == 0) return;
    size_t n = (size_t)1 << log_n;
    const fr_t* tw = ntt_twiddles(log_n, omega, st);
    ZK_LAUNCH(k_g1_bitrev, ceil_div(n, 128), 128, 0, st, d_a, log_n);
    for (unsigned s = 0; s < log_n; ++s) ZK_LAUNCH(k_g1_stage, ceil_div(n / 2, 64), 64, 0, st, d_a, tw, log_n, s);
}
void g1_scale(g1_xyzz_t* d_a, size_t n, const fr_t& scalar_mont, cudaStream_t st) {
    fr_t c = from_mont(scalar_mont);
    ZK_LAUNCH(k_g1_scale, ceil_div(n, 64), 64, 0, st, d_a, n, c);
}
void g1_from_affine(const g1_affine_t* d_in, g1_xyzz_t* d_out, size_t n, cudaStream_t st) {
    ZK_LAUNCH(k_g1_from_affine, ceil_div(n, 128), 128, 0, st, d_in, d_out, n);
}

}  // namespace zk
