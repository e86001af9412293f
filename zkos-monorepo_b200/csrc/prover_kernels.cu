// K5/K7/K8 and friends — device kernels of the batched halo2 create_proof pipeline
// (SURVEY.md §3.2 steps 1-10, §8a rows a7, a9, a11).  Semantics follow halo2_proofs v0.3.0
// plonk/{prover,evaluation}.rs, permutation/prover.rs, vanishing/prover.rs and
// poly/kzg/multiopen/shplonk/prover.rs; every kernel is checked bit-for-bit through whole-proof
// byte equality against the CPU oracle (tests/test_gpu_prover.py).
#include "prover_kernels.cuh"
#include <cstdlib>
#include "plonk_types.hpp"

namespace zk {

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ fr_t fr_from_wide(const uint64_t* w) {
    // halo2curves from_u512: lo * R^2 + hi * R^3 (Montgomery products) = (lo + hi * 2^256) mod r
    fr_t lo, hi;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lo.l[2 * i] = (uint32_t)w[i]; lo.l[2 * i + 1] = (uint32_t)(w[i] >> 32);
        hi.l[2 * i] = (uint32_t)w[4 + i]; hi.l[2 * i + 1] = (uint32_t)(w[4 + i] >> 32);
    }
    fr_reduce_raw(lo); fr_reduce_raw(hi);
    fr_t r2 = fe_r2<FrTag>();
    fr_t r3 = r2 * r2;
    return lo * r2 + hi * r3;
}
// w^i from a half-size table: tw[i] for i < half, -tw[i-half] otherwise
__device__ __forceinline__ fr_t pow_from_tw(const fr_t* tw, size_t i, size_t half) {
    return i < half ? fe_ldg(tw + i) : neg(fe_ldg(tw + (i - half)));
}
__device__ __forceinline__ fr_t fr_delta() {
    fr_t d;
    d.l[0] = 0xefd78855u; d.l[1] = 0x9a0c322bu; d.l[2] = 0x249b563cu; d.l[3] = 0x46e82d14u;
    d.l[4] = 0xe0b0b7a7u; d.l[5] = 0x5983a663u; d.l[6] = 0xaaa111adu; d.l[7] = 0x22ab452bu;
    return d;
}

// postfix expression interpreter shared by the quotient evaluation and the lookup compression
template <class FF, class FA, class FI>
__device__ __forceinline__ fr_t run_expr(const uint32_t* prog, uint32_t pc0, uint32_t pc1, const fr_t* constants, FF fixed_at, FA advice_at, FI inst_at) {
    fr_t stack[8];
    int sp = 0;
    for (uint32_t pc = pc0; pc < pc1; ++pc) {
        uint32_t op = prog[2 * pc], arg = prog[2 * pc + 1];
        switch (op) {
            case OP_CONST: stack[sp++] = fe_ldg(constants + arg); break;
            case OP_FIXED: stack[sp++] = fixed_at(arg); break;
            case OP_ADVICE: stack[sp++] = advice_at(arg); break;
            case OP_INSTANCE: stack[sp++] = inst_at(arg); break;
            case OP_NEG: stack[sp - 1] = neg(stack[sp - 1]); break;
            case OP_ADD: stack[sp - 2] = stack[sp - 2] + stack[sp - 1]; --sp; break;
            case OP_MUL: stack[sp - 2] = stack[sp - 2] * stack[sp - 1]; --sp; break;
            default: stack[sp - 1] = stack[sp - 1] * fe_ldg(constants + arg); break;  // OP_SCALE
        }
    }
    return stack[0];
}

// ---------------------------------------------------------------------------------------------
// blinding rows
// ---------------------------------------------------------------------------------------------
__global__ void k_scatter_random(fr_t* dst, size_t proof_stride, size_t col_stride, size_t row_start, const uint64_t* raw,
                                 size_t B, size_t cols, size_t rows) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * cols * rows) return;
    size_t j = t % rows, col = (t / rows) % cols, b = t / (rows * cols);
    fe_store(dst + b * proof_stride + col * col_stride + row_start + j, fr_from_wide(raw + 8 * t));
}
void launch_scatter_random(fr_t* dst, size_t proof_stride, size_t col_stride, size_t row_start, const uint64_t* raw, size_t B,
                           size_t cols, size_t rows, cudaStream_t st) {
    size_t total = B * cols * rows;
    if (!total) return;
    KtScope kt(KT_MISC, st);
    ZK_LAUNCH(k_scatter_random, ceil_div(total, 128), 128, 0, st, dst, proof_stride, col_stride, row_start, raw, B, cols, rows);
}
__global__ void k_reduce_wide(const uint64_t* raw, fr_t* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) fe_store(out + t, fr_from_wide(raw + 8 * t));
}
void launch_reduce_wide(const uint64_t* raw, fr_t* out, size_t n, cudaStream_t st) {
    if (n) ZK_LAUNCH(k_reduce_wide, ceil_div(n, 128), 128, 0, st, raw, out, n);
}

// ---------------------------------------------------------------------------------------------
// permutation argument: grand product columns
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ fr_t perm_cell(const PermArgs& a, const ColSrc& c, size_t b, size_t row, size_t n) {
    if (c.type == COL_ADVICE) return fe_load(a.adv + b * a.adv_proof_stride + (size_t)c.index * n + row);
    if (c.type == COL_FIXED) return fe_ldg(a.fixed_vals + (size_t)c.index * n + row);
    return fe_load(a.inst + b * a.inst_proof_stride + row);
}
__global__ void __launch_bounds__(128) k_perm_num_den(const PermArgs a, fr_t* num, fr_t* den, size_t B) {
    const size_t n = (size_t)1 << a.k;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * a.P * n) return;
    size_t row = t & (n - 1);
    unsigned s = (unsigned)((t >> a.k) % a.P);
    size_t b = (t >> a.k) / a.P;
    const fr_t beta = fe_ldg(&a.ch[b].beta), gamma = fe_ldg(&a.ch[b].gamma);
    fr_t bw = beta * pow_from_tw(a.omega_tw, row, n >> 1);  // beta * omega^row
    fr_t nu = fe_one<FrTag>(), de = fe_one<FrTag>();
    unsigned c0 = s * a.chunk, c1 = c0 + a.chunk < a.S ? c0 + a.chunk : a.S;
    for (unsigned c = c0; c < c1; ++c) {
        ColSrc cs = a.cols[c];
        fr_t v = perm_cell(a, cs, b, row, n) + gamma;
        de = de * (beta * fe_ldg(a.sigma_vals + (size_t)c * n + row) + v);
        nu = nu * (fe_ldg(a.delta_pows + c) * bw + v);
    }
    fe_store(num + t, nu);
    fe_store(den + t, de);
}
void launch_perm_num_den(const PermArgs& a, fr_t* num, fr_t* den, size_t B, cudaStream_t st) {
    size_t total = B * a.P << a.k;
    if (total) ZK_LAUNCH(k_perm_num_den, ceil_div(total, 128), 128, 0, st, a, num, den, B);
}

// chunked Montgomery batch inversion: each thread owns CH consecutive elements
#define ZK_INV_CH 16
__global__ void __launch_bounds__(64) k_batch_inverse(fr_t* a, size_t count) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * ZK_INV_CH;
    if (lo >= count) return;
    size_t hi = lo + ZK_INV_CH < count ? lo + ZK_INV_CH : count;
    fr_t pre[ZK_INV_CH];
    fr_t acc = fe_one<FrTag>();
    for (size_t i = lo; i < hi; ++i) {
        pre[i - lo] = acc;
        fr_t v = fe_load(a + i);
        if (!v.is_zero()) acc = acc * v;
    }
    acc = fe_inv(acc);
    for (size_t i = hi; i-- > lo;) {
        fr_t v = fe_load(a + i);
        if (v.is_zero()) continue;
        fe_store(a + i, acc * pre[i - lo]);
        acc = acc * v;
    }
}
void launch_batch_inverse(fr_t* a, size_t count, cudaStream_t st) {
    if (count) ZK_LAUNCH(k_batch_inverse, ceil_div(ceil_div(count, ZK_INV_CH), 64), 64, 0, st, a, count);
}

// exclusive prefix product of frac = num * den_inv over n rows; one CTA per (b, set)
__global__ void __launch_bounds__(1024) k_perm_scan(const fr_t* num, const fr_t* den_inv, fr_t* z, unsigned k) {
    extern __shared__ uint32_t sm_scan[];  // [8][T] limb-major
    const size_t n = (size_t)1 << k;
    const unsigned T = blockDim.x, L = (unsigned)(n / T), t = threadIdx.x;
    const fr_t* nu = num + (size_t)blockIdx.x * n;
    const fr_t* de = den_inv + (size_t)blockIdx.x * n;
    fr_t* zz = z + (size_t)blockIdx.x * n;
    fr_t p = fe_one<FrTag>();
    for (unsigned j = 0; j < L; ++j) p = p * (fe_load(nu + (size_t)t * L + j) * fe_load(de + (size_t)t * L + j));
    fr_t mine = p;
#pragma unroll
    for (int l = 0; l < 8; ++l) sm_scan[l * T + t] = p.l[l];
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        fr_t other;
        bool has = t >= d;
        if (has) {
#pragma unroll
            for (int l = 0; l < 8; ++l) other.l[l] = sm_scan[l * T + t - d];
        }
        __syncthreads();
        if (has) {
            p = p * other;
#pragma unroll
            for (int l = 0; l < 8; ++l) sm_scan[l * T + t] = p.l[l];
        }
        __syncthreads();
    }
    // exclusive prefix for this thread = inclusive of thread t-1
    fr_t run;
    if (t == 0) run = fe_one<FrTag>();
    else {
#pragma unroll
        for (int l = 0; l < 8; ++l) run.l[l] = sm_scan[l * T + t - 1];
    }
    (void)mine;
    for (unsigned j = 0; j < L; ++j) {
        size_t i = (size_t)t * L + j;
        fe_store(zz + i, run);
        run = run * (fe_load(nu + i) * fe_load(de + i));
    }
}
void launch_perm_scan(const fr_t* num, const fr_t* den_inv, fr_t* z, unsigned k, size_t BP, cudaStream_t st) {
    if (!BP) return;
    size_t n = (size_t)1 << k;
    unsigned T = n >= 4096 ? 512 : (n >= 256 ? 128 : (unsigned)n / 2);
    if (T < 1) T = 1;
    ZK_LAUNCH(k_perm_scan, (unsigned)BP, T, (size_t)T * 32, st, num, den_inv, z, k);
}

// carries[b][s] = prod_{t<s} z_local[b][t][u], u = n - bf - 1; then z[b][s][row<n-bf] *= carry, blinding rows
__global__ void k_perm_carries(const fr_t* z, fr_t* carries, unsigned k, unsigned P, unsigned bf, size_t B) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t n = (size_t)1 << k, u = n - bf - 1;
    fr_t c = fe_one<FrTag>();
    for (unsigned s = 0; s < P; ++s) {
        fe_store(carries + b * P + s, c);
        c = c * fe_load(z + (b * P + s) * n + u);
    }
}
__global__ void k_perm_finalize(fr_t* z, const fr_t* carries, unsigned k, unsigned P, unsigned bf, const uint64_t* raw, size_t B) {
    const size_t n = (size_t)1 << k;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * P * n) return;
    size_t row = t & (n - 1), bs = t >> k;
    if (row >= n - bf) {
        fe_store(z + t, fr_from_wide(raw + 8 * (bs * bf + (row - (n - bf)))));
    } else if (bs % P) {
        fe_store(z + t, fe_load(z + t) * fe_load(carries + bs));
    }
}
void launch_perm_finalize(fr_t* z, fr_t* carries, unsigned k, unsigned P, unsigned bf, const uint64_t* raw, size_t B, cudaStream_t st) {
    if (!B || !P) return;
    ZK_LAUNCH(k_perm_carries, ceil_div(B, 64), 64, 0, st, z, carries, k, P, bf, B);
    size_t total = B * P << k;
    ZK_LAUNCH(k_perm_finalize, ceil_div(total, 256), 256, 0, st, z, carries, k, P, bf, raw, B);
}

// ---------------------------------------------------------------------------------------------
// vanishing argument: ChaCha20 random polynomial (rand_chacha ChaCha20Rng, one block per coefficient)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define ZK_QR(a, b, c, d) \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
__global__ void k_chacha_poly(const uint8_t* seeds, fr_t* out, size_t n, size_t B, size_t chunk, size_t nchunks) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n) return;
    size_t b = t / n, i = t - b * n;
    // coefficient i belongs to chunk i / chunk and is the (i % chunk)-th draw of that chunk's stream
    const size_t ci = i / chunk;
    i -= ci * chunk;
    const uint32_t* key = reinterpret_cast<const uint32_t*>(seeds + 32 * (b * nchunks + ci));
    uint32_t s[16], x[16];
    s[0] = 0x61707865; s[1] = 0x3320646e; s[2] = 0x79622d32; s[3] = 0x6b206574;
#pragma unroll
    for (int j = 0; j < 8; ++j) s[4 + j] = key[j];
    s[12] = (uint32_t)i; s[13] = (uint32_t)((uint64_t)i >> 32); s[14] = 0; s[15] = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = s[j];
    for (int r = 0; r < 10; ++r) {
        ZK_QR(0, 4, 8, 12) ZK_QR(1, 5, 9, 13) ZK_QR(2, 6, 10, 14) ZK_QR(3, 7, 11, 15)
        ZK_QR(0, 5, 10, 15) ZK_QR(1, 6, 11, 12) ZK_QR(2, 7, 8, 13) ZK_QR(3, 4, 9, 14)
    }
    uint64_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = (uint64_t)(x[2 * j] + s[2 * j]) | ((uint64_t)(x[2 * j + 1] + s[2 * j + 1]) << 32);
    fe_store(out + t, fr_from_wide(w));
}
void launch_chacha_poly(const uint8_t* seeds, fr_t* out, size_t n, size_t B, size_t chunk, size_t nchunks, cudaStream_t st) {
    KtScope kt(KT_MISC, st);
    if (B * n) ZK_LAUNCH(k_chacha_poly, ceil_div(B * n, 128), 128, 0, st, seeds, out, n, B, chunk, nchunks);
}

// ---------------------------------------------------------------------------------------------
// evaluate_h: quotient numerator on the Qc quotient cosets (coset-major rows), divided by the vanishing polynomial
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_eval_h(const EvalHArgs a, fr_t* h, size_t B) {
    // Rows are coset-major: row i = c * n + r is the point g_c * omega^r, g_c = zeta * ext_omega^c, c < Qc.  `en` below is
    // the number of rows per column (Qc * n), not the size of halo2's extended domain.
    extern __shared__ uint32_t eh_sm[];
    uint32_t* s_prog = eh_sm;                                   // [2 * n_prog]
    uint32_t* s_goff = s_prog + 2 * a.n_prog;                   // [num_gates + 1]
    int32_t* s_advq = reinterpret_cast<int32_t*>(s_goff + a.num_gates + 1);   // [2 * n_adv_q]
    int32_t* s_fixq = s_advq + 2 * a.n_adv_q;
    int32_t* s_instq = s_fixq + 2 * a.n_fix_q;
    for (unsigned t = threadIdx.x; t < 2 * a.n_prog; t += blockDim.x) s_prog[t] = a.prog[t];
    for (unsigned t = threadIdx.x; t < a.num_gates + 1; t += blockDim.x) s_goff[t] = a.gate_off[t];
    for (unsigned t = threadIdx.x; t < 2 * a.n_adv_q; t += blockDim.x) s_advq[t] = a.adv_q[t];
    for (unsigned t = threadIdx.x; t < 2 * a.n_fix_q; t += blockDim.x) s_fixq[t] = a.fix_q[t];
    for (unsigned t = threadIdx.x; t < 2 * a.n_inst_q; t += blockDim.x) s_instq[t] = a.inst_q[t];
    __syncthreads();
    const size_t n = (size_t)1 << a.k;
    const size_t en = (size_t)a.Qc << a.k;
    // blocks of one row range are adjacent across the B proofs, so the proving-key columns they share stay in L2
    const size_t b = blockIdx.x % B;
    const size_t i = (size_t)(blockIdx.x / B) * blockDim.x + threadIdx.x;
    if (i >= en) return;
    const size_t t = b * en + i;
    const unsigned rs = a.ek - a.k;
    const size_t coset = i >> a.k, cbase = coset << a.k;
    const fr_t* adv = a.adv_ext + b * a.adv_ext_proof_stride;
    const fr_t y = fe_ldg(&a.ch[b].y);
    auto rot = [&](int r) -> size_t { return cbase | ((i + (size_t)(long)r) & (n - 1)); };

    auto fixed_at = [&](uint32_t q) { return fe_ldg(a.fixed_ext + (size_t)s_fixq[2 * q] * en + rot(s_fixq[2 * q + 1])); };
    auto advice_at = [&](uint32_t q) { return fe_load(adv + (size_t)s_advq[2 * q] * en + rot(s_advq[2 * q + 1])); };
    auto inst_at = [&](uint32_t q) { return fe_load(adv + (size_t)a.A * en + rot(s_instq[2 * q + 1])); };
    fr_t v = fr_t::zero();
    for (unsigned g = 0; g < a.num_gates; ++g)
        v = v * y + run_expr(s_prog, s_goff[g], s_goff[g + 1], a.constants, fixed_at, advice_at, inst_at);
    if (a.P) {
        const fr_t* z = a.z_ext + b * a.z_ext_proof_stride;
        const fr_t beta = fe_ldg(&a.ch[b].beta), gamma = fe_ldg(&a.ch[b].gamma);
        const size_t r_next = rot(1), r_last = rot(a.rotation_last);
        const fr_t l0 = fe_ldg(a.l0 + i), one = fe_one<FrTag>();
        fr_t zf = fe_load(z + i);
        v = v * y + (one - zf) * l0;
        fr_t zl = fe_load(z + (size_t)(a.P - 1) * en + i);
        v = v * y + (sqr(zl) - zl) * fe_ldg(a.l_last + i);
        for (unsigned s = 1; s < a.P; ++s)
            v = v * y + (fe_load(z + (size_t)s * en + i) - fe_load(z + (size_t)(s - 1) * en + r_last)) * l0;
        fr_t cur = beta * a.zeta * pow_from_tw(a.ext_tw, ((i & (n - 1)) << rs) + coset, (size_t)1 << (a.ek - 1));
        const fr_t delta = fr_delta();
        const fr_t lact = fe_ldg(a.l_active + i);
        for (unsigned s = 0; s < a.P; ++s) {
            unsigned c0 = s * a.chunk, c1 = c0 + a.chunk < a.S ? c0 + a.chunk : a.S;
            fr_t left = fe_load(z + (size_t)s * en + r_next), right = fe_load(z + (size_t)s * en + i);
            for (unsigned c = c0; c < c1; ++c) {
                ColSrc cs = a.cols[c];
                fr_t val = cs.type == COL_ADVICE ? fe_load(adv + (size_t)cs.index * en + i)
                         : cs.type == COL_FIXED ? fe_ldg(a.fixed_ext + (size_t)cs.index * en + i)
                                                : fe_load(adv + (size_t)a.A * en + i);
                fr_t vg = val + gamma;
                left = left * (beta * fe_ldg(a.sigma_ext + (size_t)c * en + i) + vg);
                right = right * (cur + vg);
                cur = cur * delta;
            }
            v = v * y + (left - right) * lact;
        }
    }
    if (a.lp.L) {
        const fr_t* lk = a.lk_ext + b * a.lk_ext_proof_stride;
        const fr_t theta = fe_ldg(&a.ch[b].theta), beta = fe_ldg(&a.ch[b].beta), gamma = fe_ldg(&a.ch[b].gamma);
        const fr_t l0 = fe_ldg(a.l0 + i), llast = fe_ldg(a.l_last + i), lact = fe_ldg(a.l_active + i), one = fe_one<FrTag>();
        const size_t r_next = rot(1), r_prev = rot(-1);
        for (unsigned l = 0; l < a.lp.L; ++l) {
            const fr_t* zc = lk + (size_t)(3 * l) * en;
            const fr_t* ac = zc + en;
            const fr_t* sc = ac + en;
            const uint32_t e0 = a.lp.lk_off[l], e1 = a.lp.lk_off[l + 1], em = e0 + (e1 - e0) / 2;
            fr_t cin = fr_t::zero(), ctab = fr_t::zero();
            for (uint32_t e = e0; e < em; ++e)
                cin = cin * theta + run_expr(a.lp.prog, a.lp.expr_off[e], a.lp.expr_off[e + 1], a.lp.constants, fixed_at, advice_at, inst_at);
            for (uint32_t e = em; e < e1; ++e)
                ctab = ctab * theta + run_expr(a.lp.prog, a.lp.expr_off[e], a.lp.expr_off[e + 1], a.lp.constants, fixed_at, advice_at, inst_at);
            const fr_t z = fe_load(zc + i), pa = fe_load(ac + i), ps = fe_load(sc + i);
            const fr_t a_minus_s = pa - ps;
            v = v * y + (one - z) * l0;
            v = v * y + (sqr(z) - z) * llast;
            v = v * y + (fe_load(zc + r_next) * ((pa + beta) * (ps + gamma)) - z * ((cin + beta) * (ctab + gamma))) * lact;
            v = v * y + a_minus_s * l0;
            v = v * y + (a_minus_s * (pa - fe_load(ac + r_prev))) * lact;
        }
    }
    v = v * fe_ldg(a.t_inv + coset);
    fe_store(h + t, v);
}
void launch_eval_h(const EvalHArgs& a, fr_t* h, size_t B, cudaStream_t st) {
    size_t rows = (size_t)a.Qc << a.k;
    KtScope kt(KT_EVAL_H, st);
    // (Measured dead end, round 2: four threads per row — one warp per share of the terms, combined through shared memory — for the
    // single-proof regime: bit-identical, and no faster (250 us either way at 41 k rows): the row's time is not its ~190 dependent
    // products but the interpreter's dependent operand loads, which a split does not shorten.)
    const size_t smem = ((size_t)2 * a.n_prog + a.num_gates + 1 + 2 * ((size_t)a.n_adv_q + a.n_fix_q + a.n_inst_q)) * sizeof(uint32_t);
    ZK_REQUIRE(smem <= 40 * 1024, "eval_h: gate programs too large for shared memory");
    if (B) ZK_LAUNCH(k_eval_h, (unsigned)(ceil_div(rows, 128) * B), 128, smem, st, a, h, B);
}

template <int MAXQ>
__global__ void __launch_bounds__(128) k_coset_combine(fr_t* hc, const fr_t* ginv, const fr_t* vinv, unsigned Qc, unsigned k, size_t B) {
    const size_t n = (size_t)1 << k;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n) return;
    const size_t b = t >> k, m = t & (n - 1);
    fr_t* base = hc + b * Qc * n + m;
    fr_t d[MAXQ];
#pragma unroll
    for (unsigned c = 0; c < MAXQ; ++c)
        if (c < Qc) d[c] = fe_load(base + c * n) * fe_ldg(ginv + c * n + m);
    for (unsigned j = 0; j < Qc; ++j) {
        fr_t acc = fr_t::zero();
#pragma unroll
        for (unsigned c = 0; c < MAXQ; ++c)
            if (c < Qc) acc = acc + fe_ldg(vinv + j * Qc + c) * d[c];
        fe_store(base + j * n, acc);
    }
}
void launch_coset_combine(fr_t* hc, const fr_t* ginv, const fr_t* vinv, unsigned Qc, unsigned k, size_t B, cudaStream_t st) {
    ZK_REQUIRE(Qc >= 1 && Qc <= 16, "coset combine: at most 16 quotient cosets");
    size_t total = B << k;
    KtScope kt(KT_EVAL_H, st);
    if (!total) return;
    if (Qc <= 8) ZK_LAUNCH(k_coset_combine<8>, ceil_div(total, 128), 128, 0, st, hc, ginv, vinv, Qc, k, B);
    else ZK_LAUNCH(k_coset_combine<16>, ceil_div(total, 128), 128, 0, st, hc, ginv, vinv, Qc, k, B);
}

// ---------------------------------------------------------------------------------------------
// polynomial evaluation (Horner), one CTA per job
// ---------------------------------------------------------------------------------------------
#define ZK_EVAL_T 128
__global__ void __launch_bounds__(ZK_EVAL_T) k_poly_eval(const EvalJob* jobs, fr_t* out, unsigned k) {
    __shared__ uint32_t sm[8 * ZK_EVAL_T];
    const size_t n = (size_t)1 << k;
    const unsigned T = blockDim.x, t = threadIdx.x;
    const unsigned L = (unsigned)(n / T);
    const fr_t* poly = jobs[blockIdx.x].poly;
    const fr_t x = fe_ldg(&jobs[blockIdx.x].x);
    fr_t acc = fr_t::zero();
    for (unsigned j = L; j-- > 0;) acc = acc * x + fe_load(poly + (size_t)t * L + j);
    // x^L (L is a power of two)
    fr_t m = x;
    for (unsigned l = 1; l < L; l <<= 1) m = sqr(m);
#pragma unroll
    for (int l = 0; l < 8; ++l) sm[l * T + t] = acc.l[l];
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        if ((t & (2 * d - 1)) == 0) {
            fr_t o;
#pragma unroll
            for (int l = 0; l < 8; ++l) o.l[l] = sm[l * T + t + d];
            acc = acc + o * m;
#pragma unroll
            for (int l = 0; l < 8; ++l) sm[l * T + t] = acc.l[l];
        }
        m = sqr(m);
        __syncthreads();
    }
    if (t == 0) fe_store(out + blockIdx.x, acc);
}
void launch_poly_eval(const EvalJob* jobs, fr_t* out, size_t num_jobs, unsigned k, cudaStream_t st) {
    if (!num_jobs) return;
    size_t n = (size_t)1 << k;
    unsigned T = n >= ZK_EVAL_T ? ZK_EVAL_T : (unsigned)n;
    KtScope kt(KT_POLY, st);
    ZK_LAUNCH(k_poly_eval, (unsigned)num_jobs, T, 0, st, jobs, out, k);
}

// ---------------------------------------------------------------------------------------------
// linear combinations of polynomials / low-degree correction / division by (X - pt)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_lincomb(const LinTerm* terms, const uint32_t* job_off, fr_t* const* outs, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned job = blockIdx.y;
    fr_t acc = fr_t::zero();
    for (uint32_t t = job_off[job]; t < job_off[job + 1]; ++t) acc = acc + fe_ldg(&terms[t].coef) * fe_load(terms[t].poly + i);
    fe_store(outs[job] + i, acc);
}
void launch_lincomb(const LinTerm* terms, const uint32_t* job_off, fr_t* const* outs, size_t num_jobs, size_t n, cudaStream_t st) {
    if (!num_jobs) return;
    dim3 grid(ceil_div(n, 128), (unsigned)num_jobs);
    KtScope kt(KT_POLY, st);
    ZK_LAUNCH(k_lincomb, grid, 128, 0, st, terms, job_off, outs, n);
}
__global__ void k_sub_low(fr_t* const* polys, const fr_t* low, size_t num_jobs) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_jobs * 4) return;
    size_t j = t >> 2, i = t & 3;
    fe_store(polys[j] + i, fe_load(polys[j] + i) - fe_ldg(low + t));
}
void launch_sub_low(fr_t* const* polys, const fr_t* low, size_t num_jobs, cudaStream_t st) {
    KtScope kt(KT_POLY, st);
    if (num_jobs) ZK_LAUNCH(k_sub_low, ceil_div(num_jobs * 4, 64), 64, 0, st, polys, low, num_jobs);
}

// q_{i-1} = a_i + pt * q_i  (i from n-1 down), i.e. Q_i = sum_{j>=i} a_j pt^(j-i), q[i-1] = Q_i
#define ZK_DIV_T 128
#define ZK_DIV_T_LAT 512   // a handful of jobs (single proofs): more, shorter serial segments per polynomial
__global__ void __launch_bounds__(ZK_DIV_T_LAT) k_kate_div(const DivJob* jobs, unsigned k) {
    __shared__ uint32_t sm[8 * ZK_DIV_T_LAT];
    const size_t n = (size_t)1 << k;
    const unsigned T = blockDim.x, t = threadIdx.x;
    const unsigned L = (unsigned)(n / T);
    const fr_t* a = jobs[blockIdx.x].in;
    fr_t* q = jobs[blockIdx.x].out;
    const fr_t pt = fe_ldg(&jobs[blockIdx.x].pt);
    const fr_t* low = jobs[blockIdx.x].low;  // optional: a(X) -= low[0..4) before dividing
    auto ld = [&](size_t i) { fr_t v = fe_load(a + i); if (low && i < 4) v = v - fe_ldg(low + i); return v; };
    // block value h_t = sum_{j<L} a[tL+j] pt^j
    fr_t hsum = fr_t::zero();
    for (unsigned j = L; j-- > 0;) hsum = hsum * pt + ld((size_t)t * L + j);
    fr_t m = pt;
    for (unsigned l = 1; l < L; l <<= 1) m = sqr(m);  // pt^L
    // inclusive suffix: S_t = sum_{j>=0} m^j h_{t+j}
    fr_t s = hsum;
#pragma unroll
    for (int l = 0; l < 8; ++l) sm[l * T + t] = s.l[l];
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        fr_t o;
        bool has = t + d < T;
        if (has) {
#pragma unroll
            for (int l = 0; l < 8; ++l) o.l[l] = sm[l * T + t + d];
        }
        __syncthreads();
        if (has) {
            s = s + o * m;
#pragma unroll
            for (int l = 0; l < 8; ++l) sm[l * T + t] = s.l[l];
        }
        m = sqr(m);
        __syncthreads();
    }
    // carry into this block = S_{t+1} (zero for the last block)
    fr_t run = fr_t::zero();
    if (t + 1 < T) {
#pragma unroll
        for (int l = 0; l < 8; ++l) run.l[l] = sm[l * T + t + 1];
    }
    for (unsigned j = L; j-- > 0;) {
        size_t i = (size_t)t * L + j;
        run = ld(i) + pt * run;
        if (i) fe_store(q + i - 1, run);
    }
    if (t == T - 1) fe_store(q + n - 1, fr_t::zero());
}
void launch_kate_div(const DivJob* jobs, size_t num_jobs, unsigned k, cudaStream_t st) {
    if (!num_jobs) return;
    size_t n = (size_t)1 << k;
    unsigned T = num_jobs < 2 * 148 ? ZK_DIV_T_LAT : ZK_DIV_T;   // too few polynomials to fill the GPU: split each one finer
    if (n < T) T = (unsigned)n;
    KtScope kt(KT_POLY, st);
    ZK_LAUNCH(k_kate_div, (unsigned)num_jobs, T, 0, st, jobs, k);
}

// ---------------------------------------------------------------------------------------------
// upload from pinned (device-mapped) host memory by the SMs: leaves the host->device copy engine to the
// small per-step uploads, which would otherwise queue behind an 800 MB transfer
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pull_from_host(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < count; i += 4 * stride) {
        uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < count; i += stride) dst[i] = __ldcs(src + i);
}
void launch_pull_from_host(void* dst, const void* mapped_src, size_t bytes, unsigned ctas, cudaStream_t st) {
    ZK_REQUIRE(bytes % 16 == 0, "pull_from_host: size must be a multiple of 16 bytes");
    if (bytes) ZK_LAUNCH(k_pull_from_host, ctas, 256, 0, st, (uint4*)dst, (const uint4*)mapped_src, bytes / 16);
}

// ---------------------------------------------------------------------------------------------
// keygen helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_sigma_values(const uint32_t* map_col, const uint32_t* map_row, const fr_t* delta_pows, const fr_t* omega_tw,
                               fr_t* out, size_t total, unsigned k) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t n = (size_t)1 << k;
    fe_store(out + t, fe_ldg(delta_pows + map_col[t]) * pow_from_tw(omega_tw, map_row[t], n >> 1));
}
void launch_sigma_values(const uint32_t* map_col, const uint32_t* map_row, const fr_t* delta_pows, const fr_t* omega_tw, fr_t* out,
                         size_t S, unsigned k, cudaStream_t st) {
    size_t total = S << k;
    if (total) ZK_LAUNCH(k_sigma_values, ceil_div(total, 128), 128, 0, st, map_col, map_row, delta_pows, omega_tw, out, total, k);
}
__global__ void k_gather_cosets(const fr_t* __restrict__ ext, fr_t* __restrict__ dst, unsigned k, unsigned log_e, size_t total) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t c = t >> k, j = t & (((size_t)1 << k) - 1);
    fe_store(dst + t, fe_load(ext + c + (j << log_e)));
}
void launch_gather_cosets(const fr_t* ext, fr_t* dst, unsigned k, unsigned log_e, unsigned Qc, cudaStream_t st) {
    const size_t total = (size_t)Qc << k;
    if (total) ZK_LAUNCH(k_gather_cosets, ceil_div(total, 128), 128, 0, st, ext, dst, k, log_e, total);
}

__global__ void k_one_minus_sum(const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) fe_store(out + t, fe_one<FrTag>() - fe_load(a + t) - fe_load(b + t));
}
void launch_one_minus_sum(const fr_t* a, const fr_t* b, fr_t* out, size_t n, cudaStream_t st) {
    if (n) ZK_LAUNCH(k_one_minus_sum, ceil_div(n, 128), 128, 0, st, a, b, out, n);
}
__global__ void k_vec_op(int op, const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fr_t x = fe_load(a + t);
    if (op == 3) { fe_store(out + t, to_mont(x)); return; }
    if (op == 4) { fe_store(out + t, from_mont(x)); return; }
    fr_t y = fe_load(b + t);
    fe_store(out + t, op == 0 ? x * y : op == 1 ? x + y : x - y);
}
void launch_vec_op(int op, const fr_t* a, const fr_t* b, fr_t* out, size_t n, cudaStream_t st) {
    if (n) ZK_LAUNCH(k_vec_op, ceil_div(n, 128), 128, 0, st, op, a, b, out, n);
}

}  // namespace zk
