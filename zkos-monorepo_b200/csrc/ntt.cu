// K3/K4 — BN254 Fr NTT on the device.
//
// Replaces halo2curves `best_fft::<Fr,Fr>` (natural order in and out) and the
// `EvaluationDomain::{coeff_to_extended, extended_to_coeff}` wrappers around it (row a6).
//
// Structure: a transform of N = 2^log_N points is one or two "tile passes".  A tile pass loads
// C interleaved sequences of m = 2^log_m points into shared memory (limb-major planes, bit-reversed
// and XOR-swizzled so both the bit-reversed fill and the butterfly sweeps are bank-conflict free),
// runs log_m radix-2 DIT stages there, and stores.
//   * N <= 2^12: one pass, the whole polynomial lives in one CTA's shared memory (128 KB at 2^12).
//   * larger N = n1*n2 (four-step): pass 1 transforms columns (stride n2, C adjacent columns per tile so
//     every global access is C*32 contiguous bytes), multiplies by w_N^(i2*k1), writes in place to a
//     scratch buffer; pass 2 transforms rows (contiguous) and writes transposed (k1 + n1*k2), again C
//     rows per tile so the strided stores are C*32-byte segments.  Natural order out, no separate
//     transpose or bit-reversal pass.
// Coset scaling (zeta^(i mod 3), zeta^3 = 1), zero-padding and the 1/N factor of inverse transforms
// are fused into the load / store of the passes.
#include "ntt.cuh"
#include <cooperative_groups.h>
#include <map>
#include <mutex>
#include <array>

namespace zk {

struct NttPass {
    const fr_t* in;
    fr_t* out;
    const fr_t* tw;  // w_N^i, i < N/2
    unsigned log_N, log_m, log_C, swz_q;
    unsigned tiles_per_poly;
    size_t in_poly_stride, out_poly_stride;
    size_t in_inner, in_outer_stride, out_inner, out_outer_stride;  // poly p -> (p / inner) * outer_stride + (p % inner) * poly_stride
    size_t in_tile_stride, in_sj, in_sc;
    size_t out_tile_stride, out_sj, out_sc;
    size_t in_valid;     // elements >= in_valid (index within the polynomial) read as zero
    int c_fastest_in, c_fastest_out;
    int post_twiddle;    // multiply (k, c) by w_N^((tile*C + c) * k)
    int pre_coset;       // multiply input element i by cs1 (i%3==1) / cs2 (i%3==2)
    int post_coset;      // same on output index
    int has_scale;
    int first_window_trivial;  // inputs are zero for sequence index j >= m/8: the first three stages only replicate
    fr_t scale, cs1, cs2;
    const fr_t* pre_table;     // optional: input element gi of polynomial p is multiplied by pre_table[(p % pre_count) << log_N | gi]
    unsigned pre_count;
};

// Shared-memory tile: two planes of uint4 (low / high 16 bytes of every element), element e of the tile at
// plane[phys(e)].  phys XOR-folds the upper index bits into the low three, so the 8 lanes of a quarter warp
// (one 128-byte LDS.128 / STS.128 wavefront) hit 8 different 16-byte bank groups whenever the three index
// bits they enumerate are distinct mod 3.
__device__ __forceinline__ unsigned phys(unsigned e) { return e ^ ((e >> 3) & 7) ^ ((e >> 6) & 7) ^ ((e >> 9) & 7); }
__device__ __forceinline__ fr_t tile_ld(const uint4* sm, unsigned total, unsigned e) {
    const unsigned w = phys(e);
    uint4 a = sm[w], b = sm[total + w];
    fr_t v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}
__device__ __forceinline__ void tile_st(uint4* sm, unsigned total, unsigned e, const fr_t& v) {
    const unsigned w = phys(e);
    sm[w] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[total + w] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// The log_m radix-2 DIT stages of a tile already in shared memory (bit-reversed order).  Radix-2 stages run three at a
// time on 8 register-resident elements per thread (12 butterflies between two __syncthreads), so a 256-point column costs
// 3 shared-memory round trips, not 8.
__device__ __forceinline__ void tile_stages(const NttPass& P, uint4* sm4) {
    const unsigned log_m = P.log_m, log_C = P.log_C;
    const unsigned C = 1u << log_C, log_total = log_m + log_C, total = 1u << log_total;
    if (log_m >= 3) {
        const unsigned groups = total >> 3;
        for (unsigned done = 0; done < log_m;) {
            const unsigned s0 = done + 3 <= log_m ? done : log_m - 3;  // window [s0, s0+3); the last one may overlap
            const unsigned st_begin = done - s0;
            const unsigned fb = log_C + s0;                             // the window's bit field inside e
            for (unsigned g = threadIdx.x; g < groups; g += blockDim.x) {
                unsigned e_base;
                if (fb >= 3 || log_total < fb + 6) e_base = ((g >> fb) << (fb + 3)) | (g & ((1u << fb) - 1));
                else {
                    unsigned rest = g >> 3;
                    e_base = ((rest >> fb) << (fb + 6)) | ((g & 7) << (fb + 3)) | (rest & ((1u << fb) - 1));
                }
                fr_t x[8];
                if (done == 0 && P.first_window_trivial) {
                    // bit-reversed zero-padded input: only x[0] of each group is non-zero, and butterflies with a
                    // zero partner copy it: after three stages all eight outputs equal x[0]
                    x[0] = tile_ld(sm4, total, e_base);
#pragma unroll
                    for (unsigned b = 1; b < 8; ++b) tile_st(sm4, total, e_base | (b << fb), x[0]);
                    continue;
                }
#pragma unroll
                for (unsigned b = 0; b < 8; ++b) x[b] = tile_ld(sm4, total, e_base | (b << fb));
                const unsigned jj0 = (e_base >> log_C) & ((1u << s0) - 1);  // position bits below the window
#pragma unroll
                for (unsigned st = 0; st < 3; ++st) {
                    if (st < st_begin) continue;
                    const unsigned s = s0 + st;
                    const unsigned sh = P.log_N - s - 1;
#pragma unroll
                    for (unsigned q = 0; q < 4; ++q) {
                        // pair: lo has bit st clear; q enumerates the other two bits
                        const unsigned lo = ((q >> st) << (st + 1)) | (q & ((1u << st) - 1));
                        const unsigned hi = lo | (1u << st);
                        const unsigned jj = jj0 | ((lo & ((1u << st) - 1)) << s0);
                        fr_t t = x[hi];
                        if (jj) t = t * fe_ldg(P.tw + ((size_t)jj << sh));
                        x[hi] = x[lo] - t;
                        x[lo] = x[lo] + t;
                    }
                }
#pragma unroll
                for (unsigned b = 0; b < 8; ++b) tile_st(sm4, total, e_base | (b << fb), x[b]);
            }
            __syncthreads();
            done = s0 + 3;
        }
    } else {
        // tiny transforms (m < 8): plain radix-2 stages
        for (unsigned s = 0; s < log_m; ++s) {
            const unsigned half = 1u << s;
            for (unsigned el = threadIdx.x; el < (total >> 1); el += blockDim.x) {
                unsigned c = el & (C - 1), b = el >> log_C;
                unsigned jj = b & (half - 1);
                unsigned i = ((b >> s) << (s + 1)) | jj;
                unsigned lo = (i << log_C) | c, hi = ((i + half) << log_C) | c;
                fr_t a = tile_ld(sm4, total, lo), bb = tile_ld(sm4, total, hi);
                if (jj) bb = bb * fe_ldg(P.tw + ((size_t)jj << (P.log_N - s - 1)));
                tile_st(sm4, total, lo, a + bb);
                tile_st(sm4, total, hi, a - bb);
            }
            __syncthreads();
        }
    }
}

// One tile pass.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_ntt_tile(const NttPass P) {
    extern __shared__ uint4 sm4[];
    const unsigned log_m = P.log_m, log_C = P.log_C;
    const unsigned m = 1u << log_m, C = 1u << log_C, log_total = log_m + log_C, total = 1u << log_total;
    const unsigned tile = blockIdx.x % P.tiles_per_poly;
    const size_t poly = blockIdx.x / P.tiles_per_poly;
    const fr_t* in = P.in + (poly / P.in_inner) * P.in_outer_stride + (poly % P.in_inner) * P.in_poly_stride + (size_t)tile * P.in_tile_stride;
    fr_t* out = P.out + (poly / P.out_inner) * P.out_outer_stride + (poly % P.out_inner) * P.out_poly_stride + (size_t)tile * P.out_tile_stride;

    for (unsigned el = threadIdx.x; el < total; el += blockDim.x) {
        unsigned c, j;
        if (P.c_fastest_in) { c = el & (C - 1); j = el >> log_C; } else { j = el & (m - 1); c = el >> log_m; }
        size_t off = (size_t)j * P.in_sj + (size_t)c * P.in_sc;
        size_t gi = (size_t)tile * P.in_tile_stride + off;  // index within the polynomial
        const bool valid = gi < P.in_valid;
        fr_t v = valid ? fe_load(in + off) : fr_t::zero();
        if (P.pre_coset && valid) {
            unsigned r3 = (unsigned)(gi % 3);
            if (r3 == 1) v = v * P.cs1; else if (r3 == 2) v = v * P.cs2;
        }
        if (P.pre_table && valid) v = v * fe_ldg(P.pre_table + (((size_t)(poly % P.pre_count)) << P.log_N) + gi);
        unsigned p = __brev(j) >> (32 - log_m);
        tile_st(sm4, total, (p << log_C) | c, v);
    }
    __syncthreads();

    tile_stages(P, sm4);

    const size_t halfN = (size_t)1 << (P.log_N - 1);
    for (unsigned el = threadIdx.x; el < total; el += blockDim.x) {
        unsigned c, k;
        if (P.c_fastest_out) { c = el & (C - 1); k = el >> log_C; } else { k = el & (m - 1); c = el >> log_m; }
        fr_t v = tile_ld(sm4, total, (k << log_C) | c);
        if (P.post_twiddle) {
            size_t ex = ((size_t)tile * C + c) * k;
            if (ex) {
                if (ex >= halfN) v = neg(v * fe_ldg(P.tw + (ex - halfN)));
                else v = v * fe_ldg(P.tw + ex);
            }
        }
        size_t off = (size_t)k * P.out_sj + (size_t)c * P.out_sc;
        if (P.post_coset) {
            size_t gi = (size_t)tile * P.out_tile_stride + off;
            unsigned r3 = (unsigned)(gi % 3);
            if (r3 == 1) v = v * P.cs1; else if (r3 == 2) v = v * P.cs2;
        }
        if (P.has_scale) v = v * P.scale;
        fe_store(out + off, v);
    }
}

// Single-pass transform of N = 2^13 points by a CLUSTER of two CTAs (one SM each).  A polynomial of 8192 elements is 256 KB —
// more than one SM's shared memory — so the four-step path above needed two launches, a scratch round trip through HBM/L2
// and N inter-pass twiddle products.  Here CTA c of the pair loads the elements j = c (mod 2) (bit-reversed: the top position
// bit), runs the first 12 stages in its own 128 KB tile — a size-4096 transform of the even / odd subsequence, E[k] and O[k] —
// and the last radix-2 stage  X[k] = E[k] + w^k O[k],  X[k + N/2] = E[k] - w^k O[k]  reads the partner's tile through
// distributed shared memory, each CTA producing (and storing, coalesced) half of the k range.  Same fused options as the
// tile passes (coset scaling, per-element pre-multiplier, zero padding, 1/N).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) k_ntt_cluster2(const NttPass P) {
    extern __shared__ uint4 sm4[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned total = 1u << P.log_m;          // log_m = log_N - 1, log_C = 0
    const size_t poly = blockIdx.x >> 1;
    const fr_t* in = P.in + (poly / P.in_inner) * P.in_outer_stride + (poly % P.in_inner) * P.in_poly_stride;
    fr_t* out = P.out + (poly / P.out_inner) * P.out_outer_stride + (poly % P.out_inner) * P.out_poly_stride;
    for (unsigned el = threadIdx.x; el < total; el += blockDim.x) {
        const size_t gi = 2 * (size_t)el + rank;   // index within the polynomial
        const bool valid = gi < P.in_valid;
        fr_t v = valid ? fe_load(in + gi) : fr_t::zero();
        if (P.pre_coset && valid) {
            unsigned r3 = (unsigned)(gi % 3);
            if (r3 == 1) v = v * P.cs1; else if (r3 == 2) v = v * P.cs2;
        }
        if (P.pre_table && valid) v = v * fe_ldg(P.pre_table + (((size_t)(poly % P.pre_count)) << P.log_N) + gi);
        tile_st(sm4, total, __brev(el) >> (32 - P.log_m), v);
    }
    __syncthreads();
    tile_stages(P, sm4);
    cluster.sync();   // both halves transformed (and every input consumed: in-place transforms may now be overwritten)
    const uint4* even = rank == 0 ? sm4 : cluster.map_shared_rank(sm4, 0);
    const uint4* odd = rank == 1 ? sm4 : cluster.map_shared_rank(sm4, 1);
    const unsigned half = total >> 1;
    for (unsigned i = threadIdx.x; i < half; i += blockDim.x) {
        const unsigned k = rank * half + i;
        fr_t e = tile_ld(even, total, k), t = tile_ld(odd, total, k);
        if (k) t = t * fe_ldg(P.tw + k);
        fr_t lo = e + t, hi = e - t;
        if (P.post_coset) {
            const unsigned r_lo = k % 3, r_hi = (k + total) % 3;
            if (r_lo == 1) lo = lo * P.cs1; else if (r_lo == 2) lo = lo * P.cs2;
            if (r_hi == 1) hi = hi * P.cs1; else if (r_hi == 2) hi = hi * P.cs2;
        }
        if (P.has_scale) { lo = lo * P.scale; hi = hi * P.scale; }
        fe_store(out + k, lo);
        fe_store(out + k + total, hi);
    }
    cluster.sync();   // the partner may still be reading this CTA's tile
}

// The same transform by a cluster of EIGHT CTAs of 128 threads (32 KB tile each): CTA r of the cluster transforms the
// subsequence j = brev3(r) (mod 8) — the elements whose bit-reversed position has r in its top three bits — through 10 local
// stages, and the last THREE stages run as one radix-8 butterfly per output index k' < 1024 across the eight tiles (one local,
// seven read through distributed shared memory); CTA r produces k' in [128 r, 128 r + 128) and stores eight coalesced runs.
// Against the two-CTA version: small CTAs (up to four per SM from different clusters, so one cluster's barriers are covered by
// another's butterflies) instead of one 512-thread CTA per SM.
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(128, 4) k_ntt_cluster8(const NttPass P) {
    extern __shared__ uint4 sm4[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned rho = ((rank & 1u) << 2) | (rank & 2u) | (rank >> 2);   // brev3(rank)
    const unsigned total = 1u << P.log_m;          // log_m = log_N - 3, log_C = 0
    const size_t poly = blockIdx.x >> 3;
    const fr_t* in = P.in + (poly / P.in_inner) * P.in_outer_stride + (poly % P.in_inner) * P.in_poly_stride;
    fr_t* out = P.out + (poly / P.out_inner) * P.out_outer_stride + (poly % P.out_inner) * P.out_poly_stride;
    for (unsigned el = threadIdx.x; el < total; el += blockDim.x) {
        const size_t gi = 8 * (size_t)el + rho;   // index within the polynomial
        const bool valid = gi < P.in_valid;
        fr_t v = valid ? fe_load(in + gi) : fr_t::zero();
        if (P.pre_coset && valid) {
            unsigned r3 = (unsigned)(gi % 3);
            if (r3 == 1) v = v * P.cs1; else if (r3 == 2) v = v * P.cs2;
        }
        if (P.pre_table && valid) v = v * fe_ldg(P.pre_table + (((size_t)(poly % P.pre_count)) << P.log_N) + gi);
        tile_st(sm4, total, __brev(el) >> (32 - P.log_m), v);
    }
    __syncthreads();
    tile_stages(P, sm4);
    cluster.sync();   // all eight sub-transforms done (and every input consumed: in-place transforms may now be overwritten)
    const unsigned per = total >> 3;
    for (unsigned i = threadIdx.x; i < per; i += blockDim.x) {
        const unsigned k = rank * per + i;          // local index k' in every tile; outputs (r << log_m) | k
        fr_t x[8];
#pragma unroll
        for (unsigned r = 0; r < 8; ++r) x[r] = tile_ld(r == rank ? sm4 : cluster.map_shared_rank(sm4, r), total, k);
#pragma unroll
        for (unsigned st = 0; st < 3; ++st) {
            const unsigned s = P.log_m + st;        // stage: pairs differ in bit st of r (bit s of the position)
            const unsigned sh = P.log_N - s - 1;
#pragma unroll
            for (unsigned q = 0; q < 4; ++q) {
                const unsigned lo = ((q >> st) << (st + 1)) | (q & ((1u << st) - 1));
                const unsigned hi = lo | (1u << st);
                const unsigned jj = ((lo & ((1u << st) - 1)) << P.log_m) | k;   // position bits below the stage
                fr_t t = x[hi];
                if (jj) t = t * fe_ldg(P.tw + ((size_t)jj << sh));
                x[hi] = x[lo] - t;
                x[lo] = x[lo] + t;
            }
        }
#pragma unroll
        for (unsigned r = 0; r < 8; ++r) {
            const unsigned o = (r << P.log_m) | k;
            fr_t v = x[r];
            if (P.post_coset) {
                const unsigned r3 = o % 3;
                if (r3 == 1) v = v * P.cs1; else if (r3 == 2) v = v * P.cs2;
            }
            if (P.has_scale) v = v * P.scale;
            fe_store(out + o, v);
        }
    }
    cluster.sync();   // the partners may still be reading this CTA's tile
}

// tw[i] = w^i for i < count, from pows[b] = w^(2^b)
__global__ void k_twiddles(fr_t* tw, size_t count, const fr_t* pows, unsigned nbits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t acc = fe_one<FrTag>();
    for (unsigned b = 0; b < nbits; ++b)
        if ((i >> b) & 1) acc = acc * fe_ldg(pows + b);
    fe_store(tw + i, acc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
fr_t fr_from_limbs(const uint32_t* l) { fr_t r; for (int i = 0; i < 8; ++i) r.l[i] = l[i]; return r; }
fr_t fr_from_u64(uint64_t v) {
    fr_t r = fr_t::zero(); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32);
    return to_mont(r);
}
fr_t fr_pow_u64(fr_t b, uint64_t e) {
    fr_t acc = fe_one<FrTag>();
    while (e) { if (e & 1) acc = acc * b; b = sqr(b); e >>= 1; }
    return acc;
}
fr_t fr_omega(unsigned log_n) {  // ROOT_OF_UNITY^(2^(28-log_n)) — EvaluationDomain::new
    fr_t w = fr_from_limbs(fr_consts::ROOT_OF_UNITY);
    for (unsigned i = log_n; i < fr_consts::S; ++i) w = sqr(w);
    return w;
}
fr_t fr_omega_inv(unsigned log_n) {
    fr_t w = fr_from_limbs(fr_consts::ROOT_OF_UNITY_INV);
    for (unsigned i = log_n; i < fr_consts::S; ++i) w = sqr(w);
    return w;
}
fr_t fr_pow2_inv(unsigned log_n) {  // (2^log_n)^-1
    return fr_pow_u64(fr_from_limbs(fr_consts::TWO_INV), log_n);
}

namespace {
struct TwKey {
    int dev; unsigned log_n; std::array<uint32_t, 8> w;
    bool operator<(const TwKey& o) const {
        if (dev != o.dev) return dev < o.dev;
        if (log_n != o.log_n) return log_n < o.log_n;
        return w < o.w;
    }
};
std::mutex g_tw_mu;
std::map<TwKey, DevBuf<fr_t>> g_tw;
}  // namespace

const fr_t* ntt_twiddles(unsigned log_n, const fr_t& omega, cudaStream_t st) {
    int dev; ZK_CUDA(cudaGetDevice(&dev));
    TwKey key; key.dev = dev; key.log_n = log_n;
    for (int i = 0; i < 8; ++i) key.w[i] = omega.l[i];
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) return it->second.p;
    size_t count = log_n ? ((size_t)1 << (log_n - 1)) : 1;
    std::vector<fr_t> pows(log_n ? log_n : 1);
    fr_t w = omega;
    for (unsigned b = 0; b < pows.size(); ++b) { pows[b] = w; w = sqr(w); }
    DevBuf<fr_t> d_pows(pows.size()), d_tw(count);
    ZK_CUDA(cudaMemcpyAsync(d_pows.p, pows.data(), pows.size() * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    ZK_LAUNCH(k_twiddles, ceil_div(count, 256), 256, 0, st, d_tw.p, count, d_pows.p, log_n ? log_n - 1 : 0);
    ZK_CUDA(cudaStreamSynchronize(st));
    const fr_t* p = d_tw.p;
    g_tw.emplace(key, std::move(d_tw));
    return p;
}
void fr_power_table(fr_t* d_out, unsigned log_count, const fr_t& base, cudaStream_t st) {
    std::vector<fr_t> pows(log_count ? log_count : 1);
    fr_t w = base;
    for (unsigned b = 0; b < pows.size(); ++b) { pows[b] = w; w = sqr(w); }
    DevBuf<fr_t> d_pows(pows.size());
    ZK_CUDA(cudaMemcpyAsync(d_pows.p, pows.data(), pows.size() * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    size_t count = (size_t)1 << log_count;
    ZK_LAUNCH(k_twiddles, ceil_div(count, 256), 256, 0, st, d_out, count, d_pows.p, log_count);
    ZK_CUDA(cudaStreamSynchronize(st));
}
void ntt_clear_cache() {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    g_tw.clear();
}

static void launch_pass(const NttPass& P, size_t batch, cudaStream_t st) {
    unsigned total = 1u << (P.log_m + P.log_C);
    if (total < 8) total = 8;  // phys() permutes within aligned groups of 8
    size_t smem = (size_t)total * 32;
    unsigned threads = total / 8 < 512 ? (total / 8 < 32 ? 32 : total / 8) : 512;
    static std::mutex mu; static std::map<int, size_t> cur;
    {
        std::lock_guard<std::mutex> lk(mu);
        int dev; ZK_CUDA(cudaGetDevice(&dev));
        if (cur[dev] < 200 * 1024) {
            ZK_CUDA(cudaFuncSetAttribute(k_ntt_tile<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ZK_CUDA(cudaFuncSetAttribute(k_ntt_tile<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            cur[dev] = 200 * 1024;
        }
    }
    size_t blocks = batch * P.tiles_per_poly;
    ZK_REQUIRE(blocks < (1ull << 31), "ntt: grid too large");
    KtScope kt(KT_NTT, st);
    if (threads <= 256 && smem <= 72 * 1024) ZK_LAUNCH((k_ntt_tile<256, 2>), (unsigned)blocks, threads, smem, st, P);
    else ZK_LAUNCH((k_ntt_tile<512, 1>), (unsigned)blocks, threads, smem, st, P);
}

static unsigned pick_swz(unsigned log_m, unsigned log_C) {
    unsigned q = log_C >= 5 ? 0 : 5 - log_C;
    if (q > log_m / 2) q = log_m / 2;
    return q;
}

// 2^13 (Shielder's MAX_K) runs as one launch of two-CTA clusters when there are enough polynomials to give every SM a CTA
// (measured, profiles/r02_ab_bench.md: 235.7 -> 230.4 ms per 1024 proofs; for the few transforms of a single proof the two-pass
// path has more CTAs in flight and is faster: 120 polynomials take 175 us as clusters, 72 us as two passes).  ZKGPU_NTT_CLUSTER=0 falls back to the two-pass path everywhere.
static bool ntt_use_cluster() {
    static const bool on = [] { const char* e = getenv("ZKGPU_NTT_CLUSTER"); return e ? atoi(e) != 0 : true; }();
    return on;
}

// CTAs per cluster of the 2^13 transform: 8 (default) or 2 (ZKGPU_NTT_CLUSTER=2)
static size_t ntt_cluster_min_batch() {
    static const size_t v = [] { const char* e = getenv("ZKGPU_NTT_CLUSTER_MIN"); long x = e ? atol(e) : 0; return x > 0 ? (size_t)x : NTT_CLUSTER_MIN_BATCH; }();
    return v;
}
static unsigned ntt_cluster_size() {
    static const unsigned r = [] { const char* e = getenv("ZKGPU_NTT_CLUSTER"); return (e && atoi(e) == 2) ? 2u : 8u; }();
    return r;
}

size_t ntt_scratch_elems(unsigned log_n, size_t batch) {
    if (log_n == NTT_CLUSTER_LOG && ntt_use_cluster() && batch >= ntt_cluster_min_batch()) return 0;
    return log_n > NTT_SINGLE_PASS_MAX_LOG ? batch << log_n : 0;
}

void ntt_run(const NttJob& J, cudaStream_t st) {
    const unsigned log_N = J.log_n;
    ZK_REQUIRE(log_N <= 26, "ntt: log_n too large");
    const size_t N = (size_t)1 << log_N;
    if (log_N == 0) {
        ZK_REQUIRE(J.in == J.out, "ntt: size-1 transform must be in place");
        return;
    }
    const fr_t* tw = ntt_twiddles(log_N, J.omega, st);
    NttPass P;
    memset(&P, 0, sizeof P);
    P.tw = tw; P.log_N = log_N;
    P.in_poly_stride = J.in_broadcast ? 0 : (J.in_stride ? J.in_stride : N);
    P.pre_table = J.pre_table; P.pre_count = J.pre_count ? J.pre_count : 1;
    P.out_poly_stride = J.out_stride ? J.out_stride : N;
    P.in_inner = J.in_inner ? J.in_inner : ~(size_t)0; P.in_outer_stride = J.in_outer_stride;
    P.out_inner = J.out_inner ? J.out_inner : ~(size_t)0; P.out_outer_stride = J.out_outer_stride;
    P.in_valid = J.in_valid ? J.in_valid : N;
    fr_t one = fe_one<FrTag>();
    P.cs1 = J.pre_coset ? J.cs1 : (J.post_coset ? J.cs1 : one);
    P.cs2 = J.pre_coset ? J.cs2 : (J.post_coset ? J.cs2 : one);
    if (log_N <= NTT_SINGLE_PASS_MAX_LOG) {
        P.in = J.in; P.out = J.out;
        P.log_m = log_N; P.log_C = 0; P.swz_q = pick_swz(log_N, 0);
        P.tiles_per_poly = 1;
        P.in_sj = 1; P.in_sc = 0; P.out_sj = 1; P.out_sc = 0;
        P.c_fastest_in = 0; P.c_fastest_out = 0;
        P.pre_coset = J.pre_coset; P.post_coset = J.post_coset;
        P.has_scale = J.has_scale; P.scale = J.scale;
        P.first_window_trivial = (log_N >= 3 && P.in_valid * 8 <= N) ? 1 : 0;
        launch_pass(P, J.batch, st);
        return;
    }
    if (log_N == NTT_CLUSTER_LOG && ntt_use_cluster() && J.batch >= ntt_cluster_min_batch()) {
        const unsigned log_r = ntt_cluster_size() == 2 ? 1 : 3;
        P.in = J.in; P.out = J.out;
        P.log_m = log_N - log_r; P.log_C = 0; P.swz_q = pick_swz(P.log_m, 0);
        P.tiles_per_poly = 1;
        P.pre_coset = J.pre_coset; P.post_coset = J.post_coset;
        P.has_scale = J.has_scale; P.scale = J.scale;
        P.first_window_trivial = (P.in_valid * 8 <= N) ? 1 : 0;
        ZK_REQUIRE((J.batch << log_r) < (1ull << 31), "ntt: grid too large");
        KtScope kt(KT_NTT, st);
        if (log_r == 1) {
            static DeviceOnce once;
            once.run([] { ZK_CUDA(cudaFuncSetAttribute(k_ntt_cluster2, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024)); });
            ZK_LAUNCH(k_ntt_cluster2, (unsigned)(2 * J.batch), 512, (size_t)128 * 1024, st, P);
        } else {
            ZK_LAUNCH(k_ntt_cluster8, (unsigned)(8 * J.batch), 128, (size_t)32 * 1024, st, P);
        }
        return;
    }
    ZK_REQUIRE(J.scratch != nullptr, "ntt: scratch buffer required for two-pass transforms");
    const unsigned log_n1 = log_N / 2, log_n2 = log_N - log_n1;
    const size_t n1 = (size_t)1 << log_n1, n2 = (size_t)1 << log_n2;
    ZK_REQUIRE(log_n2 <= 12, "ntt: log_n too large for two passes");
    // tile width: 8 columns (256 B segments) unless the tile would exceed ~128 KB
    auto pick_c = [](unsigned log_m) { unsigned lc = 3; while (log_m + lc > 12) --lc; return lc; };
    // pass 1: columns, in -> scratch (same layout), twiddle by w_N^(i2*k1)
    {
        NttPass A = P;
        A.in = J.in; A.out = J.scratch;
        A.out_poly_stride = N; A.out_inner = ~(size_t)0; A.out_outer_stride = 0;
        A.log_m = log_n1; A.log_C = pick_c(log_n1); A.swz_q = pick_swz(A.log_m, A.log_C);
        unsigned C = 1u << A.log_C;
        A.tiles_per_poly = (unsigned)(n2 >> A.log_C);
        A.in_tile_stride = C; A.in_sj = n2; A.in_sc = 1;
        A.out_tile_stride = C; A.out_sj = n2; A.out_sc = 1;
        A.c_fastest_in = 1; A.c_fastest_out = 1;
        A.post_twiddle = 1;
        A.pre_coset = J.pre_coset; A.post_coset = 0; A.has_scale = 0;
        // column j of pass 1 holds polynomial indices c + n2*j: zero for j >= in_valid / n2
        A.first_window_trivial = (log_n1 >= 3 && P.in_valid * 8 <= N && (P.in_valid % n2) == 0) ? 1 : 0;
        launch_pass(A, J.batch, st);
    }
    // pass 2: rows of the scratch, written transposed to out
    {
        NttPass B = P;
        B.in = J.scratch; B.out = J.out;
        B.in_poly_stride = N; B.in_valid = N; B.in_inner = ~(size_t)0; B.in_outer_stride = 0;
        B.pre_table = nullptr;
        B.log_m = log_n2; B.log_C = pick_c(log_n2); B.swz_q = pick_swz(B.log_m, B.log_C);
        unsigned C = 1u << B.log_C;
        B.tiles_per_poly = (unsigned)(n1 >> B.log_C);
        B.in_tile_stride = (size_t)C * n2; B.in_sj = 1; B.in_sc = n2;
        B.out_tile_stride = C; B.out_sj = n1; B.out_sc = 1;
        B.c_fastest_in = 0; B.c_fastest_out = 1;
        B.post_twiddle = 0;
        B.pre_coset = 0; B.post_coset = J.post_coset; B.has_scale = J.has_scale; B.scale = J.scale;
        // inner transforms of length n2 use w_{n2} = w_N^(n1): handled by the (log_N - s - 1) shift
        launch_pass(B, J.batch, st);
    }
}

}  // namespace zk
