// Lookup-argument kernels of the batched prover (halo2_proofs v0.3.0 plonk/lookup/prover.rs; the verifier-side
// terms are /root/reference/crates/halo2-verifier/src/lib/codegen/evaluator.rs:126-223):
//   k_lookup_compress : theta-compression of the input / table expressions over the Lagrange rows
//   k_lookup_sort     : bitonic sort (canonical integer order = `Fr: Ord`) of one column, one CTA per column
//   k_lookup_permute  : permute_expression_pair — first occurrence of every input value takes its table twin,
//                       repeated rows (taken from the end) receive the left-over table values in ascending order
//   k_lookup_num_den  : factors of the lookup grand product
#include "prover_kernels.cuh"
#include "plonk_types.hpp"
#include <mutex>

namespace zk {

template <class FF, class FA, class FI>
__device__ __forceinline__ fr_t lk_run_expr(const uint32_t* prog, uint32_t pc0, uint32_t pc1, const fr_t* constants, FF fixed_at, FA advice_at, FI inst_at) {
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
            default: stack[sp - 1] = stack[sp - 1] * fe_ldg(constants + arg); break;
        }
    }
    return stack[0];
}

__global__ void __launch_bounds__(128) k_lookup_compress(const LookupCompressArgs a, fr_t* out_in, fr_t* out_tab, size_t B) {
    const size_t n = (size_t)1 << a.k;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * a.lp.L * n) return;
    const size_t row = t & (n - 1);
    const unsigned l = (unsigned)((t >> a.k) % a.lp.L);
    const size_t b = (t >> a.k) / a.lp.L;
    const fr_t* adv = a.adv + b * a.adv_proof_stride;
    const fr_t* inst = a.inst + b * a.inst_proof_stride;
    auto rot = [&](int r) -> size_t { return (row + (size_t)(long)r) & (n - 1); };
    auto fixed_at = [&](uint32_t q) { return fe_ldg(a.fixed_vals + (size_t)a.lp.fix_q[2 * q] * n + rot(a.lp.fix_q[2 * q + 1])); };
    auto advice_at = [&](uint32_t q) { return fe_load(adv + (size_t)a.lp.adv_q[2 * q] * n + rot(a.lp.adv_q[2 * q + 1])); };
    auto inst_at = [&](uint32_t q) { return fe_load(inst + rot(a.lp.inst_q[2 * q + 1])); };
    const fr_t theta = fe_ldg(&a.ch[b].theta);
    const uint32_t e0 = a.lp.lk_off[l], e1 = a.lp.lk_off[l + 1], em = e0 + (e1 - e0) / 2;
    fr_t cin = fr_t::zero(), ctab = fr_t::zero();
    for (uint32_t e = e0; e < em; ++e) cin = cin * theta + lk_run_expr(a.lp.prog, a.lp.expr_off[e], a.lp.expr_off[e + 1], a.lp.constants, fixed_at, advice_at, inst_at);
    for (uint32_t e = em; e < e1; ++e) ctab = ctab * theta + lk_run_expr(a.lp.prog, a.lp.expr_off[e], a.lp.expr_off[e + 1], a.lp.constants, fixed_at, advice_at, inst_at);
    fe_store(out_in + t, cin);
    fe_store(out_tab + t, ctab);
}
void launch_lookup_compress(const LookupCompressArgs& a, fr_t* out_in, fr_t* out_tab, size_t B, cudaStream_t st) {
    size_t total = (B * a.lp.L) << a.k;
    KtScope kt(KT_LOOKUP, st);
    if (total) ZK_LAUNCH(k_lookup_compress, ceil_div(total, 128), 128, 0, st, a, out_in, out_tab, B);
}

// canonical 256-bit comparison, most significant limb first
__device__ __forceinline__ bool lt256(const fr_t& x, const fr_t& y) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (x.l[i] != y.l[i]) return x.l[i] < y.l[i];
    }
    return false;
}
__device__ __forceinline__ fr_t fr_sentinel() {
    fr_t s;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.l[i] = 0xffffffffu;  // > every canonical value
    return s;
}

// one CTA per column: src [n] Montgomery values -> dst [n] canonical values sorted ascending; rows >= usable become
// sentinels (sorted to the end).  Column c < BL is an input column, c >= BL a table column.
__global__ void __launch_bounds__(1024) k_lookup_sort(const fr_t* comp_in, const fr_t* comp_tab, fr_t* sort_a, fr_t* sort_t, unsigned k,
                                                      unsigned usable, size_t BL) {
    const unsigned n = 1u << k;
    const size_t col = blockIdx.x;
    const fr_t* src = (col < BL ? comp_in + col * n : comp_tab + (col - BL) * n);
    fr_t* dst = (col < BL ? sort_a + col * n : sort_t + (col - BL) * n);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) fe_store(dst + i, i < usable ? from_mont(fe_load(src + i)) : fr_sentinel());
    __syncthreads();
    for (unsigned size = 2; size <= n; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                unsigned i = ((t / stride) * (stride << 1)) + (t % stride);   // lower index of the pair
                unsigned j = i + stride;
                bool asc = (i & size) == 0;
                fr_t x = fe_load(dst + i), y = fe_load(dst + j);
                if (lt256(y, x) == asc) { fe_store(dst + i, y); fe_store(dst + j, x); }
            }
            __syncthreads();
        }
    }
}

// block-wide exclusive scan of one uint32 per element over `n` elements held in shared memory `v` (in place);
// returns the total.  blockDim.x threads, each owning a contiguous chunk.
__device__ uint32_t block_exclusive_scan(uint32_t* v, unsigned n, uint32_t* part) {
    const unsigned T = blockDim.x, per = (n + T - 1) / T;
    const unsigned lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint32_t sum = 0;
    for (unsigned i = lo; i < hi; ++i) sum += v[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        uint32_t add = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t total = part[T - 1];
    uint32_t run = part[threadIdx.x] - sum;
    for (unsigned i = lo; i < hi; ++i) { uint32_t x = v[i]; v[i] = run; run += x; }
    __syncthreads();
    return total;
}

// one CTA per (proof, lookup).  sort_a / sort_t hold the sorted canonical columns.
__global__ void __launch_bounds__(1024) k_lookup_permute(const fr_t* sort_a, fr_t* sort_t, fr_t* perm_in, fr_t* perm_tab, unsigned k,
                                                         unsigned usable, unsigned L, int* d_error /*[B], per proof*/) {
    extern __shared__ uint32_t sm[];
    const unsigned n = 1u << k;
    uint32_t* rep = sm;            // [n] 1 if the row repeats the previous input value -> exclusive scan
    uint32_t* keep = sm + n;       // [n] 1 if the sorted table entry is left over        -> exclusive scan
    uint32_t* part = sm + 2 * n;   // [blockDim]
    const fr_t* A = sort_a + (size_t)blockIdx.x * n;
    fr_t* T = sort_t + (size_t)blockIdx.x * n;
    fr_t* pin = perm_in + (size_t)blockIdx.x * n;
    fr_t* ptab = perm_tab + (size_t)blockIdx.x * n;
    for (unsigned r = threadIdx.x; r < n; r += blockDim.x) {
        keep[r] = r < usable ? 1u : 0u;
        rep[r] = (r < usable && r > 0 && fe_load(A + r) == fe_load(A + r - 1)) ? 1u : 0u;
    }
    __syncthreads();
    // every first occurrence removes one instance of its value from the table multiset
    for (unsigned r = threadIdx.x; r < usable; r += blockDim.x) {
        if (rep[r]) continue;
        fr_t v = fe_load(A + r);
        unsigned lo = 0, hi = usable;   // lower_bound of v in T[0, usable)
        while (lo < hi) { unsigned mid = (lo + hi) >> 1; if (lt256(fe_load(T + mid), v)) lo = mid + 1; else hi = mid; }
        if (lo >= usable || !(fe_load(T + lo) == v)) atomicExch(d_error + blockIdx.x / L, 1);
        else keep[lo] = 0;              // distinct values hit distinct positions
    }
    __syncthreads();
    const uint32_t R = block_exclusive_scan(rep, n, part);    // rep[r] = number of repeated rows below r
    block_exclusive_scan(keep, n, part);                      // keep[p] = rank of T[p] among the left-over entries
    // compact the left-over entries in place behind the table (ascending order is preserved): use perm_tab as staging
    for (unsigned p = threadIdx.x; p < usable; p += blockDim.x) {
        bool kept = (p + 1 < n ? keep[p + 1] : R) != keep[p];
        if (kept) fe_store(ptab + keep[p], fe_load(T + p));
    }
    __syncthreads();
    for (unsigned p = threadIdx.x; p < R; p += blockDim.x) fe_store(T + p, fe_load(ptab + p));   // T[0, R) = left-over list
    __syncthreads();
    for (unsigned r = threadIdx.x; r < usable; r += blockDim.x) {
        fr_t a = fe_load(A + r);
        bool repeated = (r + 1 < n ? rep[r + 1] : R) != rep[r];
        // repeated rows are consumed from the end: the j-th repeated row from the end takes left-over entry j
        fr_t s = repeated ? fe_load(T + (R - 1 - rep[r])) : a;
        fe_store(pin + r, to_mont(a));
        fe_store(ptab + r, to_mont(s));
    }
}

void launch_lookup_permute(const fr_t* comp_in, const fr_t* comp_tab, fr_t* perm_in, fr_t* perm_tab, fr_t* sort_a, fr_t* sort_t, unsigned k,
                           size_t usable, size_t BL, unsigned L, int* d_error, cudaStream_t st) {
    if (!BL) return;
    const unsigned n = 1u << k;
    unsigned threads = n / 2 < 1024 ? (n / 2 < 32 ? 32 : n / 2) : 1024;
    KtScope kt(KT_LOOKUP, st);
    ZK_LAUNCH(k_lookup_sort, (unsigned)(2 * BL), threads, 0, st, comp_in, comp_tab, sort_a, sort_t, k, (unsigned)usable, BL);
    size_t smem = ((size_t)2 * n + threads) * sizeof(uint32_t);
    static DeviceOnce attr_once;   // per device; two pipeline workers may arrive here together
    attr_once.run([] { ZK_CUDA(cudaFuncSetAttribute(k_lookup_permute, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); });
    ZK_LAUNCH(k_lookup_permute, (unsigned)BL, threads, smem, st, sort_a, sort_t, perm_in, perm_tab, k, (unsigned)usable, L, d_error);
}

__global__ void __launch_bounds__(128) k_lookup_num_den(const fr_t* comp_in, const fr_t* comp_tab, const fr_t* perm_in, const fr_t* perm_tab,
                                                        const Challenges* ch, fr_t* num, fr_t* den, unsigned k, unsigned L, size_t B) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (B * L) << k) return;
    const size_t b = (t >> k) / L;
    const fr_t beta = fe_ldg(&ch[b].beta), gamma = fe_ldg(&ch[b].gamma);
    fe_store(num + t, (fe_load(comp_in + t) + beta) * (fe_load(comp_tab + t) + gamma));
    fe_store(den + t, (fe_load(perm_in + t) + beta) * (fe_load(perm_tab + t) + gamma));
}
void launch_lookup_num_den(const fr_t* comp_in, const fr_t* comp_tab, const fr_t* perm_in, const fr_t* perm_tab, const Challenges* ch,
                           fr_t* num, fr_t* den, unsigned k, unsigned L, size_t B, cudaStream_t st) {
    size_t total = (B * L) << k;
    if (total) ZK_LAUNCH(k_lookup_num_den, ceil_div(total, 128), 128, 0, st, comp_in, comp_tab, perm_in, perm_tab, ch, num, den, k, L, B);
}

}  // namespace zk
