// K2 — BN254 G1 multi-scalar multiplication on the device.
//
// Replaces halo2curves `best_multiexp` behind `ParamsKZG::{commit,
// commit_lagrange}` (row a4).  Pippenger with signed c-bit digits:
//   1. k_msm_count    : scalars Montgomery -> canonical, recode into signed digits, histogram the
//                       (group, |digit|) keys
//   2. k_msm_scan     : exclusive prefix sum of the histogram (one CTA per MSM)
//   3. k_msm_scatter  : recode again and scatter (point index | sign) into bucket-sorted order
//   4. k_msm_buckets  : one thread per bucket walks its run and accumulates in XYZZ (madd 8M+2S)
//   5. k_msm_reduce   : sum_b b*B_b per group: per-thread running sums over a bucket segment, a small
//                       scalar multiple for the segment base, shared-memory tree across the block
//   6. k_msm_combine  : Horner over windows (plain mode only)
// Fixed-base ("precomp") mode — the proof hot path, where every commitment uses the same SRS bases —
// stores T[w][i] = 2^(c*w)*G_i once at SRS registration so all windows share ONE bucket set: the
// per-MSM cost drops from W*(n madd + 2*nb add) to n*W madd + 2*nb add and step 6 disappears.
// The group sum is independent of accumulation order and every exceptional case of the addition
// formulas is handled, so the affine-normalised result is bit-exact against the CPU oracle.
#include "msm.cuh"
#include "msm_digits.cuh"
#include <mutex>
#include <cstdlib>

namespace zk {

struct MsmDims { unsigned c, W, G, nb; unsigned precomp; unsigned n; unsigned tstride; unsigned heavy; unsigned diff_off; size_t inner, outer_stride; };  // heavy: runs longer than this go to k_msm_heavy  // tstride: table stride (points per window)  // diff_off: MsmPlan::diff_offset

// mixed additions queued for the bucket kernels by the throughput path's digit sort (one atomic per MSM): what the roofline of
// bucket accumulation is computed from, since zero digits and difference mode make it depend on the data (zkgpu_msm_additions)
__device__ unsigned long long g_msm_entries = 0;

__global__ void k_msm_count(const fr_t* __restrict__ scalars, size_t total, MsmDims D, uint32_t* __restrict__ counts) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    size_t m = idx / D.n;
    size_t i_ = idx - m * D.n;
    fr_t s = from_mont(fe_load(scalars + (m / D.inner) * D.outer_stride + (m % D.inner) * D.n + i_));
    uint32_t* cm = counts + m * ((size_t)D.G * D.nb);
    for_each_digit(s.l, D.c, D.W, [&](unsigned w, uint32_t mag, bool) {
        unsigned g = D.precomp ? 0 : w;
        atomicAdd(cm + (size_t)g * D.nb + (mag - 1), 1u);
    });
}

// one CTA per MSM; offsets has K+1 entries per MSM
__global__ void k_msm_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, unsigned K) {
    __shared__ uint32_t part[1024];
    const uint32_t* c = counts + (size_t)blockIdx.x * K;
    uint32_t* o = offsets + (size_t)blockIdx.x * (K + 1);
    unsigned per = (K + blockDim.x - 1) / blockDim.x;
    unsigned lo = threadIdx.x * per, hi = lo + per < K ? lo + per : K;
    uint32_t sum = 0;
    for (unsigned i = lo; i < hi; ++i) sum += c[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread sums
    for (unsigned d = 1; d < blockDim.x; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (unsigned i = lo; i < hi; ++i) { o[i] = run; run += c[i]; }
    if (threadIdx.x == blockDim.x - 1) o[K] = part[threadIdx.x];
}

__global__ void k_msm_scatter(const fr_t* __restrict__ scalars, size_t total, MsmDims D, uint32_t* __restrict__ counts,
                              const uint32_t* __restrict__ offsets, uint32_t* __restrict__ entries) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    size_t m = idx / D.n;
    uint32_t i = (uint32_t)(idx - m * D.n);
    fr_t s = from_mont(fe_load(scalars + (m / D.inner) * D.outer_stride + (m % D.inner) * D.n + i));
    const size_t K = (size_t)D.G * D.nb;
    uint32_t* cm = counts + m * K;
    const uint32_t* om = offsets + m * (K + 1);
    uint32_t* em = entries + m * ((size_t)D.n * D.W);
    for_each_digit(s.l, D.c, D.W, [&](unsigned w, uint32_t mag, bool negative) {
        unsigned g = D.precomp ? 0 : w;
        size_t key = (size_t)g * D.nb + (mag - 1);
        uint32_t pos = atomicSub(cm + key, 1u) - 1;
        uint32_t ref = D.precomp ? w * D.tstride + i : i;
        em[om[key] + pos] = ref | (negative ? 0x80000000u : 0u);
    });
}

// Single-kernel digit sort for MSMs whose bucket table fits in shared memory (the prover's fixed-base MSMs: 4096
// buckets): one CTA per MSM histograms the signed digits, scans, and scatters with SHARED-memory atomics; it also
// emits the bucket order by decreasing run length.  Replaces k_msm_count + k_msm_scan + k_msm_scatter + k_msm_order
// and their global atomics (the scalars are read twice, from L2).
// CT: the window width as a compile-time constant (0 = run-time D.c): with CT the window loop unrolls and the limb index of
// every digit is static, so the scalar stays in registers instead of a dynamically indexed local-memory array.
template <unsigned CT>
__global__ void __launch_bounds__(1024, 1) k_msm_sort_smem(const fr_t* __restrict__ scalars, MsmDims D, uint32_t* __restrict__ offsets,
                                                        uint32_t* __restrict__ entries, uint32_t* __restrict__ order) {
    extern __shared__ uint32_t sm_sort[];
    const unsigned K = D.G * D.nb, T = blockDim.x;
    const unsigned c = CT ? CT : D.c, W = CT ? 254 / CT + 1 : D.W;
    uint32_t* cnt = sm_sort;              // [K]    histogram -> cursors
    uint32_t* part = sm_sort + K;         // [T]
    uint32_t* hist = part + T;            // [heavy + 2] run-length histogram for the ordering
    const size_t m = blockIdx.x;
    const fr_t* sc = scalars + (m / D.inner) * D.outer_stride + (m % D.inner) * D.n;
    uint32_t* om = offsets + m * ((size_t)K + 1);
    uint32_t* em = entries + m * ((size_t)D.n * D.W);
    uint32_t* ord = order + m * (size_t)K;
    __shared__ uint32_t nz_cnt[2];
    for (unsigned i = threadIdx.x; i < K; i += T) cnt[i] = 0;
    for (unsigned i = threadIdx.x; i < D.heavy + 2; i += T) hist[i] = 0;
    if (threadIdx.x < 2) nz_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    // Which half of the table: the scalars against the plain bases, or their differences against the prefix sums of the bases,
    // whichever has clearly fewer non-zero terms (Montgomery words are zero / equal exactly when the values are).
    bool diff = false;
    if (D.diff_off) {
        uint32_t ns = 0, nd = 0;
        for (unsigned i = threadIdx.x; i < D.n; i += T) {
            const fr_t a = fe_load(sc + i);
            const fr_t b = i + 1 < D.n ? fe_load(sc + i + 1) : fr_t::zero();
            ns += a.is_zero() ? 0u : 1u;
            nd += a == b ? 0u : 1u;
        }
        for (unsigned o = 16; o; o >>= 1) { ns += __shfl_down_sync(0xffffffffu, ns, o); nd += __shfl_down_sync(0xffffffffu, nd, o); }
        if (lane == 0) { atomicAdd(&nz_cnt[0], ns); atomicAdd(&nz_cnt[1], nd); }
        __syncthreads();
        diff = nz_cnt[1] + nz_cnt[1] / 8 < nz_cnt[0];
    }
    const uint32_t ref_off = diff ? D.diff_off : 0u;
    // Histogram.  The digit loop is warp-uniform (every lane walks all W windows of its scalar, zero digits included), so the
    // lanes can vote: a column whose scalars are all equal — a constant grand product, a selector — sends every lane of a warp
    // to the SAME counter in every window, and 32 same-address shared atomics serialise.  When the voting lanes agree on the
    // key, one lane adds their count (warp-aggregated atomic); otherwise every lane issues its own.
    const uint32_t hmax = 1u << (c - 1);
    for (unsigned base = 0; base < D.n; base += T) {
        const unsigned i = base + threadIdx.x;
        const bool valid = i < D.n;
        bool flip = false;
        fr_t s = fr_t::zero();
        if (valid) s = digit_scalar(fe_load(sc + i), diff && i + 1 < D.n ? fe_load(sc + i + 1) : fr_t::zero(), diff, flip);
        uint32_t carry = 0;
#pragma unroll
        for (unsigned w = 0; w < W; ++w) {
            uint32_t d = get_bits(s.l, w * c, c) + carry, mag;
            if (d > hmax) { carry = 1; mag = (1u << c) - d; } else { carry = 0; mag = d; }
            const bool nz = mag != 0;
            const unsigned key = (D.precomp ? 0 : w) * D.nb + (mag - 1);
            const unsigned am = __ballot_sync(0xffffffffu, nz);
            if (!am) continue;
            const unsigned key0 = __shfl_sync(0xffffffffu, key, __ffs(am) - 1);
            if (__all_sync(0xffffffffu, !nz || key == key0)) { if (lane == (unsigned)__ffs(am) - 1) atomicAdd(&cnt[key0], (uint32_t)__popc(am)); }
            else if (nz) atomicAdd(&cnt[key], 1u);
        }
    }
    __syncthreads();
    // run-length histogram, then exclusive scan of the counts (in place) -> offsets
    const unsigned per = (K + T - 1) / T, lo = threadIdx.x * per, hi = lo + per < K ? lo + per : K;
    uint32_t sum = 0;
    for (unsigned i = lo; i < hi; ++i) { uint32_t c = cnt[i]; sum += c; atomicAdd(&hist[c > D.heavy ? D.heavy + 1 : c], 1u); }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {   // descending exclusive scan of the run-length histogram
        uint32_t pos = 0;
        for (int sz = (int)D.heavy + 1; sz >= 0; --sz) { uint32_t c = hist[sz]; hist[sz] = pos; pos += c; }
    }
    uint32_t run = part[threadIdx.x] - sum;
    __syncthreads();
    for (unsigned i = lo; i < hi; ++i) {
        uint32_t c = cnt[i];
        ord[atomicAdd(&hist[c > D.heavy ? D.heavy + 1 : c], 1u)] = i;
        om[i] = run; cnt[i] = run; run += c;
    }
    if (threadIdx.x == T - 1) { om[K] = part[T - 1]; atomicAdd(&g_msm_entries, (unsigned long long)part[T - 1]); }
    __syncthreads();
    for (unsigned base = 0; base < D.n; base += T) {
        const unsigned i = base + threadIdx.x;
        const bool valid = i < D.n;
        bool flip = false;
        fr_t s = fr_t::zero();
        if (valid) s = digit_scalar(fe_load(sc + i), diff && i + 1 < D.n ? fe_load(sc + i + 1) : fr_t::zero(), diff, flip);
        uint32_t carry = 0;
#pragma unroll
        for (unsigned w = 0; w < W; ++w) {
            uint32_t d = get_bits(s.l, w * c, c) + carry, mag;
            bool negative;
            if (d > hmax) { carry = 1; mag = (1u << c) - d; negative = !flip; } else { carry = 0; mag = d; negative = flip; }
            const bool nz = mag != 0;
            const unsigned key = (D.precomp ? 0 : w) * D.nb + (mag - 1);
            const unsigned am = __ballot_sync(0xffffffffu, nz);
            if (!am) continue;
            const unsigned leader = __ffs(am) - 1;
            const unsigned key0 = __shfl_sync(0xffffffffu, key, leader);
            uint32_t pos = 0;
            if (__all_sync(0xffffffffu, !nz || key == key0)) {
                if (lane == leader) pos = atomicAdd(&cnt[key0], (uint32_t)__popc(am));
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(am & ((1u << lane) - 1u));
            } else if (nz) pos = atomicAdd(&cnt[key], 1u);
            if (nz) {
                uint32_t ref = (D.precomp ? w * D.tstride + i : i) + ref_off;
                em[pos] = ref | (negative ? 0x80000000u : 0u);
            }
        }
    }
}

// Buckets of one MSM ordered by decreasing run length (counting sort on the length, one CTA per MSM), so the
// 32 lanes of a warp in k_msm_buckets walk runs of (nearly) equal length instead of idling behind the longest.
#define ZK_HEAVY_MIN 192
#define ZK_HEAVY_MAX 8190
__global__ void __launch_bounds__(1024) k_msm_order(const uint32_t* __restrict__ offsets, uint32_t* __restrict__ order, unsigned K, unsigned ZK_HEAVY) {
    extern __shared__ uint32_t hist[];  // ZK_HEAVY + 2 bins
    const uint32_t* o = offsets + (size_t)blockIdx.x * (K + 1);
    uint32_t* ord = order + (size_t)blockIdx.x * K;
    for (unsigned i = threadIdx.x; i < ZK_HEAVY + 2; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (unsigned key = threadIdx.x; key < K; key += blockDim.x) {
        uint32_t sz = o[key + 1] - o[key];
        atomicAdd(&hist[sz > ZK_HEAVY ? ZK_HEAVY + 1 : sz], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t pos = 0;
        for (int sz = (int)ZK_HEAVY + 1; sz >= 0; --sz) { uint32_t c = hist[sz]; hist[sz] = pos; pos += c; }
    }
    __syncthreads();
    for (unsigned key = threadIdx.x; key < K; key += blockDim.x) {
        uint32_t sz = o[key + 1] - o[key];
        ord[atomicAdd(&hist[sz > ZK_HEAVY ? ZK_HEAVY + 1 : sz], 1u)] = key;
    }
}

// Buckets holding more than ZK_HEAVY entries (repeated scalars: selector-like 0/1 columns, constant grand
// products) are deferred to k_msm_heavy so that one thread never walks a long run alone.
__global__ void __launch_bounds__(128) k_msm_buckets(const g1_affine_t* __restrict__ bases, const uint32_t* __restrict__ offsets,
                                                     const uint32_t* __restrict__ entries, MsmDims D, size_t M,
                                                     g1_xyzz_t* __restrict__ buckets, const uint32_t* __restrict__ order,
                                                     uint32_t* __restrict__ heavy_count, uint64_t* __restrict__ heavy_list) {
    const size_t K = (size_t)D.G * D.nb;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * K) return;
    size_t m = idx / K;
    const size_t key = order[idx];
    idx = m * K + key;
    const uint32_t* om = offsets + m * (K + 1);
    const uint32_t* em = entries + m * ((size_t)D.n * D.W);
    uint32_t b = om[key], e = om[key + 1];
    if (e - b > D.heavy) {
        heavy_list[atomicAdd(heavy_count, 1u)] = idx;
        return;
    }
    g1_xyzz_t acc = g1_xyzz_t::identity();
    if (b < e) {
        // software pipeline: the next point is in flight while the current one is added
        uint32_t ref = em[b];
        const g1_affine_t* p = bases + (ref & 0x7fffffffu);
        g1_affine_t q;
        q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
        for (uint32_t t = b + 1; t < e; ++t) {
            uint32_t ref_n = em[t];
            const g1_affine_t* pn = bases + (ref_n & 0x7fffffffu);
            g1_affine_t qn;
            qn.x = fe_ldg(&pn->x); qn.y = fe_ldg(&pn->y);
            xyzz_madd(acc, q, (ref >> 31) != 0);
            q = qn; ref = ref_n;
        }
        xyzz_madd(acc, q, (ref >> 31) != 0);
    }
    g1_xyzz_t* o = buckets + idx;
    fe_store(&o->x, acc.x); fe_store(&o->y, acc.y); fe_store(&o->zz, acc.zz); fe_store(&o->zzz, acc.zzz);
}

__device__ __forceinline__ g1_xyzz_t xyzz_load(const g1_xyzz_t* p) {
    g1_xyzz_t r;
    r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz);
    return r;
}
__device__ __forceinline__ void xyzz_store(g1_xyzz_t* p, const g1_xyzz_t& v) {
    fe_store(&p->x, v.x); fe_store(&p->y, v.y); fe_store(&p->zz, v.zz); fe_store(&p->zzz, v.zzz);
}

// One CTA per heavy bucket (grid-stride over the worklist): every thread accumulates a strided slice of the
// run, then a shared-memory tree folds the 128 partial sums.
__global__ void __launch_bounds__(128) k_msm_heavy(const g1_affine_t* __restrict__ bases, const uint32_t* __restrict__ offsets,
                                                   const uint32_t* __restrict__ entries, MsmDims D, g1_xyzz_t* __restrict__ buckets,
                                                   const uint32_t* __restrict__ heavy_count, const uint64_t* __restrict__ heavy_list) {
    __shared__ g1_xyzz_t part[128];
    const size_t K = (size_t)D.G * D.nb;
    const uint32_t count = *heavy_count;
    for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
        const size_t idx = heavy_list[h];
        const size_t m = idx / K, key = idx - m * K;
        const uint32_t* om = offsets + m * (K + 1);
        const uint32_t* em = entries + m * ((size_t)D.n * D.W);
        const uint32_t b = om[key], e = om[key + 1];
        g1_xyzz_t acc = g1_xyzz_t::identity();
        for (uint32_t t = b + threadIdx.x; t < e; t += blockDim.x) {
            uint32_t ref = em[t];
            const g1_affine_t* p = bases + (ref & 0x7fffffffu);
            g1_affine_t q;
            q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
            xyzz_madd(acc, q, (ref >> 31) != 0);
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        for (unsigned s = blockDim.x >> 1; s > 0; s >>= 1) {
            if (threadIdx.x < s) part[threadIdx.x] = xyzz_add(part[threadIdx.x], part[threadIdx.x + s]);
            __syncthreads();
        }
        if (threadIdx.x == 0) xyzz_store(buckets + idx, part[0]);
        __syncthreads();
    }
}

// Latency variant for small batches (single proofs): ZK_SPLIT threads share one bucket, each accumulating a strided
// slice of the run, then fold through shared memory — 8x shorter dependency chains, more CTAs in flight.
#define ZK_SPLIT 8
__global__ void __launch_bounds__(128) k_msm_buckets_split(const g1_affine_t* __restrict__ bases, const uint32_t* __restrict__ offsets,
                                                           const uint32_t* __restrict__ entries, MsmDims D, size_t M,
                                                           g1_xyzz_t* __restrict__ buckets, uint32_t* __restrict__ heavy_count,
                                                           uint64_t* __restrict__ heavy_list) {
    __shared__ g1_xyzz_t part[128];
    const size_t K = (size_t)D.G * D.nb;
    const unsigned lane = threadIdx.x % ZK_SPLIT;
    size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZK_SPLIT;
    const bool live = idx < M * K;
    g1_xyzz_t acc = g1_xyzz_t::identity();
    bool heavy = false;
    if (live) {
        size_t m = idx / K, key = idx - m * K;
        const uint32_t* om = offsets + m * (K + 1);
        const uint32_t* em = entries + m * ((size_t)D.n * D.W);
        uint32_t b = om[key], e = om[key + 1];
        heavy = e - b > D.heavy;
        if (heavy) { if (lane == 0) heavy_list[atomicAdd(heavy_count, 1u)] = idx; }
        else
            for (uint32_t t = b + lane; t < e; t += ZK_SPLIT) {
                uint32_t ref = em[t];
                const g1_affine_t* p = bases + (ref & 0x7fffffffu);
                g1_affine_t q;
                q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
                xyzz_madd(acc, q, (ref >> 31) != 0);
            }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned s = ZK_SPLIT >> 1; s > 0; s >>= 1) {
        if (lane < s) part[threadIdx.x] = xyzz_add(part[threadIdx.x], part[threadIdx.x + s]);
        __syncthreads();
    }
    if (live && !heavy && lane == 0) xyzz_store(buckets + idx, part[threadIdx.x]);
}

// block per (m, g); T = blockDim.x threads, each owns L = nb/T consecutive buckets (L a power of two).
//   per thread:  S_t = sum_j B[lo+j],  W_t = sum_j (j+1) B[lo+j]   (running sums: 2L additions)
//   group sum    sum_t (W_t + lo_t S_t),  lo_t = t L,   and   sum_t t S_t = sum_{u>=1} sfx[u]  with sfx the inclusive suffix sums of S
// so the weights cost one suffix scan (log2 T additions, every lane busy, no data-dependent branches) and log2 L doublings,
// instead of a per-thread double-and-add by lo_t whose add/skip pattern diverges inside every warp.
#define ZK_REDUCE_T 128
#define ZK_REDUCE_T_LAT 512   // few MSMs in flight: more, shorter segments per bucket group
__global__ void __launch_bounds__(ZK_REDUCE_T_LAT) k_msm_reduce(const g1_xyzz_t* __restrict__ buckets, MsmDims D,
                                                                g1_xyzz_t* __restrict__ groups) {
    extern __shared__ uint4 reduce_smem[];
    g1_xyzz_t* part = reinterpret_cast<g1_xyzz_t*>(reduce_smem);
    const unsigned T = blockDim.x, t = threadIdx.x;
    const unsigned L = D.nb / T;  // host guarantees T | nb
    const g1_xyzz_t* B = buckets + (size_t)blockIdx.x * D.nb;
    unsigned lo = t * L;  // bucket index lo has weight lo+1
    g1_xyzz_t run = g1_xyzz_t::identity(), acc = g1_xyzz_t::identity();
    for (unsigned j = L; j-- > 0;) {
        run = xyzz_add(run, xyzz_load(B + lo + j));
        acc = xyzz_add(acc, run);
    }
    // inclusive suffix scan of the segment sums
    part[t] = run;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        const bool has = t + d < T;
        g1_xyzz_t other;
        if (has) other = part[t + d];
        __syncthreads();
        if (has) { run = xyzz_add(run, other); part[t] = run; }
        __syncthreads();
    }
    if (t >= 1) {
        for (unsigned l = 1; l < L; l <<= 1) run = xyzz_dbl(run);
        acc = xyzz_add(acc, run);
    }
    part[t] = acc;
    __syncthreads();
    for (unsigned s = T >> 1; s > 0; s >>= 1) {
        if (t < s) part[t] = xyzz_add(part[t], part[t + s]);
        __syncthreads();
    }
    if (t == 0) xyzz_store(groups + blockIdx.x, part[0]);
}

// Latency variant of the bucket reduction for a handful of MSMs: ZK_REDUCE_R CTAs of 128 threads share one bucket
// group (one warp per scheduler, so the dependent additions of a thread run at single-warp latency, on R SMs), the
// per-thread weights come from a suffix scan instead of a per-thread small scalar multiplication, and a second
// tiny kernel folds the R partial sums.
#define ZK_REDUCE_R 8
__global__ void __launch_bounds__(128) k_msm_reduce_multi(const g1_xyzz_t* __restrict__ buckets, MsmDims D, g1_xyzz_t* __restrict__ partial) {
    __shared__ g1_xyzz_t sfx[128];   // suffix sums of the segment sums
    __shared__ g1_xyzz_t wsum[128];  // tree over the weighted segment sums
    const unsigned T = blockDim.x, t = threadIdx.x;
    const unsigned grp = blockIdx.x / ZK_REDUCE_R, c = blockIdx.x % ZK_REDUCE_R;
    const unsigned seg = D.nb / ZK_REDUCE_R;     // buckets of this CTA
    const unsigned L = seg / T;                  // buckets per thread (host guarantees divisibility, L a power of two)
    const g1_xyzz_t* B = buckets + (size_t)grp * D.nb + (size_t)c * seg;
    const unsigned lo = t * L;
    g1_xyzz_t run = g1_xyzz_t::identity(), acc = g1_xyzz_t::identity();
    for (unsigned j = L; j-- > 0;) {
        run = xyzz_add(run, xyzz_load(B + lo + j));
        acc = xyzz_add(acc, run);
    }
    // inclusive suffix scan: sfx[t] = sum_{u >= t} S_u
    sfx[t] = run;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {
        g1_xyzz_t other = g1_xyzz_t::identity();
        const bool has = t + d < T;
        if (has) other = sfx[t + d];
        __syncthreads();
        if (has) sfx[t] = xyzz_add(sfx[t], other);
        __syncthreads();
    }
    // sum_t (W_t + t L S_t) = sum_t W_t + L * sum_{u >= 1} sfx[u]
    g1_xyzz_t v = acc;
    if (t >= 1) {
        g1_xyzz_t x = sfx[t];
        for (unsigned l = 1; l < L; l <<= 1) x = xyzz_dbl(x);
        v = xyzz_add(v, x);
    }
    wsum[t] = v;
    __syncthreads();
    for (unsigned s2 = T >> 1; s2 > 0; s2 >>= 1) {
        if (t < s2) wsum[t] = xyzz_add(wsum[t], wsum[t + s2]);
        __syncthreads();
    }
    if (t == 0) {
        // bucket b of this CTA has global weight c*seg + b + 1: add (c * seg) * (sum of the CTA's buckets)
        g1_xyzz_t r = wsum[0];
        if (c) r = xyzz_add(r, xyzz_mul_small(sfx[0], c * seg));
        xyzz_store(partial + blockIdx.x, r);
    }
}
__global__ void k_msm_reduce_fold(const g1_xyzz_t* __restrict__ partial, g1_xyzz_t* __restrict__ groups, unsigned count) {
    unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    g1_xyzz_t acc = xyzz_load(partial + (size_t)g * ZK_REDUCE_R);
    for (unsigned c = 1; c < ZK_REDUCE_R; ++c) acc = xyzz_add(acc, xyzz_load(partial + (size_t)g * ZK_REDUCE_R + c));
    xyzz_store(groups + g, acc);
}

// ---------------------------------------------------------------------------------------------
// Latency path: a handful of fixed-base MSMs (the commitments of ONE proof).
// The throughput kernels above give every bucket one thread and every bucket group one CTA; with a few MSMs in flight that is a
// few long dependent chains on an otherwise idle GPU (round 1: 6 commitment rounds x ~1 ms of a 6.4 ms proof).  Here the SRS
// carries a second, NARROW-window table (c = 10: 512 buckets, 26 windows), so that
//   * every bucket holds hundreds of points and is accumulated by TPB threads (strided slices + a shared-memory tree), and
//   * the bucket reduction  sum_b (b+1) B_b = sum_b sfx[b]  (sfx = inclusive suffix sums) is ONE CTA per MSM: a suffix scan and a
//     tree, 2 log2(nb) dependent additions instead of ~45, followed by the affine normalisation in the same kernel.
// The digit sort aggregates its histogram in shared memory per CTA (512 counters would serialise ~400 global atomics each).
// MSM m reads its scalars at sc[m] and its points from table + (basis bit m) * table_stride, so commitments over g and over
// g_lagrange share one launch group.
// ---------------------------------------------------------------------------------------------
struct LatArgs {
    const fr_t* sc[ZK_LAT_MAX_M];
    uint32_t basis_mask;
    uint32_t table_stride;   // points between the two bases' tables
};

__global__ void __launch_bounds__(256) k_lat_count(LatArgs A, MsmDims D, uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t lat_sm[];   // [nb]
    const unsigned m = blockIdx.y;   // grid (chunks of 256 scalars, M)
    for (unsigned b = threadIdx.x; b < D.nb; b += blockDim.x) lat_sm[b] = 0;
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.n) {
        fr_t s = from_mont(fe_load(A.sc[m] + i));
        for_each_digit(s.l, D.c, D.W, [&](unsigned, uint32_t mag, bool) { atomicAdd(&lat_sm[mag - 1], 1u); });
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < D.nb; b += blockDim.x)
        if (lat_sm[b]) atomicAdd(counts + (size_t)m * D.nb + b, lat_sm[b]);
}
// cursors[m][b] start at offsets[m][b]; every CTA reserves a range per bucket and fills it from shared-memory counters
__global__ void __launch_bounds__(256) k_lat_scatter(LatArgs A, MsmDims D, uint32_t* __restrict__ cursors, uint32_t* __restrict__ entries) {
    extern __shared__ uint32_t lat_sm[];   // [nb] counts -> bases, [nb] local cursors
    uint32_t* base = lat_sm; uint32_t* local = lat_sm + D.nb;
    const unsigned m = blockIdx.y;
    for (unsigned b = threadIdx.x; b < D.nb; b += blockDim.x) { base[b] = 0; local[b] = 0; }
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    fr_t s = fr_t::zero();
    if (i < D.n) {
        s = from_mont(fe_load(A.sc[m] + i));
        for_each_digit(s.l, D.c, D.W, [&](unsigned, uint32_t mag, bool) { atomicAdd(&base[mag - 1], 1u); });
    }
    __syncthreads();
    for (unsigned b = threadIdx.x; b < D.nb; b += blockDim.x)
        if (base[b]) base[b] = atomicAdd(cursors + (size_t)m * (D.nb + 1) + b, base[b]);
    __syncthreads();
    if (i < D.n) {
        uint32_t* em = entries + (size_t)m * ((size_t)D.n * D.W);
        const uint32_t tb = ((A.basis_mask >> m) & 1u) * A.table_stride;
        for_each_digit(s.l, D.c, D.W, [&](unsigned w, uint32_t mag, bool negative) {
            uint32_t pos = base[mag - 1] + atomicAdd(&local[mag - 1], 1u);
            em[pos] = (tb + w * D.tstride + i) | (negative ? 0x80000000u : 0u);
        });
    }
}
__global__ void k_lat_init_cursors(const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursors, size_t total) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) cursors[t] = offsets[t];
}

// TPB threads per bucket, 128 / TPB buckets per CTA
#define ZK_LAT_HEAVY_PER_THREAD 48   // a bucket with more entries per thread than this goes to k_lat_heavy (512 threads per bucket)
template <unsigned TPB>
__global__ void __launch_bounds__(128, 4) k_lat_buckets(const g1_affine_t* __restrict__ table, const uint32_t* __restrict__ offsets,
                                                     const uint32_t* __restrict__ entries, MsmDims D, size_t M, g1_xyzz_t* __restrict__ buckets,
                                                     uint32_t* __restrict__ heavy_count, uint64_t* __restrict__ heavy_list) {
    __shared__ g1_xyzz_t part[128];
    const unsigned lane = threadIdx.x % TPB;
    const size_t idx = ((size_t)blockIdx.x * 128 + threadIdx.x) / TPB;   // (m, bucket)
    bool live = idx < M * D.nb;
    g1_xyzz_t acc = g1_xyzz_t::identity();
    if (live) {
        const size_t m = idx / D.nb, key = idx - m * D.nb;
        const uint32_t* om = offsets + m * ((size_t)D.nb + 1);
        const uint32_t* em = entries + m * ((size_t)D.n * D.W);
        const uint32_t b = om[key], e = om[key + 1];
        const uint32_t heavy = D.heavy > ZK_LAT_HEAVY_PER_THREAD * TPB ? D.heavy : ZK_LAT_HEAVY_PER_THREAD * TPB;   // D.heavy: 4x the mean run
        if (e - b > heavy) {   // skewed scalars (a constant column puts all n points of a window in one bucket)
            if (lane == 0) heavy_list[atomicAdd(heavy_count, 1u)] = idx;
            live = false;
        } else
            for (uint32_t t = b + lane; t < e; t += TPB) {
                const uint32_t ref = em[t];
                const g1_affine_t* p = table + (ref & 0x7fffffffu);
                g1_affine_t q;
                q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
                xyzz_madd(acc, q, (ref >> 31) != 0);
            }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned s2 = TPB >> 1; s2 > 0; s2 >>= 1) {
        if (lane < s2) part[threadIdx.x] = xyzz_add(part[threadIdx.x], part[threadIdx.x + s2]);
        __syncthreads();
    }
    if (live && lane == 0) xyzz_store(buckets + idx, part[threadIdx.x]);
}

// heavy buckets of the latency path: ZK_LAT_HEAVY_SPLIT CTAs of 512 threads share one bucket (strided slices of its run); the last
// CTA of a bucket to finish folds the partial sums.  `done` counts finished CTAs per heavy bucket (zeroed by the host).
#define ZK_LAT_HEAVY_SPLIT 4
__global__ void __launch_bounds__(512) k_lat_heavy(const g1_affine_t* __restrict__ table, const uint32_t* __restrict__ offsets,
                                                   const uint32_t* __restrict__ entries, MsmDims D, g1_xyzz_t* __restrict__ buckets,
                                                   const uint32_t* __restrict__ heavy_count, const uint64_t* __restrict__ heavy_list,
                                                   g1_xyzz_t* __restrict__ partial, uint32_t* __restrict__ done) {
    extern __shared__ uint4 lat_heavy_sm[];
    g1_xyzz_t* part = reinterpret_cast<g1_xyzz_t*>(lat_heavy_sm);
    __shared__ uint32_t last;
    const uint32_t count = *heavy_count;
    const unsigned slice = blockIdx.x % ZK_LAT_HEAVY_SPLIT, stride = blockDim.x * ZK_LAT_HEAVY_SPLIT;
    for (uint32_t h = blockIdx.x / ZK_LAT_HEAVY_SPLIT; h < count; h += gridDim.x / ZK_LAT_HEAVY_SPLIT) {
        const size_t idx = heavy_list[h];
        const size_t m = idx / D.nb, key = idx - m * D.nb;
        const uint32_t* om = offsets + m * ((size_t)D.nb + 1);
        const uint32_t* em = entries + m * ((size_t)D.n * D.W);
        const uint32_t b = om[key], e = om[key + 1];
        g1_xyzz_t acc = g1_xyzz_t::identity();
        for (uint32_t t = b + slice * blockDim.x + threadIdx.x; t < e; t += stride) {
            const uint32_t ref = em[t];
            const g1_affine_t* p = table + (ref & 0x7fffffffu);
            g1_affine_t q;
            q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
            xyzz_madd(acc, q, (ref >> 31) != 0);
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        for (unsigned s2 = blockDim.x >> 1; s2 > 0; s2 >>= 1) {
            if (threadIdx.x < s2) part[threadIdx.x] = xyzz_add(part[threadIdx.x], part[threadIdx.x + s2]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            xyzz_store(partial + (size_t)h * ZK_LAT_HEAVY_SPLIT + slice, part[0]);
            __threadfence();
            last = atomicAdd(done + h, 1u) == ZK_LAT_HEAVY_SPLIT - 1;
            if (last) {
                __threadfence();
                g1_xyzz_t sum = g1_xyzz_t::identity();
                for (unsigned j = 0; j < ZK_LAT_HEAVY_SPLIT; ++j) sum = xyzz_add(sum, xyzz_load(partial + (size_t)h * ZK_LAT_HEAVY_SPLIT + j));
                xyzz_store(buckets + idx, sum);
            }
        }
        __syncthreads();
    }
}

// Bucket reduction of the latency path, two levels so that no SM has to push hundreds of point additions through its own
// IMAD pipe (one 512-thread CTA per MSM took 265 us, pipe-bound on a single SM):
//   A  one WARP per 32 consecutive buckets (grid: nb/32 x M): S = sum B, V = sum (lane+1) B as the sum of the inclusive suffix sums
//   B  one small CTA per MSM over the nb/32 pairs: total = sum_j V_j + 32 sum_j j S_j, the second sum again as suffix sums of S;
//      then the affine normalisation
__global__ void __launch_bounds__(32) k_lat_reduce_a(const g1_xyzz_t* __restrict__ buckets, MsmDims D, g1_xyzz_t* __restrict__ pairs) {
    __shared__ g1_xyzz_t part[32];
    const unsigned t = threadIdx.x, chunk = blockIdx.x, m = blockIdx.y;
    g1_xyzz_t v = xyzz_load(buckets + (size_t)m * D.nb + chunk * 32 + t);
    part[t] = v;
    __syncwarp();
    for (unsigned d = 1; d < 32; d <<= 1) {
        const bool has = t + d < 32;
        g1_xyzz_t other;
        if (has) other = part[t + d];
        __syncwarp();
        if (has) { v = xyzz_add(v, other); part[t] = v; }
        __syncwarp();
    }
    g1_xyzz_t* out = pairs + ((size_t)m * (D.nb / 32) + chunk) * 2;
    if (t == 0) xyzz_store(out, v);   // S: sfx[0]
    for (unsigned s2 = 16; s2 > 0; s2 >>= 1) {
        if (t < s2) part[t] = xyzz_add(part[t], part[t + s2]);
        __syncwarp();
    }
    if (t == 0) xyzz_store(out + 1, part[0]);
}
__global__ void __launch_bounds__(32) k_lat_reduce_b(const g1_xyzz_t* __restrict__ pairs, MsmDims D, g1_affine_t* __restrict__ out) {
    __shared__ g1_xyzz_t part[32];
    const unsigned T = D.nb / 32, t = threadIdx.x;   // T <= 16 pairs; blockDim = 32, lanes >= T idle
    const g1_xyzz_t* pr = pairs + (size_t)blockIdx.x * T * 2;
    g1_xyzz_t s = t < T ? xyzz_load(pr + 2 * t) : g1_xyzz_t::identity();
    part[t] = s;
    __syncwarp();
    for (unsigned d = 1; d < T; d <<= 1) {
        const bool has = t + d < T;
        g1_xyzz_t other;
        if (has) other = part[t + d];
        __syncwarp();
        if (has) { s = xyzz_add(s, other); part[t] = s; }
        __syncwarp();
    }
    g1_xyzz_t v = t < T ? xyzz_load(pr + 2 * t + 1) : g1_xyzz_t::identity();
    if (t >= 1 && t < T) {
        for (unsigned l = 1; l < 32; l <<= 1) s = xyzz_dbl(s);   // 32 * sfxS[t]
        v = xyzz_add(v, s);
    }
    part[t] = v;
    __syncwarp();
    for (unsigned s2 = 16; s2 > 0; s2 >>= 1) {
        if (t < s2) part[t] = xyzz_add(part[t], part[t + s2]);
        __syncwarp();
    }
    if (t == 0) {
        g1_affine_t a = xyzz_to_affine(part[0]);
        fe_store(&out[blockIdx.x].x, a.x); fe_store(&out[blockIdx.x].y, a.y);
    }
}

// ---------------------------------------------------------------------------------------------
// Direct path: the commitments of one proof WITHOUT buckets.
// With a few MSMs in flight the bucket method pays for its structure, not for its additions: a digit sort, bucket folds and a
// bucket reduction of ~25 dependent point additions per round (~0.35 ms for one MSM).  The SRS bases are fixed, so the table can
// hold every multiple a digit can ask for: D[w][i][d-1] = d * 2^(c w) * G_i for d = 1 .. 2^(c-1) (c = 8: 32 windows x n x 128
// points = 2.1 GB per basis at n = 2^13; HBM has 180 GB).  An MSM is then the plain sum of n * W table entries
//     sum_i s_i G_i = sum_{i,w} sign(d_iw) * D[w][i][|d_iw| - 1]
// — no sort, no buckets, no reduction, nothing that a constant column can skew: every thread adds its share of entries into one
// accumulator (k_direct_sum: units of 128 points x 8 windows, then a shared-memory tree per CTA) and a second small kernel
// folds the CTAs' partial sums and normalises (k_direct_fold).
// Signed digits come from the recoding  d_w = ((s + H) >> c w) & (2^c - 1)) - 2^(c-1),  H = 2^(c-1) * sum_w 2^(c w):  every window's
// digit is independent of the others (no carry chain), in [-2^(c-1), 2^(c-1) - 1]; c * W = 256 so the sum is exact mod 2^256 and
// s + H < 2^256 for s < r.
// ---------------------------------------------------------------------------------------------
#define ZK_DIRECT_WG 8           // windows per unit of work
// c = 8: W = 32 windows, D = 128 multiples (2.1 GB per basis at n = 2^13); c = 9: W = 29, D = 256 (3.9 GB per basis, 9 % fewer additions)
__host__ __device__ constexpr unsigned direct_W(unsigned c) { return 254 / c + 1; }
__host__ __device__ constexpr unsigned direct_D(unsigned c) { return 1u << (c - 1); }
// tmp[i * D + d - 1] = d * base[i] (XYZZ), one thread per point
__global__ void __launch_bounds__(64) k_direct_multiples(const g1_affine_t* __restrict__ base, g1_xyzz_t* __restrict__ tmp, unsigned n, unsigned ZK_DIRECT_D) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p;
    p.x = fe_load(&base[i].x); p.y = fe_load(&base[i].y);
    g1_xyzz_t acc = g1_xyzz_t::identity();
    for (unsigned d = 0; d < ZK_DIRECT_D; ++d) {
        xyzz_madd(acc, p, false);
        xyzz_store(tmp + (size_t)i * ZK_DIRECT_D + d, acc);
    }
}
// affine normalisation of the D multiples of one point with ONE inversion (Montgomery's trick over u_d = ZZ_d * ZZZ_d);
// pre[i * D + d] is scratch for the prefix products
__global__ void __launch_bounds__(64) k_direct_normalize(const g1_xyzz_t* __restrict__ tmp, fq_t* __restrict__ pre, g1_affine_t* __restrict__ out, unsigned n,
                                                         unsigned ZK_DIRECT_D) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const g1_xyzz_t* t = tmp + (size_t)i * ZK_DIRECT_D;
    fq_t* pr = pre + (size_t)i * ZK_DIRECT_D;
    g1_affine_t* o = out + (size_t)i * ZK_DIRECT_D;
    const fq_t one = fe_one<FqTag>();
    fq_t acc = one;
    for (unsigned d = 0; d < ZK_DIRECT_D; ++d) {
        fe_store(pr + d, acc);
        fq_t zz = fe_load(&t[d].zz);
        if (!zz.is_zero()) acc = acc * (zz * fe_load(&t[d].zzz));
    }
    fq_t inv = fe_inv(acc);
    for (unsigned d = ZK_DIRECT_D; d-- > 0;) {
        fq_t zz = fe_load(&t[d].zz), zzz = fe_load(&t[d].zzz);
        if (zz.is_zero()) { fe_store(&o[d].x, fq_t::zero()); fe_store(&o[d].y, fq_t::zero()); continue; }   // multiple of the identity
        fq_t tinv = inv * fe_load(pr + d);       // 1 / (ZZ ZZZ)
        inv = inv * (zz * zzz);
        fe_store(&o[d].x, fe_load(&t[d].x) * (zzz * tinv));
        fe_store(&o[d].y, fe_load(&t[d].y) * (zz * tinv));
    }
}

struct DirectArgs {
    const fr_t* sc[ZK_LAT_MAX_M];
    uint32_t basis_mask;
    size_t table_stride;   // points between the two bases' tables
    unsigned n, tstride;   // scalars per MSM; points per window in the table (the SRS size)
};
// A unit of work = 128 points x WG windows (one madd per thread and window).  grid: (ctas_per_msm, M); CTA c of an MSM takes the
// units [c U / C, (c + 1) U / C), U = (n / 128) * (W / WG): the host picks C so that ALL CTAs of the launch are resident at once
// (no partial last wave) and a thread's chain is as long as that allows, which amortises the CTA's fold tree.
template <unsigned CW>
__global__ void __launch_bounds__(128, 4) k_direct_sum(DirectArgs A, const g1_affine_t* __restrict__ table, unsigned units, g1_xyzz_t* __restrict__ partial) {
    __shared__ g1_xyzz_t part[128];
    constexpr unsigned W = direct_W(CW), D = direct_D(CW), groups = (W + ZK_DIRECT_WG - 1) / ZK_DIRECT_WG;
    const unsigned m = blockIdx.y, C = gridDim.x;
    const g1_affine_t* T = table + ((A.basis_mask >> m) & 1u) * A.table_stride;
    g1_xyzz_t acc = g1_xyzz_t::identity();
    const unsigned u0 = (unsigned)(((uint64_t)blockIdx.x * units) / C), u1 = (unsigned)(((uint64_t)(blockIdx.x + 1) * units) / C);
    unsigned cur_chunk = ~0u;
    uint32_t sl[9];   // s + H: nine limbs (c = 9: 29 windows x 9 bits = 261 bits)
#pragma unroll 1
    for (unsigned u = u0; u < u1; ++u) {
        const unsigned chunk = u / groups, wg = u % groups;     // consecutive units of a chunk share the scalar
        const unsigned i = chunk * 128 + threadIdx.x;
        if (i >= A.n) continue;
        if (chunk != cur_chunk) {
            fr_t s = from_mont(fe_load(A.sc[m] + i));
            // s + H, H = 2^(c-1) * sum_w 2^(c w): 0x80 in every byte for c = 8; bits 8, 17, 26, ... for c = 9
            uint32_t carry = 0;
#pragma unroll
            for (int l = 0; l < 9; ++l) {
                uint32_t hl = 0;
                if (CW == 8) hl = l < 8 ? 0x80808080u : 0u;
                else {
#pragma unroll
                    for (unsigned w = 0; w < W; ++w) { const unsigned bit = CW * w + CW - 1; if ((bit >> 5) == (unsigned)l) hl |= 1u << (bit & 31); }
                }
                uint64_t v = (uint64_t)(l < 8 ? s.l[l] : 0u) + hl + carry; sl[l] = (uint32_t)v; carry = (uint32_t)(v >> 32);
            }
            cur_chunk = chunk;
        }
        // the (up to) 8 digits of this window group: CW * 8 bits starting at bit CW * 8 * wg — three limbs cover them (selects, not a
        // dynamically indexed array)
        const unsigned bit0 = CW * ZK_DIRECT_WG * wg, l0 = bit0 >> 5, sh0 = bit0 & 31;
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
        for (unsigned l = 0; l < 9; ++l) {
            if (l == l0) w0 = sl[l];
            if (l == l0 + 1) w1 = sl[l];
            if (l == l0 + 2) w2 = sl[l];
            if (l == l0 + 3) w3 = sl[l];
        }
        const uint64_t lo64 = ((uint64_t)w1 << 32) | w0, hi64 = ((uint64_t)w3 << 32) | w2;
        auto digit = [&](unsigned j) -> int {
            const unsigned shv = sh0 + CW * j;     // < 32 + 72
            const uint64_t v = shv == 0 ? lo64 : shv < 64 ? (lo64 >> shv) | (hi64 << (64 - shv)) : hi64 >> (shv - 64);
            return (int)(v & ((1u << CW) - 1)) - (int)D;
        };
        const unsigned nw = W - wg * ZK_DIRECT_WG < ZK_DIRECT_WG ? W - wg * ZK_DIRECT_WG : ZK_DIRECT_WG;   // windows in this group
        // the table entry of window j + 1 is in flight (a random 64-byte read from a multi-gigabyte table: DRAM latency) while
        // window j's point is added; a zero digit loads entry 0 and skips the addition
        auto entry = [&](unsigned j) -> const g1_affine_t* {
            const int d = digit(j);
            const unsigned mag = d < 0 ? (unsigned)(-d) : (unsigned)d;
            return T + ((size_t)(wg * ZK_DIRECT_WG + j) * A.tstride + i) * D + (mag ? mag - 1 : 0);
        };
        const g1_affine_t* p = entry(0);
        g1_affine_t q;
        q.x = fe_ldg(&p->x); q.y = fe_ldg(&p->y);
#pragma unroll 1
        for (unsigned j = 0; j < nw; ++j) {
            g1_affine_t qn = q;
            if (j + 1 < nw) { const g1_affine_t* pn = entry(j + 1); qn.x = fe_ldg(&pn->x); qn.y = fe_ldg(&pn->y); }
            const int d = digit(j);
            if (d != 0) xyzz_madd(acc, q, d < 0);
            q = qn;
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned s2 = 64; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) part[threadIdx.x] = xyzz_add(part[threadIdx.x], part[threadIdx.x + s2]);
        __syncthreads();
    }
    if (threadIdx.x == 0) xyzz_store(partial + (size_t)m * C + blockIdx.x, part[0]);
}
// one CTA per MSM folds its `count` partial sums (count <= 256, a power of two) and writes the affine result
__global__ void __launch_bounds__(256) k_direct_fold(const g1_xyzz_t* __restrict__ partial, unsigned count, g1_affine_t* __restrict__ out) {
    extern __shared__ uint4 direct_fold_sm[];
    g1_xyzz_t* part = reinterpret_cast<g1_xyzz_t*>(direct_fold_sm);
    const unsigned t = threadIdx.x;
    part[t] = t < count ? xyzz_load(partial + (size_t)blockIdx.x * count + t) : g1_xyzz_t::identity();
    __syncthreads();
    for (unsigned s2 = blockDim.x >> 1; s2 > 0; s2 >>= 1) {
        if (t < s2) part[t] = xyzz_add(part[t], part[t + s2]);
        __syncthreads();
    }
    if (t == 0) {
        g1_affine_t a = xyzz_to_affine(part[0]);
        fe_store(&out[blockIdx.x].x, a.x); fe_store(&out[blockIdx.x].y, a.y);
    }
}

__global__ void k_msm_combine(const g1_xyzz_t* __restrict__ groups, MsmDims D, size_t M, g1_xyzz_t* __restrict__ out) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    g1_xyzz_t acc = g1_xyzz_t::identity();
    for (unsigned g = D.G; g-- > 0;) {
        for (unsigned i = 0; i < D.c; ++i) acc = xyzz_dbl(acc);
        acc = xyzz_add(acc, xyzz_load(groups + m * D.G + g));
    }
    xyzz_store(out + m, acc);
}

__global__ void k_precompute_table(const g1_affine_t* __restrict__ bases, g1_affine_t* __restrict__ table, MsmDims D) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.n) return;
    g1_affine_t p;
    p.x = fe_load(&bases[i].x); p.y = fe_load(&bases[i].y);
    fe_store(&table[i].x, p.x); fe_store(&table[i].y, p.y);
    g1_xyzz_t cur = g1_xyzz_t::from_affine(p);
    for (unsigned w = 1; w < D.W; ++w) {
        for (unsigned j = 0; j < D.c; ++j) cur = xyzz_dbl(cur);
        g1_affine_t a = xyzz_to_affine(cur);
        g1_affine_t* o = table + (size_t)w * D.tstride + i;
        fe_store(&o->x, a.x); fe_store(&o->y, a.y);
        cur = g1_xyzz_t::from_affine(a);  // keep ZZ = 1 so later doublings stay cheap and exact
    }
}

// Prefix sums of the bases (Hillis-Steele over global memory: n log n additions, once per SRS).
__global__ void k_prefix_lift(const g1_affine_t* __restrict__ in, g1_xyzz_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p;
    p.x = fe_load(&in[i].x); p.y = fe_load(&in[i].y);
    xyzz_store(out + i, p.is_identity() ? g1_xyzz_t::identity() : g1_xyzz_t::from_affine(p));
}
__global__ void k_prefix_step(const g1_xyzz_t* __restrict__ in, g1_xyzz_t* __restrict__ out, size_t n, size_t d) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_xyzz_t a = xyzz_load(in + i);
    if (i >= d) a = xyzz_add(xyzz_load(in + i - d), a);
    xyzz_store(out + i, a);
}

__global__ void k_normalize(const g1_xyzz_t* __restrict__ in, g1_affine_t* __restrict__ out, size_t m) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    g1_affine_t a = xyzz_to_affine(xyzz_load(in + i));
    fe_store(&out[i].x, a.x); fe_store(&out[i].y, a.y);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
MsmPlan msm_plan(size_t n, bool precomp, unsigned force_c) {
    MsmPlan p;
    p.n = n; p.precomp = precomp;
    unsigned best = 0; double best_cost = 0;
    for (unsigned c = 2; c <= 16; ++c) {
        unsigned W = 254 / c + 1;
        double nb = (double)(1u << (c - 1));
        double cost = precomp ? (double)n * W * 10 + nb * 28 : W * ((double)n * 10 + nb * 28);
        if (!best || cost < best_cost) { best = c; best_cost = cost; }
    }
    p.c = force_c ? force_c : best;
    ZK_REQUIRE(p.c >= 2 && p.c <= 16, "msm: window must be in [2,16]");
    p.W = 254 / p.c + 1;
    p.nb = 1u << (p.c - 1);
    p.G = precomp ? 1 : p.W;
    ZK_REQUIRE(n * p.W < (1ull << 31), "msm: n*W must fit 31 bits");
    return p;
}

void MsmWorkspace::ensure(const MsmPlan& p, size_t M) {
    size_t K = p.K();
    counts.ensure(M * K);
    offsets.ensure(M * (K + 1));
    entries.ensure(M * p.entries_per_msm());
    buckets.ensure(M * K);
    groups.ensure(M * p.G);
    order.ensure(M * (K + 1));   // M * K bucket ranks; the latency path keeps its (K + 1)-strided scatter cursors here
    heavy_count.ensure(1);
    heavy_list.ensure(M * p.entries_per_msm() / ZK_HEAVY_MIN + 1);
}

static MsmDims dims_of(const MsmPlan& p) {
    MsmDims D; D.c = p.c; D.W = p.W; D.G = p.G; D.nb = p.nb; D.precomp = p.precomp ? 1 : 0; D.n = (unsigned)p.n; D.tstride = (unsigned)(p.tstride ? p.tstride : p.n);
    D.inner = p.inner ? p.inner : ~(size_t)0; D.outer_stride = p.outer_stride;
    D.diff_off = (unsigned)p.diff_offset;
    // a run is "heavy" when it is several times the mean run length (and at least ZK_HEAVY_MIN)
    size_t mean = p.entries_per_msm() / (p.K() ? p.K() : 1);
    size_t heavy = 4 * mean < ZK_HEAVY_MIN ? ZK_HEAVY_MIN : 4 * mean;
    D.heavy = (unsigned)(heavy > ZK_HEAVY_MAX ? ZK_HEAVY_MAX : heavy);
    return D;
}

void msm_run(const MsmPlan& plan, const fr_t* d_scalars, const g1_affine_t* d_bases, size_t M, g1_xyzz_t* d_out,
             MsmWorkspace& ws, cudaStream_t st) {
    if (M == 0) return;
    ZK_REQUIRE(plan.n > 0 && plan.n < (1ull << 31), "msm: bad n");
    ws.ensure(plan, M);
    MsmDims D = dims_of(plan);
    const size_t K = plan.K();
    const size_t total = M * plan.n;
    ZK_REQUIRE(K <= (1u << 30), "msm: too many buckets");
    const size_t sort_smem = (K + 1024 + D.heavy + 2) * sizeof(uint32_t);
    if (K <= 8192 && plan.n <= (1u << 20) && M >= 32) {   // one CTA per MSM: needs enough MSMs to fill the GPU
        KtScope kt(KT_MSM_SORT, st);
        // measured (profiles/r02_ab_bench.md): 1024 threads 75 ms per 1024 proofs, 512 threads 116 ms, 256 threads 168 ms
        static const unsigned sort_t = [] { const char* e = getenv("ZKGPU_SORT_T"); unsigned v = e ? (unsigned)atoi(e) : 0; return (v == 256 || v == 512 || v == 1024) ? v : 1024u; }();
        switch (D.c) {
#define ZK_SORT_CASE(CT)                                                                                                                          \
    case CT: {                                                                                                                                    \
        static DeviceOnce once;                                                                                                                   \
        once.run([] { ZK_CUDA(cudaFuncSetAttribute(k_msm_sort_smem<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); });            \
        ZK_LAUNCH(k_msm_sort_smem<CT>, (unsigned)M, sort_t, sort_smem, st, d_scalars, D, ws.offsets.p, ws.entries.p, ws.order.p);                 \
    } break;
            ZK_SORT_CASE(11) ZK_SORT_CASE(12) ZK_SORT_CASE(13) ZK_SORT_CASE(14)
            default: ZK_SORT_CASE(0)
#undef ZK_SORT_CASE
        }
    } else {
        KtScope kt(KT_MSM_SORT, st);
        // the histogram is zeroed on the launching stream every call (k_msm_scatter counts it back down to zero, but a fresh
        // allocation or a call that failed between the two kernels must not leak into this one)
        ZK_CUDA(cudaMemsetAsync(ws.counts.p, 0, M * K * sizeof(uint32_t), st));
        ZK_LAUNCH(k_msm_count, ceil_div(total, 256), 256, 0, st, d_scalars, total, D, ws.counts.p);
        unsigned scan_threads = K >= 1024 ? 1024 : 32;
        while (scan_threads < K && scan_threads < 1024) scan_threads <<= 1;
        ZK_LAUNCH(k_msm_scan, (unsigned)M, scan_threads, 0, st, ws.counts.p, ws.offsets.p, (unsigned)K);
        ZK_LAUNCH(k_msm_scatter, ceil_div(total, 256), 256, 0, st, d_scalars, total, D, ws.counts.p, ws.offsets.p, ws.entries.p);
        ZK_LAUNCH(k_msm_order, (unsigned)M, K >= 1024 ? 1024 : 256, (D.heavy + 2) * sizeof(uint32_t), st, ws.offsets.p, ws.order.p, (unsigned)K, D.heavy);
    }
    {
        KtScope kt(KT_MSM_BUCKETS, st);
        ZK_CUDA(cudaMemsetAsync(ws.heavy_count.p, 0, sizeof(uint32_t), st));
        static const size_t split_below = [] { const char* e = getenv("ZKGPU_SPLIT_BELOW"); long v = e ? atol(e) : -1; return v >= 0 ? (size_t)v : (size_t)148 * 512; }();   // measured: 24 MSMs x 4096 buckets already run faster one thread per bucket
        if (M * K < split_below)   // latency regime: too few buckets to fill the GPU with one thread each
            ZK_LAUNCH(k_msm_buckets_split, ceil_div(M * K * ZK_SPLIT, 128), 128, 0, st, d_bases, ws.offsets.p, ws.entries.p, D, M, ws.buckets.p,
                      ws.heavy_count.p, ws.heavy_list.p);
        else
            ZK_LAUNCH(k_msm_buckets, ceil_div(M * K, 128), 128, 0, st, d_bases, ws.offsets.p, ws.entries.p, D, M, ws.buckets.p,
                      ws.order.p, ws.heavy_count.p, ws.heavy_list.p);
        ZK_LAUNCH(k_msm_heavy, 148 * 8, 128, 0, st, d_bases, ws.offsets.p, ws.entries.p, D, ws.buckets.p, ws.heavy_count.p, ws.heavy_list.p);
    }
    KtScope kt(KT_MSM_REDUCE, st);
    if (M * plan.G <= 64 && plan.nb >= ZK_REDUCE_R * 128) {
        // latency regime: R CTAs per bucket group, then fold
        g1_xyzz_t* groups = plan.precomp ? d_out : ws.groups.p;
        ws.partial.ensure(M * plan.G * ZK_REDUCE_R);
        ZK_LAUNCH(k_msm_reduce_multi, (unsigned)(M * plan.G * ZK_REDUCE_R), 128, 0, st, ws.buckets.p, D, ws.partial.p);
        ZK_LAUNCH(k_msm_reduce_fold, ceil_div(M * plan.G, 32), 32, 0, st, ws.partial.p, groups, (unsigned)(M * plan.G));
        if (!plan.precomp) ZK_LAUNCH(k_msm_combine, ceil_div(M, 64), 64, 0, st, ws.groups.p, D, M, d_out);
        return;
    }
    static const unsigned reduce_t = [] { const char* e = getenv("ZKGPU_REDUCE_T"); unsigned v = e ? (unsigned)atoi(e) : 0; return (v == 32 || v == 64 || v == 128 || v == 256) ? v : (unsigned)ZK_REDUCE_T; }();
    unsigned Tmax = M * plan.G < 148 * 4 ? ZK_REDUCE_T_LAT : reduce_t;
    unsigned T = plan.nb < Tmax ? plan.nb : Tmax;
    g1_xyzz_t* groups = plan.precomp ? d_out : ws.groups.p;
    static DeviceOnce reduce_once;
    reduce_once.run([] { ZK_CUDA(cudaFuncSetAttribute(k_msm_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, ZK_REDUCE_T_LAT * (int)sizeof(g1_xyzz_t))); });
    ZK_LAUNCH(k_msm_reduce, (unsigned)(M * plan.G), T, T * sizeof(g1_xyzz_t), st, ws.buckets.p, D, groups);
    if (!plan.precomp) {
        ZK_LAUNCH(k_msm_combine, ceil_div(M, 64), 64, 0, st, ws.groups.p, D, M, d_out);
    }
}

unsigned msm_lat_window() {
    static const unsigned c = [] { const char* e = getenv("ZKGPU_LAT_C"); int v = e ? atoi(e) : 9; return (unsigned)((v >= 6 && v <= 10) ? v : v == 0 ? 0 : 9); }();
    return c;
}

void msm_lat_run(const MsmPlan& plan, const fr_t* const* d_scalars, uint32_t basis_mask, size_t table_stride, const g1_affine_t* d_tables,
                 size_t M, g1_affine_t* d_out_affine, MsmWorkspace& ws, cudaStream_t st) {
    if (M == 0) return;
    ZK_REQUIRE(M <= ZK_LAT_MAX_M && plan.precomp && plan.nb <= 512 && plan.nb >= 32, "msm_lat_run: unsupported shape");
    ws.ensure(plan, M);
    MsmDims D = dims_of(plan);
    LatArgs A;
    for (size_t m = 0; m < ZK_LAT_MAX_M; ++m) A.sc[m] = d_scalars[m < M ? m : 0];
    A.basis_mask = basis_mask; A.table_stride = (uint32_t)table_stride;
    ZK_REQUIRE(2 * table_stride < (1ull << 31), "msm_lat_run: table too large for 31-bit entries");
    const size_t K = plan.nb;
    const dim3 grid(ceil_div(plan.n, 256), (unsigned)M);
    {
        KtScope kt(KT_MSM_SORT, st);
        ZK_CUDA(cudaMemsetAsync(ws.counts.p, 0, M * K * sizeof(uint32_t), st));
        ZK_LAUNCH(k_lat_count, grid, 256, K * sizeof(uint32_t), st, A, D, ws.counts.p);
        unsigned scan_threads = 32;
        while (scan_threads < K && scan_threads < 1024) scan_threads <<= 1;
        ZK_LAUNCH(k_msm_scan, (unsigned)M, scan_threads, 0, st, ws.counts.p, ws.offsets.p, (unsigned)K);
        // the (K+1)-strided cursor array lives in `order` (unused by this path)
        ZK_LAUNCH(k_lat_init_cursors, ceil_div(M * (K + 1), 256), 256, 0, st, ws.offsets.p, ws.order.p, M * (K + 1));
        ZK_LAUNCH(k_lat_scatter, grid, 256, 2 * K * sizeof(uint32_t), st, A, D, ws.order.p, ws.entries.p);
    }
    {
        KtScope kt(KT_MSM_BUCKETS, st);
        ZK_CUDA(cudaMemsetAsync(ws.heavy_count.p, 0, sizeof(uint32_t), st));
        // Threads per bucket (TPB).  A CTA of 128 threads serves 128 / TPB buckets; a thread issues mean / TPB mixed additions plus
        // log2(TPB) fold additions (paid by the whole warp however few lanes are still active), and 592 CTAs are resident at a time.
        // Measured on the withdraw proof (profiles/r02_single_proof.md; c = 10, 416 points per bucket on average):
        //   24 MSMs: TPB 16 1311 us, 32 1487 us, 8 1490-1634 us   |   8 MSMs: TPB 32 372 us, 16 482 us
        //    5 MSMs: TPB 32 405 us, 16 514 us                     |   1 MSM : TPB 128 128 us
        const size_t nbk = M * K;
        unsigned tpb = nbk >= 8192 ? 16 : nbk >= 1536 ? 32 : 128;
        static const unsigned tpb_force = [] { const char* e = getenv("ZKGPU_LAT_TPB"); int v = e ? atoi(e) : 0; return (unsigned)((v == 16 || v == 32 || v == 64 || v == 128) ? v : 0); }();
        if (tpb_force) tpb = tpb_force;
#define ZK_LAT_BUCKETS(TPB) ZK_LAUNCH(k_lat_buckets<TPB>, ceil_div(nbk * TPB, 128), 128, 0, st, d_tables, ws.offsets.p, ws.entries.p, D, M, ws.buckets.p, \
                                      ws.heavy_count.p, ws.heavy_list.p)
        if (tpb == 16) ZK_LAT_BUCKETS(16);
        else if (tpb == 32) ZK_LAT_BUCKETS(32);
        else if (tpb == 64) ZK_LAT_BUCKETS(64);
        else ZK_LAT_BUCKETS(128);
#undef ZK_LAT_BUCKETS
        static DeviceOnce once;
        once.run([] { ZK_CUDA(cudaFuncSetAttribute(k_lat_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * (int)sizeof(g1_xyzz_t))); });
        // partial sums + completion counters for as many heavy buckets as the worklist can hold
        const size_t max_heavy = M * plan.entries_per_msm() / ZK_HEAVY_MIN + 1;
        ws.heavy_partial.ensure(max_heavy * ZK_LAT_HEAVY_SPLIT);
        ws.heavy_done.ensure(max_heavy);
        ZK_CUDA(cudaMemsetAsync(ws.heavy_done.p, 0, max_heavy * sizeof(uint32_t), st));
        ZK_LAUNCH(k_lat_heavy, 128 * ZK_LAT_HEAVY_SPLIT / 4, 512, 512 * sizeof(g1_xyzz_t), st, d_tables, ws.offsets.p, ws.entries.p, D, ws.buckets.p,
                  ws.heavy_count.p, ws.heavy_list.p, ws.heavy_partial.p, ws.heavy_done.p);
    }
    KtScope kt(KT_MSM_REDUCE, st);
    ws.partial.ensure(M * (K / 32) * 2);
    ZK_LAUNCH(k_lat_reduce_a, dim3((unsigned)(K / 32), (unsigned)M), 32, 0, st, ws.buckets.p, D, ws.partial.p);
    ZK_LAUNCH(k_lat_reduce_b, (unsigned)M, 32, 0, st, ws.partial.p, D, d_out_affine);
}

// ---- direct path, host side ----
size_t msm_direct_points_per_basis(size_t n) { const unsigned c = msm_lat_window(); return (size_t)direct_W(c) * n * direct_D(c); }
bool msm_direct_enabled() {
    static const bool on = [] { const char* e = getenv("ZKGPU_DIRECT"); return e ? atoi(e) != 0 : true; }();
    return on && (msm_lat_window() == 8 || msm_lat_window() == 9);
}
// window_table: T[w][i] = 2^(8 w) * base[i] (the latency plan's table for c = 8, W = 32); direct[(w * n + i) * D + d - 1] = d * T[w][i]
void msm_direct_build(const g1_affine_t* d_window_table, size_t n, g1_affine_t* d_direct, cudaStream_t st) {
    const unsigned c = msm_lat_window(), W = direct_W(c), D = direct_D(c);
    DevBuf<g1_xyzz_t> tmp(n * D);
    DevBuf<fq_t> pre(n * D);
    for (unsigned w = 0; w < W; ++w) {
        ZK_LAUNCH(k_direct_multiples, ceil_div(n, 64), 64, 0, st, d_window_table + (size_t)w * n, tmp.p, (unsigned)n, D);
        ZK_LAUNCH(k_direct_normalize, ceil_div(n, 64), 64, 0, st, tmp.p, pre.p, d_direct + (size_t)w * n * D, (unsigned)n, D);
    }
    ZK_CUDA(cudaStreamSynchronize(st));
}
void msm_direct_run(const fr_t* const* d_scalars, uint32_t basis_mask, size_t table_stride, const g1_affine_t* d_direct, size_t n, size_t tstride,
                    size_t M, g1_affine_t* d_out_affine, MsmWorkspace& ws, cudaStream_t st) {
    if (M == 0) return;
    ZK_REQUIRE(M <= ZK_LAT_MAX_M && n >= 1 && n <= tstride, "msm_direct_run: unsupported shape");
    DirectArgs A;
    for (size_t m = 0; m < ZK_LAT_MAX_M; ++m) A.sc[m] = d_scalars[m < M ? m : 0];
    A.basis_mask = basis_mask; A.table_stride = table_stride; A.n = (unsigned)n; A.tstride = (unsigned)tstride;
    // CTAs per MSM: all CTAs of the launch resident at once (148 SMs x 4 CTAs of 128 threads at 128 registers), at most one unit each
    const unsigned c = msm_lat_window();
    const unsigned units = ceil_div(n, 128) * ((direct_W(c) + ZK_DIRECT_WG - 1) / ZK_DIRECT_WG);
    unsigned ctas = (unsigned)(((size_t)148 * 4) / M);
    if (ctas > units) ctas = units;
    if (ctas > 256) ctas = 256;
    if (ctas < 1) ctas = 1;
    unsigned folded = 1;
    while (folded < ctas) folded <<= 1;       // partial sums per MSM, padded to a power of two for the fold tree
    ws.partial.ensure(M * ctas);
    {
        KtScope kt(KT_MSM_BUCKETS, st);
        if (c == 8) ZK_LAUNCH(k_direct_sum<8>, dim3(ctas, (unsigned)M), 128, 0, st, A, d_direct, units, ws.partial.p);
        else ZK_LAUNCH(k_direct_sum<9>, dim3(ctas, (unsigned)M), 128, 0, st, A, d_direct, units, ws.partial.p);
    }
    KtScope kt(KT_MSM_REDUCE, st);
    ZK_LAUNCH(k_direct_fold, (unsigned)M, folded, folded * sizeof(g1_xyzz_t), st, ws.partial.p, ctas, d_out_affine);
}

unsigned long long msm_entries_counter(bool reset) {
    unsigned long long v = 0;
    ZK_CUDA(cudaMemcpyFromSymbol(&v, g_msm_entries, sizeof v));
    if (reset) { const unsigned long long z = 0; ZK_CUDA(cudaMemcpyToSymbol(g_msm_entries, &z, sizeof z)); }
    return v;
}

bool msm_diff_enabled() {
    static const bool on = [] { const char* e = getenv("ZKGPU_MSM_DIFF"); return e ? atoi(e) != 0 : true; }();
    return on;
}

void msm_prefix_bases(const g1_affine_t* d_bases, size_t n, g1_affine_t* d_out, cudaStream_t st) {
    if (n == 0) return;
    DevBuf<g1_xyzz_t> a, b;
    a.alloc(n); b.alloc(n);
    const unsigned grid = (unsigned)ceil_div(n, 128);
    ZK_LAUNCH(k_prefix_lift, grid, 128, 0, st, d_bases, a.p, n);
    g1_xyzz_t *cur = a.p, *nxt = b.p;
    for (size_t d = 1; d < n; d <<= 1) {
        ZK_LAUNCH(k_prefix_step, grid, 128, 0, st, cur, nxt, n, d);
        g1_xyzz_t* t = cur; cur = nxt; nxt = t;
    }
    ZK_LAUNCH(k_normalize, grid, 128, 0, st, cur, d_out, n);
    ZK_CUDA(cudaStreamSynchronize(st));   // the scratch buffers go away with this frame
}

void msm_precompute_table(const MsmPlan& plan, const g1_affine_t* d_bases, g1_affine_t* d_table, cudaStream_t st) {
    MsmDims D = dims_of(plan);
    ZK_LAUNCH(k_precompute_table, ceil_div(plan.n, 64), 64, 0, st, d_bases, d_table, D);
}

void g1_normalize(const g1_xyzz_t* d_in, g1_affine_t* d_out, size_t m, cudaStream_t st) {
    if (!m) return;
    KtScope kt(KT_MISC, st);
    ZK_LAUNCH(k_normalize, ceil_div(m, 64), 64, 0, st, d_in, d_out, m);
}

}  // namespace zk
