// Internal interface of the NTT module (ntt.cu).
#pragma once
#include "common.cuh"
#include "fp.cuh"
#include <cstring>

namespace zk {

static const unsigned NTT_SINGLE_PASS_MAX_LOG = 12;  // 2^12 * 32 B = 128 KB of shared memory
static const unsigned NTT_CLUSTER_LOG = 13;          // 2^13: one launch of two-CTA clusters (distributed shared memory)
static const size_t NTT_CLUSTER_MIN_BATCH = 592;     // ... from 592 polynomials on (148 SMs x 8 CTAs of a pair: several full waves)

struct NttJob {
    const fr_t* in = nullptr;   // batch polynomials, stride in_stride (0 = N)
    fr_t* out = nullptr;        // may equal `in`
    fr_t* scratch = nullptr;    // batch * N elements, needed when log_n > NTT_SINGLE_PASS_MAX_LOG
    size_t batch = 1;
    unsigned log_n = 0;
    size_t in_stride = 0, out_stride = 0;
    // optional two-level batch addressing: polynomial p lives at (p / inner) * outer_stride + (p % inner) * stride
    size_t in_inner = 0, in_outer_stride = 0, out_inner = 0, out_outer_stride = 0;
    size_t in_valid = 0;        // 0 = N; inputs at index >= in_valid read as zero (zero-padding)
    fr_t omega;
    int pre_coset = 0, post_coset = 0;  // multiply element i by cs1 / cs2 when i%3 == 1 / 2
    fr_t cs1, cs2;
    int has_scale = 0;
    fr_t scale;
    // optional per-element pre-multiplier: polynomial p multiplies input element i by pre_table[(p % pre_count) * N + i]
    // (coset evaluation on g_c * H: table row c holds g_c^i); with in_broadcast the `in_inner` polynomials of one group
    // all read the same input polynomial (stride 0 inside the group)
    const fr_t* pre_table = nullptr;
    unsigned pre_count = 1;
    int in_broadcast = 0;
};

void ntt_run(const NttJob& job, cudaStream_t st);
size_t ntt_scratch_elems(unsigned log_n, size_t batch);
const fr_t* ntt_twiddles(unsigned log_n, const fr_t& omega, cudaStream_t st);
// d_out[i] = base^i for i < 2^log_count (device buffer)
void fr_power_table(fr_t* d_out, unsigned log_count, const fr_t& base, cudaStream_t st);
void ntt_clear_cache();

// host-side Fr helpers (run the same generated limb code on the CPU)
fr_t fr_from_limbs(const uint32_t* l);
fr_t fr_from_u64(uint64_t v);
fr_t fr_pow_u64(fr_t b, uint64_t e);
fr_t fr_omega(unsigned log_n);
fr_t fr_omega_inv(unsigned log_n);
fr_t fr_pow2_inv(unsigned log_n);

}  // namespace zk
