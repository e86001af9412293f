// Internal interface of the MSM module (msm.cu).
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace zk {

// Pippenger layout.  Signed c-bit digits, W = 254/c + 1 windows, nb = 2^(c-1) buckets per group.
//   plain   : G = W bucket groups (one per window), entries index the caller's bases directly
//   precomp : G = 1 — every window shares one bucket set because the bases carry a table
//             T[w][i] = 2^(c*w) * base[i] (fixed-base SRS path); entries index w*n + i
struct MsmPlan {
    unsigned c = 0, W = 0, G = 0, nb = 0;
    bool precomp = false;
    size_t n = 0;        // scalars per MSM
    size_t tstride = 0;  // precomp: points per window in the table (0 = n)
    // precomp: the table continues, diff_offset points further on, with the same windows over the PREFIX SUMS P_i = G_0 + ... + G_i
    // of the bases (0 = no such half).  sum s_i G_i = sum (s_i - s_{i+1}) P_i, so a column that is constant over long stretches
    // (a grand product outside the rows that carry copies, a sorted lookup column, padding) costs its number of CHANGES.
    size_t diff_offset = 0;
    // optional two-level scalar addressing: MSM m reads scalars at (m / inner) * outer_stride + (m % inner) * n
    size_t inner = 0, outer_stride = 0;
    size_t K() const { return (size_t)G * nb; }            // buckets per MSM
    size_t entries_per_msm() const { return n * W; }
};
MsmPlan msm_plan(size_t n, bool precomp, unsigned force_c = 0);

struct MsmWorkspace {
    DevBuf<uint32_t> counts;    // M * K      (zeroed on the stream by every call that uses it)
    DevBuf<uint32_t> offsets;   // M * (K+1)
    DevBuf<uint32_t> entries;   // M * n * W
    DevBuf<g1_xyzz_t> buckets;  // M * K
    DevBuf<g1_xyzz_t> groups;   // M * G
    DevBuf<g1_xyzz_t> partial;  // latency regime: R partial sums per bucket group
    DevBuf<uint32_t> order;     // M * K bucket keys by decreasing run length
    DevBuf<uint32_t> heavy_count;  // worklist of buckets with long runs (k_msm_heavy)
    DevBuf<uint64_t> heavy_list;
    DevBuf<g1_xyzz_t> heavy_partial;   // latency path: partial sums of the CTAs sharing a heavy bucket
    DevBuf<uint32_t> heavy_done;       //               and their completion counters
    void ensure(const MsmPlan& p, size_t M);
};

// out[m] = sum_i scalars[m*n + i] * bases[i]  (XYZZ, not normalised).  Scalars in Montgomery form.
void msm_run(const MsmPlan& plan, const fr_t* d_scalars, const g1_affine_t* d_bases_or_table, size_t M,
             g1_xyzz_t* d_out, MsmWorkspace& ws, cudaStream_t st);
// Latency path for M <= ZK_LAT_MAX_M fixed-base MSMs over a narrow-window table (`plan` from msm_plan(n, true, c_lat)): MSM m reads
// n scalars at d_scalars[m] (host array of device pointers) and points from d_tables + (bit m of basis_mask) * table_stride.
// Affine results (normalised) in d_out_affine[m].
#define ZK_LAT_MAX_M 32
unsigned msm_lat_window();   // window width of the latency tables (ZKGPU_LAT_C, default 9; 0 disables the latency path)
void msm_lat_run(const MsmPlan& plan, const fr_t* const* d_scalars, uint32_t basis_mask, size_t table_stride, const g1_affine_t* d_tables,
                 size_t M, g1_affine_t* d_out_affine, MsmWorkspace& ws, cudaStream_t st);
// Direct path (no buckets): a table of EVERY multiple d * 2^(8 w) * G_i, d = 1 .. 128, built from the c = 8 window table at SRS
// registration (msm_direct_points_per_basis(n) points per basis); an MSM is the plain sum of n * 32 table entries.
bool msm_direct_enabled();     // ZKGPU_DIRECT (default on) and the latency tables use c = 8 or 9
size_t msm_direct_points_per_basis(size_t n);
void msm_direct_build(const g1_affine_t* d_window_table, size_t n, g1_affine_t* d_direct, cudaStream_t st);
void msm_direct_run(const fr_t* const* d_scalars, uint32_t basis_mask, size_t table_stride, const g1_affine_t* d_direct, size_t n, size_t tstride,
                    size_t M, g1_affine_t* d_out_affine, MsmWorkspace& ws, cudaStream_t st);
// table[w*n + i] = 2^(c*w) * bases[i], affine
void msm_precompute_table(const MsmPlan& plan, const g1_affine_t* d_bases, g1_affine_t* d_table, cudaStream_t st);
// d_out[i] = d_bases[0] + ... + d_bases[i] (affine), i < n; one-off work at SRS registration
void msm_prefix_bases(const g1_affine_t* d_bases, size_t n, g1_affine_t* d_out, cudaStream_t st);
unsigned long long msm_entries_counter(bool reset);   // current device; see g_msm_entries in msm.cu
bool msm_diff_enabled();   // ZKGPU_MSM_DIFF (default on)
// affine normalisation of m points (one inversion each)
void g1_normalize(const g1_xyzz_t* d_in, g1_affine_t* d_out, size_t m, cudaStream_t st);

}  // namespace zk
