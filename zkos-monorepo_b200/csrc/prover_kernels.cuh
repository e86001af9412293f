// Device-side building blocks of the batched create_proof pipeline (prover_kernels.cu).
// Every launcher works on a sub-batch of B independent proofs of the same circuit.
#pragma once
#include "common.cuh"
#include "fp.cuh"

namespace zk {

struct ColSrc { uint32_t type, index; };  // COL_ADVICE / COL_FIXED / COL_INSTANCE

// Per-proof Fiat-Shamir challenges (Montgomery), device resident: [B] of this struct
struct Challenges { fr_t theta, beta, gamma, y, x; };

// dst[b*proof_stride + col*col_stride + row_start + j] = from_u512(raw[(b*cols + col)*rows + j])
void launch_scatter_random(fr_t* dst, size_t proof_stride, size_t col_stride, size_t row_start, const uint64_t* raw_wide,
                           size_t B, size_t cols, size_t rows, cudaStream_t st);

struct PermArgs {
    const fr_t* adv;  size_t adv_proof_stride;   // [B][A][n] Lagrange values
    const fr_t* inst; size_t inst_proof_stride;  // [B][n]
    const fr_t* fixed_vals;                      // [F][n]
    const fr_t* sigma_vals;                      // [S][n]
    const ColSrc* cols;                          // [S] permutation columns (device)
    const fr_t* delta_pows;                      // [S] delta^c
    const fr_t* omega_tw;                        // omega^i, i < n/2
    const Challenges* ch;                        // [B]
    unsigned k, S, chunk, P;
};
// num/den: [B][P][n]
void launch_perm_num_den(const PermArgs& a, fr_t* num, fr_t* den, size_t B, cudaStream_t st);
// in-place batch inversion of `count` elements (zeros stay zero)
void launch_batch_inverse(fr_t* a, size_t count, cudaStream_t st);
// z_local[b][s][row] = prod_{i<row} num*den_inv  (exclusive prefix product), one CTA per (b,s)
void launch_perm_scan(const fr_t* num, const fr_t* den_inv, fr_t* z, unsigned k, size_t BP, cudaStream_t st);
// chain the sets (z_s *= prod_{t<s} z_t[u]) and overwrite the last bf rows with blinding values
void launch_perm_finalize(fr_t* z, fr_t* carries /*[B][P] scratch*/, unsigned k, unsigned P, unsigned bf, const uint64_t* raw_wide /*[B][P][bf] x 8 u64*/, size_t B, cudaStream_t st);

// random polynomial of the vanishing argument, filled in chunks of `chunk` coefficients as halo2's vanishing prover does with
// rayon: coefficient i of proof b = Fr::random of ChaCha20(seed[b][i / chunk]) block i % chunk
void launch_chacha_poly(const uint8_t* seeds /*[B][nchunks][32]*/, fr_t* out /*[B][n]*/, size_t n, size_t B, size_t chunk, size_t nchunks, cudaStream_t st);

// ---- lookup arguments (halo2 lookup/prover.rs) ------------------------------------------------------------
// Expression programs of all lookups, flattened: lookup l owns expressions [lk_off[l], lk_off[l+1]) — first half
// inputs, second half tables — and expression e owns instructions [expr_off[e], expr_off[e+1]) of `prog`.
struct LookupProgs {
    const uint32_t* prog; const uint32_t* expr_off; const uint32_t* lk_off;
    const fr_t* constants; const int32_t* adv_q; const int32_t* fix_q; const int32_t* inst_q;
    unsigned L;
};
struct LookupCompressArgs {
    const fr_t* adv; size_t adv_proof_stride;    // [B][A][n] Lagrange values (blinded)
    const fr_t* inst; size_t inst_proof_stride;  // [B][n]
    const fr_t* fixed_vals;                      // [F][n]
    const Challenges* ch;
    LookupProgs lp;
    unsigned k;
};
// compressed input / table values: out_in, out_tab [B][L][n]
void launch_lookup_compress(const LookupCompressArgs& a, fr_t* out_in, fr_t* out_tab, size_t B, cudaStream_t st);
// permute_expression_pair for B*L (input, table) pairs: sorts canonical values (rows < usable), assigns the table
// column; perm_in / perm_tab [B*L][n] receive rows < usable (blinding rows are written separately).
// sort_a / sort_t: scratch [B*L][n].  d_error[b] (one flag per proof, b = pair / L) is set to 1 if an input value of
// proof b is missing from its table.
void launch_lookup_permute(const fr_t* comp_in, const fr_t* comp_tab, fr_t* perm_in, fr_t* perm_tab, fr_t* sort_a, fr_t* sort_t,
                           unsigned k, size_t usable, size_t BL, unsigned L, int* d_error, cudaStream_t st);
// num = (in + beta)(tab + gamma), den = (perm_in + beta)(perm_tab + gamma); all [B][L][n]
void launch_lookup_num_den(const fr_t* comp_in, const fr_t* comp_tab, const fr_t* perm_in, const fr_t* perm_tab, const Challenges* ch,
                           fr_t* num, fr_t* den, unsigned k, unsigned L, size_t B, cudaStream_t st);

struct EvalHArgs {
    // per-proof columns on the quotient cosets
    const fr_t* adv_ext; size_t adv_ext_proof_stride;  // [B][A+1][Qc*n], instance coset at column A
    const fr_t* z_ext;   size_t z_ext_proof_stride;    // [B][P][Qc*n]
    // proving key
    const fr_t* fixed_ext;   // [F][Qc*n]
    const fr_t* sigma_ext;   // [S][Qc*n]
    const fr_t* l0; const fr_t* l_last; const fr_t* l_active;  // [Qc*n]
    const fr_t* t_inv;       // [2^(ek-k)]
    const fr_t* ext_tw;      // ext_omega^i, i < en/2
    const fr_t* delta_pows;  // unused by the kernel (delta is a constant) — kept for symmetry
    const ColSrc* cols;      // [S]
    const Challenges* ch;    // [B]
    // gate programs
    const uint32_t* prog;        // concatenated (op,arg) pairs
    const uint32_t* gate_off;    // [num_gates+1] offsets into prog (in instructions)
    const fr_t* constants;
    const int32_t* adv_q;  // [num_advice_queries][2] = (column, rotation)
    const int32_t* fix_q;
    const int32_t* inst_q;
    // lookups: quotient cosets [B][L][3][Qc*n] in the order (z, permuted input, permuted table)
    const fr_t* lk_ext; size_t lk_ext_proof_stride;
    LookupProgs lp;
    unsigned num_gates, A, S, chunk, P, k, ek;
    // table sizes (entries): the gate programs and the query tables are staged in shared memory by every CTA, so that an operand
    // costs one global load (the column value) instead of a chain of three (instruction -> query -> value)
    unsigned n_prog, n_adv_q, n_fix_q, n_inst_q;
    unsigned Qc;  // cosets of the size-n subgroup the quotient is evaluated on (= number of quotient pieces)
    int rotation_last;
    fr_t zeta;  // coset generator (Montgomery)
};
void launch_eval_h(const EvalHArgs& a, fr_t* h /*[B][Qc*n]*/, size_t B, cudaStream_t st);

// Quotient pieces from its values on Qc cosets.  hc [B][Qc][n] holds, for coset c, the coefficients of h(g_c X) mod (X^n - 1)
// (size-n iNTT of the coset values); overwritten in place with the pieces h_j, h(X) = sum_j X^(jn) h_j(X):
//   h_j[m] = sum_c vinv[j*Qc + c] * ginv[c][m] * hc[c][m],   ginv[c][m] = g_c^-m,  vinv = inverse of V[c][j] = (g_c^n)^j
void launch_coset_combine(fr_t* hc, const fr_t* ginv /*[Qc][n]*/, const fr_t* vinv /*[Qc*Qc]*/, unsigned Qc, unsigned k, size_t B, cudaStream_t st);

struct EvalJob { const fr_t* poly; fr_t x; };
// out[j] = poly_j(x_j) for polynomials of n = 2^k coefficients
void launch_poly_eval(const EvalJob* jobs, fr_t* out, size_t num_jobs, unsigned k, cudaStream_t st);

struct LinTerm { const fr_t* poly; fr_t coef; };
// out_j[i] = sum_{t in [off_j, off_{j+1})} coef_t * poly_t[i], i < n
void launch_lincomb(const LinTerm* terms, const uint32_t* job_off, fr_t* const* outs, size_t num_jobs, size_t n, cudaStream_t st);
// a_j[i] -= low_j[i] for i < cnt (low-degree correction r(X))
void launch_sub_low(fr_t* const* polys, const fr_t* low /*[jobs][4]*/, size_t num_jobs, cudaStream_t st);
// out_j = in_j / (X - pt_j) (remainder dropped), n coefficients in, n out (top coefficient zero)
struct DivJob { const fr_t* in; fr_t* out; const fr_t* low; fr_t pt; };  // low: optional 4 coefficients subtracted from `in` on load
void launch_kate_div(const DivJob* jobs, size_t num_jobs, unsigned k, cudaStream_t st);

// dst (device) <- mapped_src (device-visible pointer of pinned host memory), copied by `ctas` CTAs instead of the copy engine
void launch_pull_from_host(void* dst, const void* mapped_src, size_t bytes, unsigned ctas, cudaStream_t st);

// sigma values for keygen: out[c][row] = delta^{map_col} * omega^{map_row}
void launch_sigma_values(const uint32_t* map_col, const uint32_t* map_row, const fr_t* delta_pows, const fr_t* omega_tw, fr_t* out,
                         size_t S, unsigned k, cudaStream_t st);
// quotient cosets out of halo2's extended domain: dst[c * n + j] = ext[c + (j << log_e)], c < Qc, j < n = 2^k
// (extended point i = zeta * ext_omega^i, so the entries i = c mod 2^log_e are the coset g_c * H in the order omega^j)
void launch_gather_cosets(const fr_t* ext, fr_t* dst, unsigned k, unsigned log_e, unsigned Qc, cudaStream_t st);
// out[i] = 1 - a[i] - b[i]
void launch_one_minus_sum(const fr_t* a, const fr_t* b, fr_t* out, size_t n, cudaStream_t st);
// elementwise helpers (also exported through the C ABI for synthetic witness generation):
// op 0 mul, 1 add, 2 sub, 3 to_mont(a), 4 from_mont(a) (b unused for 3/4)
void launch_vec_op(int op, const fr_t* a, const fr_t* b, fr_t* out, size_t n, cudaStream_t st);
// dst[i] = from_u512(raw[i])
void launch_reduce_wide(const uint64_t* raw_wide, fr_t* out, size_t n, cudaStream_t st);

}  // namespace zk
