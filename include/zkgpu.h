/* libzkgpu — C ABI of the B200-native BN254 proving backend.
 *
 * This is the drop-in boundary for the data-parallel core of Shielder's halo2 prover.  halo2 v0.3.0
 * has no backend trait: `best_multiexp` / `best_fft` are free functions in un-vendored dependencies
 * (halo2curves 0.6.1 msm.rs / fft.rs, halo2_proofs v0.3.0 arithmetic.rs, poly/domain.rs,
 * poly/kzg/commitment.rs — /root/reference/Cargo.lock:2332-2334,2367-2368), reached from the
 * reference only through
 *     shielder_circuits::generate_proof(&params, &pk, circuit, &public_input, rng)
 *       (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111,
 *        /root/reference/crates/shielder-account/src/call_data.rs:489-501,
 *        /root/reference/tee/crates/shielder-prover-tee/src/circuits/mod.rs:70-78)
 * and the ParamsKZG data seam (/root/reference/crates/powers-of-tau/lib.rs:64,71,263,280).
 * Each entry point below names the upstream function whose body a `[patch]`ed halo2curves /
 * halo2_proofs would replace with a call to it (binding shown in INTEGRATION.md).
 *
 * Conventions
 *   - Field elements: 4 x u64, little-endian limbs, Montgomery form (R = 2^256) — the in-memory
 *     layout of Rust `bn256::Fr` / `Fq` and of halo2 `SerdeFormat::RawBytes`
 *     (verified on /root/reference/resources/ppot_0080_11_raw).
 *   - G1 affine: (x, y) as 8 x u64; identity = (0,0) (`G1Affine::identity()`).
 *   - G1 projective results: Jacobian (X, Y, Z) as 12 x u64, returned normalised (Z = 1, or
 *     (0,1,0) for the identity) — projective representatives are not canonical, the group element is.
 *   - All pointers are host memory unless the function name ends in `_dev`.
 *   - All functions are synchronous, thread-safe and re-entrant, return 0 on success and a negative code
 *     otherwise; zkgpu_last_error() gives the thread-local message.  The upstream functions are infallible
 *     and panic on misuse; the Rust shim maps a non-zero return to `panic!`.
 *   - There is no process-wide lock.  The primitive calls (MSM / NTT / G1-FFT) serialise per device on the
 *     scratch buffers they share; proving calls serialise per proving key and device; zkgpu_prove callers
 *     never serialise: concurrent requests are coalesced into lock-step batches.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with ZKGPU_ERR_CUDA.
 */
#ifndef ZKGPU_H
#define ZKGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKGPU_OK 0
#define ZKGPU_ERR_CUDA (-1)
#define ZKGPU_ERR_ARG (-2)
#define ZKGPU_ERR_STATE (-3)
#define ZKGPU_ERR_INTERNAL (-4)
#define ZKGPU_ERR_WITNESS (-5)   /* create_proof failed for THIS witness (plonk::Error::ConstraintSystemFailure) */

/* Selects the CUDA devices this process proves on: bit i of `device_mask` = CUDA device i,
 * 0 = every visible device.  Idempotent for the same mask; once per process, thread-safe.  Without a call the first
 * compute entry point selects device 0.  Key material (SRS tables, proving keys) is replicated on every selected
 * device; the primitive calls run on the lowest selected device, zkgpu_prove_batch / zkgpu_prove on all of them. */
int zkgpu_init(int device_mask);
void zkgpu_shutdown(void);
/* number of selected devices / CUDA index of the i-th one (-1 if out of range) */
int zkgpu_device_count(void);
int zkgpu_device_index(int slot);
const char* zkgpu_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int zkgpu_abi_version(void);

/* Keccak-256 with the EVM's padding: the hash behind the Keccak256 transcript (known answers at
 * /root/reference/crates/shielder-account/src/secrets.rs:75-106) and behind the contract-side `commitment` public inputs
 * (/root/reference/contracts/Shielder.sol:351-356).  Host only, needs no GPU. */
int zkgpu_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);

/* ---- MSM: halo2curves::msm::best_multiexp(coeffs, bases) -> G1 -------------------------------- */
int zkgpu_msm_g1(const uint64_t* scalars, const uint64_t* bases_affine, size_t n, uint64_t out_jacobian[12]);

/* ---- best_multiexp with RESIDENT bases, points split across the selected devices ("partition points into G
 * contiguous shards; each GPU runs a full Pippenger on its shard -> one point; gather G points over NVLink P2P, G - 1 EC adds").
 * zkgpu_bases_register uploads shard g of the n points to device g only.  zkgpu_msm_g1_bases(handle, scalars, n, out, kernel_ms):
 * scalars = n field elements in host memory, or NULL to reuse the scalars the previous call left in HBM (kernel-only timing);
 * kernel_ms (may be NULL) receives the slowest shard's device time.  In the one-process-per-GPU deployment every rank registers its
 * own shard and the 64-byte partial results are combined with zkgpu_g1_sum_affine. */
int zkgpu_bases_register(const uint64_t* bases_affine, size_t n, uint64_t* handle_out);
int zkgpu_bases_release(uint64_t bases);
int zkgpu_msm_g1_bases(uint64_t bases, const uint64_t* scalars, size_t n, uint64_t out_jacobian[12], double* kernel_ms);

/* ---- SRS-resident MSM: ParamsKZG::{commit, commit_lagrange} (poly/kzg/commitment.rs) -----------
 * zkgpu_srs_register uploads g and g_lagrange (n = 2^k points each, as ParamsKZG holds them;
 * crates/powers-of-tau/lib.rs:71 builds it, :280 `get_g`) once and precomputes, on every selected device, the fixed-base window
 * tables of both bases (plain and over the prefix sums of the basis, 2 x 10 MiB per basis at k = 13) and, when they fit a quarter of
 * the free device memory, the tables of all window multiples that serve single commitments (7.8 GB at k = 13; ZKGPU_DIRECT=0 or
 * ZKGPU_LAT_C=0 to do without).  basis: 0 = g (monomial, `commit`), 1 = g_lagrange (`commit_lagrange`). */
int zkgpu_srs_register(const uint64_t* g, const uint64_t* g_lagrange, uint32_t k, uint64_t* handle_out);
int zkgpu_srs_release(uint64_t srs);
int zkgpu_msm_g1_srs(uint64_t srs, int basis, const uint64_t* scalars, size_t n, uint64_t out_jacobian[12]);
/* m independent commitments over the same basis in one launch; scalars is m x n, out is m x 8 (affine). */
int zkgpu_msm_g1_srs_batch(uint64_t srs, int basis, const uint64_t* scalars, size_t n, size_t m, uint64_t* out_affine);

/* ---- NTT: halo2curves::fft::best_fft::<Fr, Fr>(a, omega, log_n), in place, natural order ------- */
int zkgpu_ntt_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n);
int zkgpu_ntt_fr_batch(uint64_t* a /* m x n */, const uint64_t omega[4], uint32_t log_n, size_t m);

/* ---- EvaluationDomain (halo2_proofs poly/domain.rs) ------------------------------------------- */
/* lagrange_to_coeff (inverse=1) / coeff_to_lagrange (inverse=0) on m polynomials of 2^k values */
int zkgpu_domain_ntt_fr(uint64_t* a, uint32_t k, int inverse, size_t m);
/* coeff_to_extended: n = 2^k coefficients -> 2^ext_k evaluations on the zeta-coset */
int zkgpu_coset_ntt_fr(const uint64_t* coeffs, uint32_t k, uint32_t ext_k, uint64_t* out);
/* extended_to_coeff: 2^ext_k coset evaluations (in place) -> coefficients; the first
 * 2^k * quotient_degree entries are the result (the upstream `truncate`), the rest is zeroed. */
int zkgpu_coset_intt_fr(uint64_t* evals, uint32_t k, uint32_t ext_k, uint32_t quotient_degree);

/* ---- G1 FFT: best_fft::<Fr, G1> as used by g_to_lagrange (ParamsKZG::from_parts / downsize) ----
 * points: n x 12 u64 Jacobian, transformed in place and returned normalised (Z = 1). */
int zkgpu_fft_g1(uint64_t* points_jacobian, const uint64_t omega[4], uint32_t log_n);
/* g_to_lagrange(g, k): n^-1 * FFT_{omega^-1}(g), affine in, affine out */
int zkgpu_g_to_lagrange(const uint64_t* g_affine, uint32_t k, uint64_t* out_affine);

/* ParamsKZG::setup(k, SmallRng::seed_from_u64(seed)) — the seeded SRS of the reference's prove/verify tests
 * (/root/reference/crates/halo2-verifier/src/generator.rs:118-119): g[i] = G * s^i, g_lagrange = g_to_lagrange(g). */
int zkgpu_params_setup(uint32_t k, uint64_t seed, uint64_t* g_out, uint64_t* g_lagrange_out);
/* The same from a RUNNING SmallRng: rng_state is the caller's xoshiro256++ state (rand 0.8.5 SmallRng on 64-bit targets), advanced
 * by the one `Fr::random` that setup draws — generator.rs:117-120 hands one rng to setup, witness generation and the prover in turn. */
int zkgpu_params_setup_rng(uint32_t k, uint64_t rng_state[4], uint64_t* g_out, uint64_t* g_lagrange_out);

/* g_lagrange_out may be NULL (then k up to 24: bases for the large-MSM sweep). */
/* g_out[i] = s^(start + i) * G, i < count: a slice of that setup's `g` (s = the first Fr::random of SmallRng::seed_from_u64(seed)),
 * so one rank of a point-sharded MSM builds only its shard. */
int zkgpu_setup_powers(uint64_t seed, uint64_t start, size_t count, uint64_t* g_out);

/* halo2_proofs::arithmetic::eval_polynomial(poly, point): sum_i coeffs[i] * x^i. */
int zkgpu_eval_polynomial(const uint64_t* coeffs, size_t n, const uint64_t x[4], uint64_t out[4]);

/* Counts the points that are neither the identity (0,0) nor on y^2 = x^3 + 3: the `G1Affine::from_xy(..).unwrap()`
 * check of the ptau reader (/root/reference/crates/powers-of-tau/lib.rs:206-224). */
int zkgpu_g1_on_curve(const uint64_t* points_affine, size_t n, uint64_t* bad_count);

/* Host-side sum of n affine points: the combine step of a point-sharded MSM (one partial result per GPU,
 * gathered by the caller).  Needs no GPU. */
int zkgpu_g1_sum_affine(const uint64_t* points_affine, size_t n, uint64_t out_affine[8]);

/* ---- vectorised Fr helpers (halo2curves `Fr` ops / `Fr::random`), used to build synthetic circuits and
 * witnesses in bench.py without the CPU oracle.  op: 0 mul, 1 add, 2 sub, 3 to Montgomery, 4 from Montgomery. */
int zkgpu_fr_vec_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
int zkgpu_fr_to_mont(const uint64_t* canonical, uint64_t* out, size_t n);
int zkgpu_fr_from_mont(const uint64_t* mont, uint64_t* out, size_t n);
int zkgpu_fr_random(uint64_t seed, uint64_t* out, size_t n);
/* n x `Fr::random(&mut rng)` from a running SmallRng (state advanced in place) */
int zkgpu_fr_random_rng(uint64_t rng_state[4], uint64_t* out, size_t n);

/* ---- Poseidon2 (t = 8, rate 7) and the note-tree Merkle path: witness-side hashing --------------------------------
 * shielder_bindings::hash::poseidon_hash = hash_variable_length (/root/reference/crates/shielder_bindings/src/hash.rs:16-27,
 * src/utils.rs:14-30): m hashes of `len` (1..7; anything else is an error, as upstream panics) field elements each, the same
 * permutation as the on-chain Poseidon2T8Assembly (/root/reference/poseidon2-solidity/generate_t8.py).  inputs m x len x 4 u64,
 * out m x 4 u64, Montgomery form like every other field buffer. */
int zkgpu_poseidon2_hash_batch(const uint64_t* inputs, size_t len, size_t m, uint64_t* out);
int zkgpu_poseidon2_hash_batch_dev(const void* d_inputs, size_t len, size_t m, void* d_out, void* stream);
/* m Merkle paths in the layout of MerkleTree.getMerklePath without the trailing root
 * (/root/reference/contracts/MerkleTree.sol:88-113; `vec_to_path`, shielder_bindings/src/utils.rs:36-41): height x 7 elements,
 * leaf level first.  roots[i] = hash(top level); consistent[i] (may be NULL) = 1 iff every level contains the hash of the
 * level below it, i.e. the path is one `_addNote` could have produced (:134-147). */
int zkgpu_merkle_root_batch(const uint64_t* paths, size_t height, size_t m, uint64_t* roots, uint8_t* consistent);

/* ---- device-resident variants (inputs already in HBM; `stream` is a cudaStream_t or NULL) ------
 * Used by the prover pipeline and by bench.py's kernel-only ("value") measurement. */
int zkgpu_ntt_fr_batch_dev(void* d_a, const uint64_t omega[4], uint32_t log_n, size_t m, void* d_scratch, void* stream);
int zkgpu_msm_g1_srs_batch_dev(uint64_t srs, int basis, const void* d_scalars, size_t n, size_t m, void* d_out_affine, void* stream);
/* ---- batched prover: halo2_proofs::plonk::{keygen_vk, keygen_pk, create_proof} -------------------
 * The funnel every Shielder prover host goes through is
 *   shielder_circuits::generate_proof(&params, &pk, circuit, &public_input, rng) -> Vec<u8>
 * (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111,
 *  /root/reference/crates/shielder-account/src/call_data.rs:489-501,
 *  /root/reference/tee/crates/shielder-prover-tee/src/circuits/mod.rs:70-78), one circuit + one instance
 * column per proof, Keccak256 EVM transcript, SHPLONK.  zkgpu_prove_batch runs m such proofs of the same
 * circuit in lock step on the GPU.
 *
 * zkgpu_pk_create = generate_keys_with_min_k's keygen_vk + keygen_pk for a fixed k
 * (/root/reference/crates/shielder_bindings/build.rs:22): `circuit_blob` is the serialised constraint
 * system + fixed assignment + copy constraints (layout: zkgpu/circuits.py Circuit._serialize), i.e. what
 * halo2's `ConstraintSystem` and keygen `Assembly` hold after `Circuit::configure` / `synthesize`.
 * The SRS handle must have been registered with the same k (ParamsKZG::downsize first). */
int zkgpu_pk_create(uint64_t srs, const uint8_t* circuit_blob, size_t blob_len, uint64_t* pk_out);
/* ProvingKey::read for the artefact the reference ships: `pk_bin` = `k: u32 LE` ‖ `ProvingKey::to_bytes(RawBytesUnchecked)`
 * (`marshall_pk`: /root/reference/crates/shielder_bindings/build.rs:19-33, read at src/circuits/mod.rs:35-48,89-101, cached by
 * /root/reference/crates/shielder-cli/src/shielder_ops/pk.rs:68-126).  The file's fixed / permutation values, coefficient forms and
 * extended cosets and the verifying key's commitments are used as they are (nothing is recomputed).  Upstream rebuilds the constraint
 * system from the circuit TYPE; here `cs_blob` carries it: the same layout as `circuit_blob` with magic 0x5a4b4354, without the
 * fixed assignment and copy constraints, followed by num_selectors: u32 and vk.transcript_repr() (32 bytes) — written by the Rust
 * exporter shown in INTEGRATION.md (layout: zkgpu/circuits.py Circuit.cs_blob). */
int zkgpu_pk_load(uint64_t srs, const uint8_t* cs_blob, size_t cs_blob_len, const uint8_t* pk_bin, size_t pk_bin_len, uint64_t* pk_out);
int zkgpu_pk_release(uint64_t pk);
/* info[0..14] = k, n, num_advice, num_fixed, degree, blinding_factors, num_perm_sets, num_quotients,
 *               num_evals, proof_len, extended_k, num_perm_columns, num_rotation_sets, sub_batch, replicas (devices) */
int zkgpu_pk_info(uint64_t pk, uint64_t info[16]);
/* VerifyingKey parts: fixed commitments (F x 8 u64 affine), permutation commitments (S x 8), transcript_repr (4) */
int zkgpu_pk_vk(uint64_t pk, uint64_t* fixed_commitments, uint64_t* perm_commitments, uint64_t digest[4]);
/* ---- the proof's randomness: `rng: &mut impl RngCore` of generate_proof --------------------------------------------
 * The reference passes a RUNNING `SmallRng` in its seeded tests — the same rng has already produced the SRS and the witness
 * (/root/reference/crates/halo2-verifier/src/generator.rs:117-130) — and `OsRng` / `thread_rng()` in production
 * (/root/reference/crates/shielder-account/src/call_data.rs:499, /root/reference/crates/shielder_bindings/src/circuits/deposit.rs:108).
 * rng_mode selects what `rng_data` holds for each proof:
 *   ZKGPU_RNG_SEED_U64      m x u64        SmallRng::seed_from_u64(seed): TEST / PARITY ONLY — 64 bits of entropy and a
 *                                          non-cryptographic generator do not hide a witness.
 *   ZKGPU_RNG_XOSHIRO_STATE m x 4 x u64    a running SmallRng (xoshiro256++ state), IN/OUT: the state after the proof's last
 *                                          draw is written back, so the host's rng continues exactly as `&mut rng` does
 *                                          upstream.  Reproduces the reference's seeded prove-and-verify tests.  Not a CSPRNG.
 *   ZKGPU_RNG_CHACHA20_SEED m x 32 bytes   ChaCha20Rng::from_seed(seed) (rand_chacha 0.3.1): PRODUCTION — the caller draws 32
 *                                          bytes per proof from `OsRng` / `thread_rng()`; every blinding value of the proof comes
 *                                          from that 256-bit-keyed ChaCha20 stream. */
#define ZKGPU_RNG_SEED_U64 0
#define ZKGPU_RNG_XOSHIRO_STATE 1
#define ZKGPU_RNG_CHACHA20_SEED 2
/* per-proof status */
#define ZKGPU_PROOF_OK 0
#define ZKGPU_PROOF_LOOKUP_FAILED 1   /* a lookup input is not in its table: plonk::Error::ConstraintSystemFailure for that proof */

/* m proofs.  advice: m x num_advice x n field elements — the assigned advice columns `create_proof` holds
 * after witness synthesis (rows >= n - (blinding_factors + 1) are overwritten with blinding values);
 * instance: m x num_instance public inputs; proofs_out: m x proof_len bytes, each exactly what
 * `transcript.finalize()` returns (/root/reference/crates/halo2-verifier/src/lib/verifier_contract.rs:14-20).
 * With several selected devices the batch is split into contiguous shards, one per device.
 * A witness that makes create_proof fail fails ALONE, as one request of the reference's prover server does
 * (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:189-190): status_out[i] (may be NULL) receives its
 * ZKGPU_PROOF_* code, its proof bytes are zeroed, and the other m - 1 proofs are complete and valid; the call returns 0. */
int zkgpu_prove_batch_rng(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, size_t m,
                          int rng_mode, void* rng_data, uint8_t* proofs_out, size_t proof_len, int32_t* status_out);
/* same with the advice columns already resident in HBM of one selected device (that device proves the whole batch) */
int zkgpu_prove_batch_rng_dev(uint64_t pk, const void* d_advice, const uint64_t* instance, size_t num_instance, size_t m,
                              int rng_mode, void* rng_data, uint8_t* proofs_out, size_t proof_len, int32_t* status_out);
/* Test / parity form of the above: rng_mode = ZKGPU_RNG_SEED_U64 (rng_seeds[i] seeds proof i,
 * /root/reference/crates/shielder-setup/lib.rs:29-40) and all-or-nothing: ZKGPU_ERR_ARG if any proof failed. */
int zkgpu_prove_batch(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, size_t m,
                      const uint64_t* rng_seeds, uint8_t* proofs_out, size_t proof_len);
/* ONE proof, blocking: the call a per-request host makes where the reference calls generate_proof — one tokio task per
 * client, up to 100 in flight (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:157-195,
 * /root/reference/tee/crates/shielder-prover-server/src/command_line_args.rs:24-27).  Callable from any number of threads:
 * requests waiting at the same time are coalesced into lock-step batches across every selected device; no caller waits on a
 * lock around the GPU.  rng_data as above for ONE proof.  A failing witness returns ZKGPU_ERR_WITNESS to its caller only. */
int zkgpu_prove(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, int rng_mode, void* rng_data,
                uint8_t* proof_out, size_t proof_len);
/* coalescer statistics of a proving key: out = {requests served, batches run, largest batch, dispatcher threads} */
int zkgpu_prove_stats(uint64_t pk, uint64_t out[4]);
/* halo2's vanishing prover draws the random polynomial in chunks of n / rayon::current_num_threads() coefficients, each from its
 * own ChaCha20 stream seeded from the proof's rng, so proof bytes depend on the host's rayon thread count.  Tell the
 * library the thread count of the host it replaces (default 1 = a single stream, the `multicore` feature off / RAYON_NUM_THREADS=1). */
int zkgpu_set_rayon_threads(unsigned num_threads);
/* zkgpu_prove_batch with the advice columns already resident in HBM */
int zkgpu_prove_batch_dev(uint64_t pk, const void* d_advice, const uint64_t* instance, size_t num_instance, size_t m,
                          const uint64_t* rng_seeds, uint8_t* proofs_out, size_t proof_len);
/* profiling aid: accumulated wall-clock seconds per prover step (0 upload, 1 advice commit, 2 permutation +
 * random poly, 3 quotient, 4 evaluations, 5 SHPLONK h, 6 SHPLONK L); reset != 0 zeroes the counters */
void zkgpu_prover_step_seconds(double out[8], int reset);
/* test hook: called with (stage name, field elements of proof 0 of each sub-batch, byte count) */
void zkgpu_set_trace(void (*fn)(const char* name, const void* data, size_t bytes));

/* Per-kernel-class device timing (bench.py's roofline line): when enabled, CUDA events are recorded on the
 * launching stream around every launch group of a class.  Slots: 0 MSM bucket accumulation, 1 MSM digit
 * sort (count/scan/scatter), 2 MSM bucket reduction, 3 NTT tile passes, 4 quotient evaluation,
 * 5 permutation / lookup grand products, 6 polynomial evaluation / SHPLONK algebra, 7 lookup compression +
 * permuted columns, 8 the rest (blinding scatter, ChaCha20 polynomial, affine normalisation), 9 device idle time at the
 * Fiat-Shamir round trips (the host hashes the transcript; hidden by the other pipeline workers outside this timing mode). */
void zkgpu_kernel_timing(int enable);
int zkgpu_kernel_times(int slot, double* total_ms, uint64_t* launches, int reset);
/* mixed point additions queued for bucket accumulation by the batched fixed-base MSM path since the last reset, summed over the
 * devices (zero digits are skipped and a column may be committed through its differences, so the count depends on the data;
 * bench.py's roofline uses it as the algorithmic work of the dominant kernel).  Synchronises the devices. */
int zkgpu_msm_additions(uint64_t* total, int reset);
/* the library's CUDA stream (cudaStream_t) on the primary device after zkgpu_init, for event timing by the caller */
void* zkgpu_stream(void);
/* number of kernel launches issued by this library in this process so far */
uint64_t zkgpu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ZKGPU_H */
