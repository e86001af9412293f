//! Raw bindings to `libzkgpu.so` — one `extern "C"` item per entry point of `include/zkgpu.h`.
//!
//! Memory layout contract: `bn256::Fr` / `Fq` are four little-endian `u64` limbs in Montgomery form and
//! `G1Affine` is `{x: Fq, y: Fq}` (identity = (0, 0)), so slices of them pass as `*const u64` without copying.
//! Every function returns 0 on success; `check` turns anything else into a panic, which is how the upstream
//! functions behave on misuse (`assert_eq!(coeffs.len(), bases.len())`).
use std::os::raw::{c_char, c_int, c_void};

extern "C" {
    /// bit i of `device_mask` selects CUDA device i; 0 = every visible device
    pub fn zkgpu_init(device_mask: c_int) -> c_int;
    pub fn zkgpu_device_count() -> c_int;
    pub fn zkgpu_device_index(slot: c_int) -> c_int;
    pub fn zkgpu_shutdown();
    pub fn zkgpu_last_error() -> *const c_char;
    pub fn zkgpu_abi_version() -> c_int;
    pub fn zkgpu_keccak256(data: *const u8, len: usize, out: *mut u8) -> c_int;

    pub fn zkgpu_msm_g1(scalars: *const u64, bases_affine: *const u64, n: usize, out_jacobian: *mut u64) -> c_int;
    pub fn zkgpu_srs_register(g: *const u64, g_lagrange: *const u64, k: u32, handle_out: *mut u64) -> c_int;
    pub fn zkgpu_srs_release(srs: u64) -> c_int;
    pub fn zkgpu_msm_g1_srs(srs: u64, basis: c_int, scalars: *const u64, n: usize, out_jacobian: *mut u64) -> c_int;
    pub fn zkgpu_msm_g1_srs_batch(srs: u64, basis: c_int, scalars: *const u64, n: usize, m: usize, out_affine: *mut u64) -> c_int;

    pub fn zkgpu_ntt_fr(a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn zkgpu_ntt_fr_batch(a: *mut u64, omega: *const u64, log_n: u32, m: usize) -> c_int;
    pub fn zkgpu_domain_ntt_fr(a: *mut u64, k: u32, inverse: c_int, m: usize) -> c_int;
    pub fn zkgpu_coset_ntt_fr(coeffs: *const u64, k: u32, ext_k: u32, out: *mut u64) -> c_int;
    pub fn zkgpu_coset_intt_fr(evals: *mut u64, k: u32, ext_k: u32, quotient_degree: u32) -> c_int;

    pub fn zkgpu_fft_g1(points_jacobian: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn zkgpu_g_to_lagrange(g_affine: *const u64, k: u32, out_affine: *mut u64) -> c_int;
    pub fn zkgpu_params_setup(k: u32, seed: u64, g_out: *mut u64, g_lagrange_out: *mut u64) -> c_int;
    pub fn zkgpu_params_setup_rng(k: u32, rng_state: *mut u64, g_out: *mut u64, g_lagrange_out: *mut u64) -> c_int;
    pub fn zkgpu_fr_random_rng(rng_state: *mut u64, out: *mut u64, n: usize) -> c_int;
    pub fn zkgpu_setup_powers(seed: u64, start: u64, count: usize, g_out: *mut u64) -> c_int;
    pub fn zkgpu_eval_polynomial(coeffs: *const u64, n: usize, x: *const u64, out: *mut u64) -> c_int;
    pub fn zkgpu_bases_register(bases_affine: *const u64, n: usize, handle_out: *mut u64) -> c_int;
    pub fn zkgpu_bases_release(bases: u64) -> c_int;
    pub fn zkgpu_msm_g1_bases(bases: u64, scalars: *const u64, n: usize, out_jacobian: *mut u64, kernel_ms: *mut f64) -> c_int;
    pub fn zkgpu_g1_sum_affine(points_affine: *const u64, n: usize, out_affine: *mut u64) -> c_int;
    pub fn zkgpu_g1_on_curve(points_affine: *const u64, n: usize, bad_count: *mut u64) -> c_int;

    pub fn zkgpu_poseidon2_hash_batch(inputs: *const u64, len: usize, m: usize, out: *mut u64) -> c_int;
    pub fn zkgpu_poseidon2_hash_batch_dev(d_inputs: *const c_void, len: usize, m: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn zkgpu_merkle_root_batch(paths: *const u64, height: usize, m: usize, roots: *mut u64, consistent: *mut u8) -> c_int;

    pub fn zkgpu_pk_create(srs: u64, circuit_blob: *const u8, blob_len: usize, pk_out: *mut u64) -> c_int;
    pub fn zkgpu_pk_load(srs: u64, cs_blob: *const u8, cs_blob_len: usize, pk_bin: *const u8, pk_bin_len: usize, pk_out: *mut u64) -> c_int;
    pub fn zkgpu_pk_release(pk: u64) -> c_int;
    pub fn zkgpu_pk_info(pk: u64, info: *mut u64) -> c_int;
    pub fn zkgpu_pk_vk(pk: u64, fixed_commitments: *mut u64, perm_commitments: *mut u64, digest: *mut u64) -> c_int;
    pub fn zkgpu_prove_batch(pk: u64, advice: *const u64, instance: *const u64, num_instance: usize, m: usize,
                             rng_seeds: *const u64, proofs_out: *mut u8, proof_len: usize) -> c_int;
    pub fn zkgpu_set_rayon_threads(num_threads: u32) -> c_int;
    /// rng_mode: RNG_SEED_U64 (tests), RNG_XOSHIRO_STATE (running SmallRng, state written back), RNG_CHACHA20_SEED (production)
    pub fn zkgpu_prove_batch_rng(pk: u64, advice: *const u64, instance: *const u64, num_instance: usize, m: usize, rng_mode: c_int,
                                 rng_data: *mut c_void, proofs_out: *mut u8, proof_len: usize, status_out: *mut i32) -> c_int;
    pub fn zkgpu_prove_batch_rng_dev(pk: u64, d_advice: *const c_void, instance: *const u64, num_instance: usize, m: usize, rng_mode: c_int,
                                     rng_data: *mut c_void, proofs_out: *mut u8, proof_len: usize, status_out: *mut i32) -> c_int;
    /// one blocking proof; concurrent callers are coalesced into batches inside the library
    pub fn zkgpu_prove(pk: u64, advice: *const u64, instance: *const u64, num_instance: usize, rng_mode: c_int, rng_data: *mut c_void,
                       proof_out: *mut u8, proof_len: usize) -> c_int;
    pub fn zkgpu_prove_stats(pk: u64, out: *mut u64) -> c_int;
    pub fn zkgpu_prove_batch_dev(pk: u64, d_advice: *const c_void, instance: *const u64, num_instance: usize, m: usize,
                                 rng_seeds: *const u64, proofs_out: *mut u8, proof_len: usize) -> c_int;
}

pub const RNG_SEED_U64: c_int = 0;
pub const RNG_XOSHIRO_STATE: c_int = 1;
pub const RNG_CHACHA20_SEED: c_int = 2;
pub const ERR_WITNESS: c_int = -5;

/// Panics with the library's thread-local message on a non-zero return code.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(zkgpu_last_error()) }.to_string_lossy().into_owned();
        panic!("zkgpu error {rc}: {msg}");
    }
}
