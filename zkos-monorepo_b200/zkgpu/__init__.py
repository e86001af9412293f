"""zkgpu — Python host mirror of the reference-facing prover interface, over libzkgpu's C ABI.

The reference's host side is Rust (no cargo in this image), so the product's host logic lives in C++
inside ``libzkgpu.so``; this module is only the thin ctypes binding used by the tests and the
benchmark.  Names follow the upstream halo2 / halo2curves API that Shielder reaches through
``shielder_circuits::generate_proof`` (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111):

    best_multiexp(coeffs, bases)          halo2curves::msm::best_multiexp
    best_fft(a, omega, log_n)             halo2curves::fft::best_fft
    ParamsKZG(k, g, g_lagrange).commit / .commit_lagrange      halo2_proofs poly/kzg/commitment.rs
    EvaluationDomain(j, k).{lagrange_to_coeff, coeff_to_lagrange, coeff_to_extended, extended_to_coeff}
    g_to_lagrange(g, k)                   ParamsKZG::from_parts(.., None, ..)  (powers-of-tau/lib.rs:71)

Arrays are numpy uint64 in the Rust memory layout: field elements (..., 4) Montgomery limbs,
affine points (..., 8).  There is NO CPU fallback: if the CUDA library is missing or no GPU is
visible, every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libzkgpu.so")
_lib = None


class ZkGpuError(RuntimeError):
    pass


def lib():
    """Loads libzkgpu.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZkGpuError("libzkgpu.so not built (%s); run __graft_entry__.build() — there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.zkgpu_last_error.restype = C.c_char_p
        _lib.zkgpu_launch_count.restype = C.c_uint64
    return _lib


def _chk(rc):
    if rc != 0:
        raise ZkGpuError("libzkgpu error %d: %s" % (rc, lib().zkgpu_last_error().decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def init(device=0, mask=None):
    """zkgpu_init(device_mask): `device` selects one CUDA device (mask 1 << device); `mask` selects several (bit i = device i,
    0 = every visible device) for the single-process multi-GPU mode."""
    _chk(lib().zkgpu_init(int(mask) if mask is not None else 1 << int(device)))


def device_count():
    return int(lib().zkgpu_device_count())


def shutdown():
    lib().zkgpu_shutdown()


def launch_count():
    return int(lib().zkgpu_launch_count())


def _keccak256(data):
    out = C.create_string_buffer(32)
    _chk(lib().zkgpu_keccak256(bytes(data), C.c_size_t(len(data)), out))
    return out.raw


def _jac_to_affine(j):
    """normalised Jacobian (x, y, 1) / (0, 1, 0) -> affine (x, y) / (0, 0)"""
    out = np.zeros(8, dtype=np.uint64)
    if j[8:].any():
        out[:] = j[:8]
    return out


def best_multiexp(coeffs, bases):
    """sum_i coeffs[i] * bases[i]; returns the affine-normalised point (8 u64)."""
    coeffs, bases = _u64(coeffs), _u64(bases)
    n = coeffs.size // 4
    if bases.size // 8 != n:
        raise ZkGpuError("best_multiexp: coeffs.len() != bases.len()")  # upstream assert_eq!
    out = np.zeros(12, dtype=np.uint64)
    _chk(lib().zkgpu_msm_g1(_p(coeffs), _p(bases), C.c_size_t(n), _p(out)))
    return _jac_to_affine(out)


def best_fft(a, omega, log_n, batch=1):
    """In-place semantics of halo2curves best_fft; returns the transformed copy."""
    a = np.array(a, dtype=np.uint64, copy=True)
    omega = _u64(omega)
    if a.size != 4 * batch << log_n:
        raise ZkGpuError("best_fft: a.len() != 1 << log_n")
    _chk(lib().zkgpu_ntt_fr_batch(_p(a), _p(omega), C.c_uint32(log_n), C.c_size_t(batch)))
    return a


def g_to_lagrange(g, k):
    g = _u64(g)
    out = np.empty((1 << k, 8), dtype=np.uint64)
    _chk(lib().zkgpu_g_to_lagrange(_p(g), C.c_uint32(k), _p(out)))
    return out


def fft_g1(points_jacobian, omega, log_n):
    pts = np.array(points_jacobian, dtype=np.uint64, copy=True)
    _chk(lib().zkgpu_fft_g1(_p(pts), _p(_u64(omega)), C.c_uint32(log_n)))
    return pts


def params_setup(k, seed, lagrange=True):
    """ParamsKZG::setup(k, SmallRng::seed_from_u64(seed)) -> (g, g_lagrange) as (n, 8) uint64 arrays
    (g_lagrange is None with lagrange=False, which allows k up to 24)"""
    n = 1 << k
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64) if lagrange else None
    _chk(lib().zkgpu_params_setup(C.c_uint32(k), C.c_uint64(seed), _p(g), _p(gl) if lagrange else None))
    return g, gl


def params_setup_rng(k, rng_state):
    """ParamsKZG::setup(k, &mut rng) with a running SmallRng: rng_state (4,) uint64 is advanced in place"""
    n = 1 << k
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64)
    _chk(lib().zkgpu_params_setup_rng(C.c_uint32(k), _p(rng_state), _p(g), _p(gl)))
    return g, gl


def fr_random_rng(rng_state, n):
    """n x Fr::random(&mut rng) from a running SmallRng (state advanced in place)"""
    out = np.empty((n, 4), dtype=np.uint64)
    _chk(lib().zkgpu_fr_random_rng(_p(rng_state), _p(out), C.c_size_t(n)))
    return out


def g1_sum(points):
    """host-side sum of affine points (n, 8) -> (8,)"""
    points = _u64(points)
    out = np.zeros(8, dtype=np.uint64)
    _chk(lib().zkgpu_g1_sum_affine(_p(points), C.c_size_t(points.size // 8), _p(out)))
    return out


def setup_powers(seed, start, count):
    """g[i] = s^(start + i) * G, i < count: a slice of ParamsKZG::setup(.., SmallRng::seed_from_u64(seed)).g"""
    g = np.empty((count, 8), dtype=np.uint64)
    _chk(lib().zkgpu_setup_powers(C.c_uint64(seed), C.c_uint64(start), C.c_size_t(count), _p(g)))
    return g


def eval_polynomial(coeffs, x):
    """halo2_proofs::arithmetic::eval_polynomial: sum_i coeffs[i] x^i"""
    coeffs, x = _u64(coeffs), _u64(x)
    out = np.empty(4, dtype=np.uint64)
    _chk(lib().zkgpu_eval_polynomial(_p(coeffs), C.c_size_t(coeffs.size // 4), _p(x), _p(out)))
    return out


class Bases:
    """best_multiexp over RESIDENT bases: the points are split into contiguous shards, one per selected device (a device only
    ever holds its shard); every MSM reduces each shard to one point and sums the partial points on the primary device."""

    def __init__(self, bases):
        bases = _u64(bases)
        self.n = bases.size // 8
        h = C.c_uint64(0)
        _chk(lib().zkgpu_bases_register(_p(bases), C.c_size_t(self.n), C.byref(h)))
        self.handle = h.value
        self.kernel_ms = 0.0

    def msm(self, scalars=None):
        """scalars (n, 4) in host memory, or None to reuse the scalars the previous call left in HBM; returns the affine point"""
        out = np.zeros(12, dtype=np.uint64)
        ms = C.c_double(0)
        ptr = _p(_u64(scalars)) if scalars is not None else None
        _chk(lib().zkgpu_msm_g1_bases(C.c_uint64(self.handle), ptr, C.c_size_t(self.n), _p(out), C.byref(ms)))
        self.kernel_ms = ms.value
        return _jac_to_affine(out)

    def release(self):
        if self.handle:
            lib().zkgpu_bases_release(C.c_uint64(self.handle))
            self.handle = 0


class ParamsKZG:
    """Device-resident SRS: ParamsKZG::{commit, commit_lagrange}; `Blind` is ignored by KZG."""

    def __init__(self, k, g, g_lagrange):
        self.k, self.n = int(k), 1 << int(k)
        g, g_lagrange = _u64(g), _u64(g_lagrange)
        if g.size != 8 * self.n or g_lagrange.size != 8 * self.n:
            raise ZkGpuError("ParamsKZG: g / g_lagrange must hold 2^k points")
        h = C.c_uint64(0)
        _chk(lib().zkgpu_srs_register(_p(g), _p(g_lagrange), C.c_uint32(self.k), C.byref(h)))
        self.handle = h.value

    def _msm(self, basis, scalars):
        scalars = _u64(scalars)
        n = scalars.size // 4
        out = np.zeros(12, dtype=np.uint64)
        _chk(lib().zkgpu_msm_g1_srs(C.c_uint64(self.handle), basis, _p(scalars), C.c_size_t(n), _p(out)))
        return _jac_to_affine(out)

    def commit(self, poly_coeffs, blind=None):
        return self._msm(0, poly_coeffs)

    def commit_lagrange(self, poly_values, blind=None):
        return self._msm(1, poly_values)

    def commit_batch(self, basis, scalars, n):
        """m commitments over the same basis in one call; scalars (m, n, 4) -> (m, 8) affine"""
        scalars = _u64(scalars)
        m = scalars.size // (4 * n)
        out = np.empty((m, 8), dtype=np.uint64)
        _chk(lib().zkgpu_msm_g1_srs_batch(C.c_uint64(self.handle), basis, _p(scalars), C.c_size_t(n), C.c_size_t(m), _p(out)))
        return out

    def commit_batch_dev(self, basis, d_scalars_ptr, n, m, d_out_ptr, stream=0):
        _chk(lib().zkgpu_msm_g1_srs_batch_dev(C.c_uint64(self.handle), basis, C.c_void_p(d_scalars_ptr), C.c_size_t(n),
                                              C.c_size_t(m), C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def release(self):
        if self.handle:
            lib().zkgpu_srs_release(C.c_uint64(self.handle))
            self.handle = 0


class EvaluationDomain:
    """halo2_proofs poly/domain.rs EvaluationDomain::new(j, k)."""

    def __init__(self, j, k):
        self.j, self.k = int(j), int(k)
        self.n = 1 << self.k
        self.quotient_poly_degree = self.j - 1
        self.extended_k = self.k
        while (1 << self.extended_k) < self.n * self.quotient_poly_degree:
            self.extended_k += 1

    def extended_len(self):
        return 1 << self.extended_k

    def _domain_ntt(self, a, inverse):
        a = np.array(a, dtype=np.uint64, copy=True)
        m = a.size // (4 * self.n)
        if m * 4 * self.n != a.size:
            raise ZkGpuError("polynomial length must be a multiple of 2^k")
        _chk(lib().zkgpu_domain_ntt_fr(_p(a), C.c_uint32(self.k), int(inverse), C.c_size_t(m)))
        return a

    def lagrange_to_coeff(self, a):
        return self._domain_ntt(a, 1)

    def coeff_to_lagrange(self, a):
        return self._domain_ntt(a, 0)

    def coeff_to_extended(self, a):
        a = _u64(a)
        if a.size != 4 * self.n:
            raise ZkGpuError("coeff_to_extended: expected 2^k coefficients")
        out = np.empty((self.extended_len(), 4), dtype=np.uint64)
        _chk(lib().zkgpu_coset_ntt_fr(_p(a), C.c_uint32(self.k), C.c_uint32(self.extended_k), _p(out)))
        return out

    def extended_to_coeff(self, a):
        a = np.array(a, dtype=np.uint64, copy=True)
        if a.size != 4 * self.extended_len():
            raise ZkGpuError("extended_to_coeff: expected 2^extended_k evaluations")
        _chk(lib().zkgpu_coset_intt_fr(_p(a), C.c_uint32(self.k), C.c_uint32(self.extended_k), C.c_uint32(self.quotient_poly_degree)))
        return a.reshape(-1, 4)[: self.n * self.quotient_poly_degree]


def ntt_batch_dev(d_ptr, omega, log_n, m, d_scratch_ptr=0, stream=0):
    _chk(lib().zkgpu_ntt_fr_batch_dev(C.c_void_p(d_ptr), _p(_u64(omega)), C.c_uint32(log_n), C.c_size_t(m),
                                      C.c_void_p(d_scratch_ptr), C.c_void_p(stream)))


class ProvingKey:
    """Device-resident halo2 ProvingKey for one circuit: keygen_vk + keygen_pk over a ParamsKZG of the
    same k (what `generate_keys_with_min_k` returns, /root/reference/crates/shielder_bindings/build.rs:22),
    and the batched `generate_proof` funnel
    (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111)."""

    INFO = ("k", "n", "num_advice", "num_fixed", "degree", "blinding_factors", "num_perm_sets", "num_quotients",
            "num_evals", "proof_len", "extended_k", "num_perm_columns", "num_rotation_sets", "sub_batch", "replicas")

    def __init__(self, params, circuit_blob, pk_bin=None):
        """keygen from a circuit blob, or — with `pk_bin` — `unmarshall_pk`: load the reference's pk.bin next to a
        constraint-system-only blob (circuits.Circuit.cs_blob)"""
        self.params = params
        h = C.c_uint64(0)
        blob = bytes(circuit_blob)
        if pk_bin is None:
            _chk(lib().zkgpu_pk_create(C.c_uint64(params.handle), blob, C.c_size_t(len(blob)), C.byref(h)))
        else:
            pk_bin = bytes(pk_bin)
            _chk(lib().zkgpu_pk_load(C.c_uint64(params.handle), blob, C.c_size_t(len(blob)), pk_bin, C.c_size_t(len(pk_bin)), C.byref(h)))
        self.handle = h.value
        info = np.zeros(16, dtype=np.uint64)
        _chk(lib().zkgpu_pk_info(C.c_uint64(self.handle), _p(info)))
        for name, v in zip(self.INFO, info):
            setattr(self, name, int(v))

    def vk(self):
        fc = np.zeros((self.num_fixed, 8), dtype=np.uint64)
        pc = np.zeros((self.num_perm_columns, 8), dtype=np.uint64)
        dg = np.zeros(4, dtype=np.uint64)
        _chk(lib().zkgpu_pk_vk(C.c_uint64(self.handle), _p(fc), _p(pc), _p(dg)))
        return fc, pc, dg

    def prove_batch(self, advice, instance, seeds):
        """advice (m, A, n, 4), instance (m, num_pi, 4), seeds (m,) -> list of m proofs (bytes)"""
        advice, instance = _u64(advice), _u64(instance)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        m = seeds.size
        if advice.size != m * self.num_advice * self.n * 4:
            raise ZkGpuError("prove_batch: advice must be m x num_advice x n field elements")
        num_pi = instance.size // (4 * m) if m else 0
        out = np.zeros(m * self.proof_len, dtype=np.uint8)
        _chk(lib().zkgpu_prove_batch(C.c_uint64(self.handle), _p(advice), _p(instance), C.c_size_t(num_pi), C.c_size_t(m),
                                     _p(seeds), _p(out), C.c_size_t(self.proof_len)))
        raw = out.tobytes()
        return [raw[i * self.proof_len:(i + 1) * self.proof_len] for i in range(m)]

    RNG_SEED_U64, RNG_XOSHIRO_STATE, RNG_CHACHA20_SEED = 0, 1, 2
    PROOF_OK, PROOF_LOOKUP_FAILED = 0, 1

    def prove_batch_rng(self, advice, instance, rng_mode, rng_data):
        """zkgpu_prove_batch_rng: rng_data is a C-contiguous numpy array — (m,) uint64 seeds, (m, 4) uint64 running SmallRng
        states (advanced IN PLACE) or (m, 32) uint8 ChaCha20 seeds.  Returns (proofs, status): a failed proof is b"" """
        advice, instance = _u64(advice), _u64(instance)
        m = {0: rng_data.size, 1: rng_data.size // 4, 2: rng_data.size // 32}[rng_mode]
        if advice.size != m * self.num_advice * self.n * 4:
            raise ZkGpuError("prove_batch: advice must be m x num_advice x n field elements")
        if not rng_data.flags["C_CONTIGUOUS"]:
            raise ZkGpuError("rng_data must be C-contiguous (it is updated in place)")
        num_pi = instance.size // (4 * m) if m else 0
        out = np.zeros(m * self.proof_len, dtype=np.uint8)
        status = np.zeros(m, dtype=np.int32)
        _chk(lib().zkgpu_prove_batch_rng(C.c_uint64(self.handle), _p(advice), _p(instance), C.c_size_t(num_pi), C.c_size_t(m),
                                         int(rng_mode), _p(rng_data), _p(out), C.c_size_t(self.proof_len), _p(status)))
        raw = out.tobytes()
        return [raw[i * self.proof_len:(i + 1) * self.proof_len] if status[i] == 0 else b"" for i in range(m)], status

    def prove_one(self, advice, instance, rng_mode, rng_data):
        """zkgpu_prove: one blocking proof; concurrent callers (threads) are coalesced into batches by the library"""
        advice, instance = _u64(advice), _u64(instance)
        out = np.zeros(self.proof_len, dtype=np.uint8)
        _chk(lib().zkgpu_prove(C.c_uint64(self.handle), _p(advice), _p(instance), C.c_size_t(instance.size // 4), int(rng_mode),
                               _p(rng_data), _p(out), C.c_size_t(self.proof_len)))
        return out.tobytes()

    def prove_stats(self):
        out = np.zeros(4, dtype=np.uint64)
        _chk(lib().zkgpu_prove_stats(C.c_uint64(self.handle), _p(out)))
        return dict(requests=int(out[0]), batches=int(out[1]), max_batch=int(out[2]), dispatchers=int(out[3]))

    def prove_batch_dev(self, d_advice_ptr, instance, seeds, out=None):
        """advice already resident in HBM (device pointer); returns the proofs as one uint8 array"""
        instance = _u64(instance)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        m = seeds.size
        num_pi = instance.size // (4 * m) if m else 0
        if out is None:
            out = np.zeros(m * self.proof_len, dtype=np.uint8)
        _chk(lib().zkgpu_prove_batch_dev(C.c_uint64(self.handle), C.c_void_p(d_advice_ptr), _p(instance), C.c_size_t(num_pi),
                                         C.c_size_t(m), _p(seeds), _p(out), C.c_size_t(self.proof_len)))
        return out

    def prove(self, advice, instance, seed):
        """generate_proof(params, pk, circuit, public_input, rng) for one proof"""
        return self.prove_batch(np.asarray(advice)[None], np.asarray(instance)[None], [seed])[0]

    def release(self):
        if self.handle:
            lib().zkgpu_pk_release(C.c_uint64(self.handle))
            self.handle = 0


TRACE_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p, C.c_size_t)


def set_rayon_threads(num_threads):
    """rayon::current_num_threads() of the host being replaced (chunking of the vanishing argument's random polynomial)."""
    _chk(lib().zkgpu_set_rayon_threads(C.c_uint(int(num_threads))))


def set_trace(fn):
    """fn(name: bytes, data_ptr, nbytes) per prover stage of proof 0, or None to disable.  Keep the returned
    object alive while tracing."""
    cb = TRACE_FN(fn) if fn is not None else C.cast(None, TRACE_FN)
    lib().zkgpu_set_trace(cb)
    return cb


# ---- Poseidon2 / note-tree Merkle path (witness-side hashing on the GPU) -----------------------------------
POSEIDON_RATE = 7       # shielder_circuits::consts::POSEIDON_RATE
ARITY, NOTE_TREE_HEIGHT = 7, 13   # /root/reference/crates/shielder-setup/lib.rs:4-5, contracts/MerkleTree.sol:17-18


def poseidon_rate():
    """shielder_bindings::hash::poseidon_rate (/root/reference/crates/shielder_bindings/src/hash.rs:10-14)"""
    return POSEIDON_RATE


def poseidon2_hash(inputs):
    """`hash_variable_length` over a batch: inputs (m, len, 4) Montgomery limbs with 1 <= len <= 7 -> (m, 4)."""
    inputs = _u64(inputs)
    if inputs.ndim != 3 or inputs.shape[2] != 4:
        raise ValueError("inputs must have shape (m, len, 4)")
    m, ln = inputs.shape[0], inputs.shape[1]
    out = np.empty((m, 4), dtype=np.uint64)
    _chk(lib().zkgpu_poseidon2_hash_batch(_p(inputs), C.c_size_t(ln), C.c_size_t(m), _p(out)))
    return out


def poseidon_hash(inputs):
    """shielder_bindings::hash::poseidon_hash (hash.rs:16-27): concatenated canonical little-endian 32-byte field
    elements in, one such word out.  A length that is not a multiple of 32 panics upstream -> ValueError."""
    from . import conversions as cv
    inputs = bytes(inputs)
    if len(inputs) % cv.FR_SIZE != 0:
        raise ValueError("Input length must be divisible by F::size()")
    elems = np.stack([cv.vec_to_f(inputs[o:o + cv.FR_SIZE]) for o in range(0, len(inputs), cv.FR_SIZE)]) if inputs else np.zeros((0, 4), np.uint64)
    return cv.field_to_bytes(poseidon2_hash(elems[None])[0])


def merkle_root(paths):
    """paths (m, height, 7, 4) as `vec_to_path` decodes them -> (roots (m, 4), consistent (m,) bool): the root every
    path hashes to and whether each level contains the hash of the level below (MerkleTree.sol:121-152)."""
    paths = _u64(paths)
    if paths.ndim != 4 or paths.shape[2] != ARITY or paths.shape[3] != 4:
        raise ValueError("paths must have shape (m, height, 7, 4)")
    m, height = paths.shape[0], paths.shape[1]
    roots = np.empty((m, 4), dtype=np.uint64)
    ok = np.empty(m, dtype=np.uint8)
    _chk(lib().zkgpu_merkle_root_batch(_p(paths), C.c_size_t(height), C.c_size_t(m), _p(roots), _p(ok)))
    return roots, ok.astype(bool)
