"""ctypes loader for oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (the CPU checker)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
U64P = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _LIB = C.CDLL(path)
        _LIB.orc_last_error.restype = C.c_char_p
    return _LIB


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def field_op(field, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
    _chk(lib().orc_field_op(field, op, _p(a), _p(b), _p(out), C.c_size_t(a.size // 4)))
    return out


def to_mont(field, canon):
    canon = np.ascontiguousarray(canon, dtype=np.uint64)
    out = np.empty_like(canon)
    _chk(lib().orc_to_mont(field, _p(canon), _p(out), C.c_size_t(canon.size // 4)))
    return out


def from_mont(field, mont):
    mont = np.ascontiguousarray(mont, dtype=np.uint64)
    out = np.empty_like(mont)
    _chk(lib().orc_from_mont(field, _p(mont), _p(out), C.c_size_t(mont.size // 4)))
    return out


def fr_from_u512(wide):
    wide = np.ascontiguousarray(wide, dtype=np.uint64)
    n = wide.size // 8
    out = np.empty((n, 4), dtype=np.uint64)
    _chk(lib().orc_fr_from_u512(_p(wide), _p(out), C.c_size_t(n)))
    return out


def fr_const(which):
    out = np.empty(4, dtype=np.uint64)
    _chk(lib().orc_fr_const(which, _p(out)))
    return out


def g1_op(op, a, b=None):
    out = np.empty(8, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
    _chk(lib().orc_g1_op(op, _p(a), _p(b), _p(out)))
    return out


def g1_on_curve(pts):
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    return bool(lib().orc_g1_on_curve(_p(pts), C.c_size_t(pts.size // 8)))


def msm(scalars, bases, threads=1):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    n = scalars.size // 4
    assert bases.size // 8 == n
    out = np.empty(8, dtype=np.uint64)
    _chk(lib().orc_msm(_p(scalars), _p(bases), C.c_size_t(n), threads, _p(out)))
    return out


def fft(a, omega, log_n, threads=1):
    a = np.array(a, dtype=np.uint64, copy=True)
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    assert a.size == 4 << log_n
    _chk(lib().orc_fft(_p(a), _p(omega), log_n, threads))
    return a


def domain(j, k):
    ek = C.c_uint(0)
    om = np.empty((4, 4), dtype=np.uint64)
    _chk(lib().orc_domain(j, k, C.byref(ek), _p(om)))
    return ek.value, om


def domain_op(j, k, which, a, threads=1):
    ek, _ = domain(j, k)
    n = 1 << k
    out_len = {0: n, 1: n, 2: 1 << ek, 3: n * (j - 1), 4: 1 << ek}[which]
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty((out_len, 4), dtype=np.uint64)
    _chk(lib().orc_domain_op(j, k, which, _p(a), _p(out), threads))
    return out


def g_to_lagrange(g, k, threads=1):
    g = np.ascontiguousarray(g, dtype=np.uint64)
    out = np.empty((1 << k, 8), dtype=np.uint64)
    _chk(lib().orc_g_to_lagrange(_p(g), k, _p(out), threads))
    return out


def eval_polynomial(poly, x):
    poly = np.ascontiguousarray(poly, dtype=np.uint64)
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    _chk(lib().orc_eval_polynomial(_p(poly), C.c_size_t(poly.size // 4), _p(x), _p(out)))
    return out


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_keccak256(data, C.c_size_t(len(data)), out)
    return out.raw


def smallrng(seed, n):
    out = np.empty(n, dtype=np.uint64)
    lib().orc_smallrng(C.c_uint64(seed), _p(out), C.c_size_t(n))
    return out


def chacha20(seed32: bytes, n):
    out = np.empty(n, dtype=np.uint64)
    lib().orc_chacha20(seed32, _p(out), C.c_size_t(n))
    return out


def random_fr(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_random_fr(C.c_uint64(seed), _p(out), C.c_size_t(n))
    return out


def random_fr_rng(state, n):
    """n x Fr::random from a running SmallRng; `state` (4,) uint64 is advanced in place"""
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_random_fr_rng(_p(state), _p(out), C.c_size_t(n))
    return out


def smallrng_state(seed):
    """the xoshiro256++ state of SmallRng::seed_from_u64(seed) (four SplitMix64 outputs)"""
    st, z, mask = [], seed, (1 << 64) - 1
    for _ in range(4):
        z = (z + 0x9E3779B97F4A7C15) & mask
        x = z
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & mask
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & mask
        st.append(x ^ (x >> 31))
    return np.array(st, dtype=np.uint64)


def srs_read(path, fmt):
    """fmt 0 = halo2 RawBytes, 1 = .ptau.  Returns dict(k, g, g_lagrange|None, g2, s_g2)."""
    k = C.c_uint(0)
    _chk(lib().orc_srs_read(path.encode(), fmt, C.byref(k), None, None, None))
    n = 1 << k.value
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64) if fmt == 0 else None
    g2s = np.empty((2, 16), dtype=np.uint64)
    _chk(lib().orc_srs_read(path.encode(), fmt, C.byref(k), _p(g), _p(gl), _p(g2s)))
    return dict(k=k.value, g=g, g_lagrange=gl, g2=g2s[0].copy(), s_g2=g2s[1].copy())


def pairing_check(p1, q1, p2, q2):
    ok = C.c_int(0)
    arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (p1, q1, p2, q2)]
    _chk(lib().orc_pairing_check(*[_p(x) for x in arrs], C.byref(ok)))
    return bool(ok.value)


def g2_on_curve(q):
    q = np.ascontiguousarray(q, dtype=np.uint64)
    return bool(lib().orc_g2_on_curve(_p(q)))


TRACE_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p, C.c_size_t)


def params_setup(k, seed, threads=8):
    """ParamsKZG::setup(k, SmallRng::seed_from_u64(seed)) as the reference's seeded tests do
    (crates/halo2-verifier/src/generator.rs:118-119).  G2 generator from the ppot fixture."""
    n = 1 << k
    gen = srs_read(RAW11, 0)["g2"]
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64)
    g2s = np.empty((2, 16), dtype=np.uint64)
    _chk(lib().orc_params_setup(k, C.c_uint64(seed), _p(np.ascontiguousarray(gen)), threads, _p(g), _p(gl), _p(g2s)))
    return dict(k=k, g=g, g_lagrange=gl, g2=g2s[0].copy(), s_g2=g2s[1].copy())


def params_setup_rng(k, state, threads=8):
    """ParamsKZG::setup(k, &mut rng) with a running SmallRng: `state` (4,) uint64 is advanced in place"""
    n = 1 << k
    gen = srs_read(RAW11, 0)["g2"]
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64)
    g2s = np.empty((2, 16), dtype=np.uint64)
    _chk(lib().orc_params_setup_rng(k, _p(state), _p(np.ascontiguousarray(gen)), threads, _p(g), _p(gl), _p(g2s)))
    return dict(k=k, g=g, g_lagrange=gl, g2=g2s[0].copy(), s_g2=g2s[1].copy())


GOLDEN = os.path.join(ROOT, "tests", "golden")
RAW11 = os.path.join(GOLDEN, "ppot_0080_11_raw.bin")


# ---- PLONK prover / verifier (oracle/plonk.hpp) ------------------------------------------------
class OracleBackend:
    """Field backend for zkgpu.circuits backed by the CPU oracle (tests only)."""

    @staticmethod
    def random(seed, count):
        return random_fr(seed, count)

    @staticmethod
    def const(v):
        import pyref
        return to_mont(0, pyref.int_to_limbs([v % pyref.R_MOD]))[0]

    @staticmethod
    def to_mont(canon):
        return to_mont(0, canon).reshape(-1, 4)

    @staticmethod
    def mul(a, b):
        return field_op(0, 0, a, b).reshape(-1, 4)

    @staticmethod
    def add(a, b):
        return field_op(0, 1, a, b).reshape(-1, 4)

    @staticmethod
    def sub(a, b):
        return field_op(0, 2, a, b).reshape(-1, 4)


def downsized_srs(k, raw=None):
    """ParamsKZG::downsize(k) of the k=11 fixture: truncate g, recompute g_lagrange (oracle)."""
    raw = raw or srs_read(RAW11, 0)
    assert k <= raw["k"]
    if k == raw["k"]:
        return dict(raw)
    g = raw["g"][: 1 << k].copy()
    return dict(k=k, g=g, g_lagrange=g_to_lagrange(g, k, threads=8), g2=raw["g2"], s_g2=raw["s_g2"])


class PlonkOracle:
    def __init__(self, blob: bytes, srs, threads=1):
        self.h = C.c_void_p()
        g2s = np.concatenate([srs["g2"], srs["s_g2"]]).astype(np.uint64)
        g, gl = np.ascontiguousarray(srs["g"]), np.ascontiguousarray(srs["g_lagrange"])
        _chk(lib().orc_plonk_new(blob, C.c_size_t(len(blob)), srs["k"], _p(g), _p(gl), _p(g2s), threads, C.byref(self.h)))
        info = np.zeros(16, dtype=np.uint64)
        _chk(lib().orc_plonk_info(self.h, _p(info)))
        (self.k, self.n, self.num_advice, self.num_fixed, self.degree, self.blinding_factors, self.num_perm_sets,
         self.num_quotients, self.num_evals, self.proof_len, self.extended_k) = [int(x) for x in info[:11]]

    def vk(self, num_perm_cols):
        fc = np.zeros((self.num_fixed, 8), dtype=np.uint64)
        pc = np.zeros((num_perm_cols, 8), dtype=np.uint64)
        dg = np.zeros(4, dtype=np.uint64)
        _chk(lib().orc_plonk_vk(self.h, _p(fc), _p(pc), _p(dg)))
        return fc, pc, dg

    def write_pk(self, num_selectors=0):
        """`marshall_pk`: k u32 LE ‖ ProvingKey::to_bytes(RawBytesUnchecked), as the reference's build.rs writes pk.bin"""
        ln = C.c_size_t(0)
        _chk(lib().orc_plonk_write_pk(self.h, num_selectors, None, C.c_size_t(0), C.byref(ln)))
        buf = C.create_string_buffer(ln.value)
        _chk(lib().orc_plonk_write_pk(self.h, num_selectors, buf, ln, C.byref(ln)))
        return buf.raw

    def check_witness(self, advice, instance):
        ok = C.c_int(0)
        advice, instance = np.ascontiguousarray(advice), np.ascontiguousarray(instance)
        _chk(lib().orc_plonk_check_witness(self.h, _p(advice), _p(instance), C.c_size_t(instance.size // 4), C.byref(ok)))
        return bool(ok.value), lib().orc_last_error().decode()

    def prove(self, advice, instance, seed):
        advice, instance = np.ascontiguousarray(advice), np.ascontiguousarray(instance)
        proof = C.create_string_buffer(self.proof_len)
        stats = np.zeros(3, dtype=np.uint64)
        _chk(lib().orc_plonk_prove(self.h, _p(advice), _p(instance), C.c_size_t(instance.size // 4), C.c_uint64(seed), proof, _p(stats)))
        self.last_stats = dict(msm=int(stats[0]), ntt=int(stats[1]), ext_ntt=int(stats[2]))
        return proof.raw

    def prove_rng(self, advice, instance, mode, rng_data):
        """create_proof with the caller's rng: mode 0 u64 seed (np.uint64 array of 1), 1 running SmallRng state ((4,) uint64,
        advanced in place), 2 ChaCha20 seed (32 bytes as a uint8 array)"""
        advice, instance = np.ascontiguousarray(advice), np.ascontiguousarray(instance)
        proof = C.create_string_buffer(self.proof_len)
        _chk(lib().orc_plonk_prove_rng(self.h, _p(advice), _p(instance), C.c_size_t(instance.size // 4), int(mode), _p(rng_data), proof))
        return proof.raw

    def verify(self, proof: bytes, instance):
        ok = C.c_int(0)
        instance = np.ascontiguousarray(instance)
        _chk(lib().orc_plonk_verify(self.h, proof, C.c_size_t(len(proof)), _p(instance), C.c_size_t(instance.size // 4), C.byref(ok)))
        return bool(ok.value)

    def verify_batch(self, proofs, instances, threads=8):
        """all proofs of a batch at once (RLC of the pairing inputs): returns (all accepted, number malformed)"""
        blob = b"".join(proofs) if not isinstance(proofs, (bytes, bytearray)) else bytes(proofs)
        instances = np.ascontiguousarray(instances, dtype=np.uint64)
        m = len(blob) // self.proof_len
        ok, bad = C.c_int(0), C.c_uint64(0)
        _chk(lib().orc_plonk_verify_batch(self.h, blob, C.c_size_t(self.proof_len), C.c_size_t(m), _p(instances),
                                          C.c_size_t(instances.size // (4 * m)), threads, C.byref(ok), C.byref(bad)))
        return bool(ok.value), int(bad.value)

    def __del__(self):
        try:
            lib().orc_plonk_free(self.h)
        except Exception:
            pass


def poseidon2_hash(inputs):
    """inputs [m][len][4] Mont-LE (len 1..7) -> [m][4]"""
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    m, ln = inputs.shape[0], inputs.shape[1]
    out = np.empty((m, 4), dtype=np.uint64)
    _chk(lib().orc_poseidon2_hash(_p(inputs), C.c_size_t(ln), C.c_size_t(m), _p(out)))
    return out


def merkle_root(paths):
    """paths [m][height][7][4] Mont-LE -> (roots [m][4], consistent [m] bool)"""
    paths = np.ascontiguousarray(paths, dtype=np.uint64)
    m, height = paths.shape[0], paths.shape[1]
    roots = np.empty((m, 4), dtype=np.uint64)
    ok = np.empty(m, dtype=np.uint8)
    _chk(lib().orc_merkle_root(_p(paths), C.c_size_t(height), C.c_size_t(m), _p(roots), _p(ok)))
    return roots, ok.astype(bool)


def set_rayon_threads(t):
    """rayon thread count emulated by the oracle prover's vanishing argument (default 1)"""
    lib().orc_plonk_set_rayon_threads(C.c_uint(int(t)))
