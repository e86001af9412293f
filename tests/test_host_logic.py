"""The product's host-side prover logic, built with g++ and run on the CPU (no GPU, no nvcc):
Keccak-256 / EVM transcript / SmallRng / codecs (csrc/host_util.hpp) and the constraint-system parser,
derived sizes and permutation assembly (csrc/plonk_types.hpp).  Pinned by the reference's own known answers
(/root/reference/crates/shielder-account/src/secrets.rs:75-106) and cross-checked with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pyref as P
from zkgpu import circuits

ROOT = O.ROOT
CSRC = os.path.join(ROOT, "zkos-monorepo_b200", "csrc")


@pytest.fixture(scope="module")
def hl(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hl") / "libhostlogic.so")
    cuda_inc = "/usr/local/cuda/include"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", cuda_inc, "-I", CSRC,
                           os.path.join(ROOT, "tests", "emu", "host_logic.cpp"), "-o", so])
    return C.CDLL(so)


def _keccak(hl, data):
    out = C.create_string_buffer(32)
    hl.hl_keccak256(data, C.c_size_t(len(data)), out)
    return out.raw


def test_keccak_known_answers(hl):
    m1 = (15).to_bytes(32, "big") + b"nullifier" + (0xFF).to_bytes(4, "big")
    assert _keccak(hl, m1).hex() == "375a07a9503d15a291307e33ad0c297c9768fea4712947172ad09f2df34d8015"
    m2 = (16).to_bytes(32, "big") + b"id" + (26).to_bytes(8, "big") + (45).to_bytes(4, "big")
    assert _keccak(hl, m2).hex() == "f4b3b097dfb3da737872bdf8b59a3b3723345dc147a0b8229608db69cfef6499"
    assert _keccak(hl, b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    rng = np.random.default_rng(1)
    for n in (1, 31, 32, 135, 136, 137, 271, 272, 273, 5000):
        d = rng.bytes(n)
        assert _keccak(hl, d) == O.keccak256(d), n


def test_smallrng_matches_oracle(hl):
    for seed in (0, 42, 2**64 - 1):
        out = np.empty(64, dtype=np.uint64)
        hl.hl_smallrng(C.c_uint64(seed), out.ctypes.data_as(C.c_void_p), C.c_size_t(64))
        assert np.array_equal(out, O.smallrng(seed, 64))


def test_challenge_reduction(hl):
    """hash (any 256-bit value) -> Fr: values >= r, the largest word, and random ones"""
    rng = np.random.default_rng(3)
    vals = [0, 1, P.R_MOD - 1, P.R_MOD, P.R_MOD + 1, 5 * P.R_MOD + 7, (1 << 256) - 1] + [int.from_bytes(rng.bytes(32), "big") for _ in range(200)]
    for v in vals:
        out = np.empty(8, dtype=np.uint32)
        hl.hl_fr_from_be_reduce(v.to_bytes(32, "big"), out.ctypes.data_as(C.c_void_p))
        got = P.limbs_to_int(out.view(np.uint64))[0]
        assert got == (v % P.R_MOD) * P.MONT_R % P.R_MOD, hex(v)


def test_transcript_matches_spec(hl):
    """Halo2Verifier.sol:101-124: challenge = keccak(state ‖ absorbed) mod r; a squeeze right after a squeeze
    hashes prev_hash ‖ 0x01.  Points / scalars are written as canonical big-endian words."""
    raw = O.srs_read(O.RAW11, 0)
    digest = O.random_fr(9, 1)[0]
    scalars = O.random_fr(10, 5)
    points = raw["g"][3:6]
    n_ch = 4
    ch = np.empty((n_ch, 4), dtype=np.uint64)
    proof = C.create_string_buffer(32 * 5 + 64 * 3)
    hl.hl_transcript(digest.ctypes.data_as(C.c_void_p), scalars.ctypes.data_as(C.c_void_p), C.c_size_t(5),
                     np.ascontiguousarray(points).ctypes.data_as(C.c_void_p), C.c_size_t(3), ch.ctypes.data_as(C.c_void_p), C.c_size_t(n_ch), proof)
    be = lambda field, m: b"".join(int(v).to_bytes(32, "big") for v in P.limbs_to_int(O.from_mont(field, m)))
    body = be(0, scalars) + be(1, points.reshape(-1, 4))
    assert proof.raw == body
    buf = be(0, digest[None]) + body
    want = []
    h = O.keccak256(buf); want.append(int.from_bytes(h, "big") % P.R_MOD)
    for _ in range(n_ch - 1):
        h = O.keccak256(h + b"\x01"); want.append(int.from_bytes(h, "big") % P.R_MOD)
    assert P.limbs_to_int(O.from_mont(0, ch)) == want


@pytest.mark.parametrize("name", sorted(circuits.SHAPES))
def test_constraint_system_sizes(hl, name):
    shape = circuits.Shape(name) if circuits.SHAPES[name]["k"] <= 9 else circuits.Shape(name, k=9)
    circ = circuits.Circuit(shape, O.OracleBackend, seed=1)
    info = np.zeros(10, dtype=np.uint64)
    err = C.create_string_buffer(256)
    assert hl.hl_cs_info(circ.blob, C.c_size_t(len(circ.blob)), info.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == 0, err.value
    want = [shape.k, shape.n, shape.degree, shape.blinding_factors, shape.chunk_len, shape.num_perm_sets, shape.num_quotients,
            shape.extended_k, shape.num_evals, shape.proof_len]
    assert [int(x) for x in info] == want
    # proof length formula of the verifier generator (codegen/util.rs:175-186)
    assert shape.proof_len == 64 * (shape.num_advice + 3 * shape.n_lookup + shape.num_perm_sets + 1 + shape.num_quotients) + 32 * shape.num_evals + 128


def test_malformed_blobs_rejected(hl):
    shape = circuits.Shape("tiny")
    blob = circuits.Circuit(shape, O.OracleBackend, seed=1).blob
    info = np.zeros(10, dtype=np.uint64)
    err = C.create_string_buffer(256)
    call = lambda b: hl.hl_cs_info(bytes(b), C.c_size_t(len(b)), info.ctypes.data_as(C.c_void_p), err, C.c_size_t(256))
    assert call(blob) == 0
    assert call(blob[:-1]) == -1 and call(blob + b"\0") == -1 and call(b"") == -1
    bad = bytearray(blob); bad[0] ^= 1
    assert call(bad) == -1 and b"magic" in err.value
    bad = bytearray(blob); bad[4:8] = (40).to_bytes(4, "little")     # k out of range
    assert call(bad) == -1


def test_permutation_assembly(hl):
    shape = circuits.Shape("tiny")
    circ = circuits.Circuit(shape, O.OracleBackend, seed=1)
    S, n = len(shape.perm_columns), shape.n
    mc = np.zeros(S * n, dtype=np.uint32); mr = np.zeros(S * n, dtype=np.uint32)
    assert hl.hl_perm_mapping(circ.blob, C.c_size_t(len(circ.blob)), mc.ctypes.data_as(C.c_void_p), mr.ctypes.data_as(C.c_void_p)) == 0
    nxt = mc.astype(np.int64) * n + mr
    assert sorted(nxt.tolist()) == list(range(S * n))                 # sigma is a permutation of the cells
    # every copy constraint joins its two cells into one cycle
    cyc = -np.ones(S * n, dtype=np.int64)
    for start in range(S * n):
        if cyc[start] >= 0:
            continue
        i = start
        while cyc[i] < 0:
            cyc[i] = start
            i = nxt[i]
    for lc, lr, rc, rr in circ.copies:
        assert cyc[lc * n + lr] == cyc[rc * n + rr]
    # cells without copy constraints map to themselves
    touched = set()
    for lc, lr, rc, rr in circ.copies:
        touched.add(lc * n + lr); touched.add(rc * n + rr)
    free = np.array([i for i in range(S * n) if i not in touched])
    assert np.array_equal(nxt[free], free)


@pytest.mark.parametrize("field,p", [(0, P.R_MOD), (1, P.Q_MOD)])
def test_binary_inversion(hl, field, p):
    """fe_inv (binary extended Euclid) == a^(p-2) == big-int inverse, in Montgomery form; 0 -> 0."""
    rng = np.random.default_rng(5 + field)
    vals = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, (p - 1) // 2, 1 << 253, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 15, (1 << 15) - 1, 1 << 31,
            (1 << 254) % p, pow(2, -1, p), pow(2, -15, p), pow(2, -510, p), pow(3, -1, p)] \
        + [pow(2, e, p) for e in range(0, 254, 7)] + [(p - pow(2, e, p)) % p for e in range(0, 254, 11)] \
        + [int.from_bytes(rng.bytes(40), "little") % p for _ in range(3000)] + [int.from_bytes(rng.bytes(k), "little") for k in range(1, 31)]
    mont = P.int_to_limbs([v * P.MONT_R % p for v in vals])
    a32 = mont.view(np.uint32)
    got = np.empty_like(a32); ref = np.empty_like(a32)
    hl.hl_inv(field, 0, a32.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)))
    hl.hl_inv(field, 1, a32.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)))
    assert np.array_equal(got, ref)
    euc = np.empty_like(a32)
    hl.hl_inv(field, 2, a32.ctypes.data_as(C.c_void_p), euc.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)))
    assert np.array_equal(euc, ref)
    want = [(pow(v, -1, p) * P.MONT_R % p) if v else 0 for v in vals]
    assert P.limbs_to_int(got.view(np.uint64)) == want


def test_quotient_from_cosets_identity():
    """The prover evaluates the quotient on Q cosets g_c * H of the size-n subgroup instead of halo2's 2^ek-point extended domain
    (DESIGN.md "Quotient cosets").  Plain-integer check of the reconstruction k_coset_combine implements: for a random h of degree
    < Q n, the size-n interpolants of its values on the cosets give back the pieces h_j through
    h_j[m] = sum_c Vinv[j][c] * g_c^-m * hhat_c[m],  V[c][j] = (g_c^n)^j."""
    import random
    p = P.R_MOD
    k, ek, Q = 3, 6, 5
    n = 1 << k
    rng = random.Random(3)
    root = pow(7, (p - 1) >> 28, p)                     # ROOT_OF_UNITY
    w_ext = pow(root, 1 << (28 - ek), p)
    w = pow(w_ext, 1 << (ek - k), p)
    zeta = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
    assert pow(zeta, 3, p) == 1 and zeta != 1 and pow(w, n, p) == 1
    h = [rng.randrange(p) for _ in range(Q * n)]
    g = [zeta * pow(w_ext, c, p) % p for c in range(Q)]
    ev = lambda x: sum(co * pow(x, i, p) for i, co in enumerate(h)) % p
    n_inv = pow(n, -1, p)
    hhat = []
    for c in range(Q):
        vals = [ev(g[c] * pow(w, t, p) % p) for t in range(n)]
        hhat.append([n_inv * sum(vals[t] * pow(w, -t * m, p) for t in range(n)) % p for m in range(n)])   # inverse NTT of size n
    # invert the Q x Q Vandermonde matrix in G_c = g_c^n (Gauss-Jordan mod p)
    G = [pow(x, n, p) for x in g]
    M = [[pow(G[c], j, p) for j in range(Q)] + [int(c == j) for j in range(Q)] for c in range(Q)]
    for col in range(Q):
        piv = next(r for r in range(col, Q) if M[r][col])
        M[col], M[piv] = M[piv], M[col]
        inv = pow(M[col][col], -1, p)
        M[col] = [v * inv % p for v in M[col]]
        for r in range(Q):
            if r != col and M[r][col]:
                f = M[r][col]
                M[r] = [(a - f * b) % p for a, b in zip(M[r], M[col])]
    vinv = [row[Q:] for row in M]
    for j in range(Q):
        for m in range(n):
            got = sum(vinv[j][c] * pow(g[c], -m, p) * hhat[c][m] for c in range(Q)) % p
            assert got == h[j * n + m]


def test_proof_rng_modes(hl):
    """csrc/host_util.hpp ProofRng: u64 seed = SmallRng::seed_from_u64, running state continues and is handed back,
    32-byte seed = ChaCha20Rng::from_seed (rand_chacha 0.3.1) — each against the oracle's generators"""
    out = np.empty(100, dtype=np.uint64)
    seed = np.array([42], dtype=np.uint64)
    hl.hl_proof_rng(0, seed.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(100))
    assert np.array_equal(out, O.smallrng(42, 100)) and seed[0] == 42
    state = O.smallrng_state(42)
    hl.hl_proof_rng(1, state.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(60))
    assert np.array_equal(out[:60], O.smallrng(42, 100)[:60])
    hl.hl_proof_rng(1, state.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(40))
    assert np.array_equal(out[:40], O.smallrng(42, 100)[60:]), "the state written back does not continue the stream"
    key = np.arange(32, dtype=np.uint8)
    hl.hl_proof_rng(2, key.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(100))
    assert np.array_equal(out, O.chacha20(key.tobytes(), 100)) and np.array_equal(key, np.arange(32, dtype=np.uint8))


def test_pk_bin_reader_against_the_oracle_writer(hl):
    """`pk.bin` = k ‖ ProvingKey::to_bytes(RawBytesUnchecked) (build.rs:19-33): the product's reader (csrc/pk_file.hpp) locates every
    section of a file the oracle wrote, with and without selector bit-vectors, and rejects truncated / padded files."""
    shape = circuits.Shape("tiny_lookup")
    circ = circuits.Circuit(shape, O.OracleBackend, seed=2)
    po = O.PlonkOracle(circ.blob, O.downsized_srs(shape.k), threads=2)
    digest = po.vk(len(shape.perm_columns))[2]
    for nsel in (0, 3):
        pk_bin = po.write_pk(nsel)
        cs_blob = circ.cs_blob(digest, num_selectors=nsel)
        info = np.zeros(7, dtype=np.uint64)
        first = np.zeros(56, dtype=np.uint32)
        err = C.create_string_buffer(256)
        rc = hl.hl_pk_file(cs_blob, C.c_size_t(len(cs_blob)), pk_bin, C.c_size_t(len(pk_bin)), info.ctypes.data_as(C.c_void_p),
                           first.ctypes.data_as(C.c_void_p), err, C.c_size_t(256))
        assert rc == 0, err.value
        en = 1 << shape.extended_k
        assert info.tolist() == [shape.k, shape.num_fixed, len(shape.perm_columns), shape.num_fixed, en, len(shape.perm_columns), en]
        f = first.view(np.uint64).reshape(7, 4)
        assert np.array_equal(f[0], circ.fixed[0][0]), "fixed_values[0][0]"
        assert np.array_equal(f[6], digest), "transcript_repr"
        # the coefficient / extended forms are the oracle's own transforms of that column
        coeffs = O.domain_op(shape.degree, shape.k, 0, circ.fixed[0])
        assert np.array_equal(f[1], coeffs[0]) and np.array_equal(f[2], O.domain_op(shape.degree, shape.k, 2, coeffs)[0])
        for bad in (pk_bin[:-1], pk_bin + b"\\0", pk_bin[:1000]):
            assert hl.hl_pk_file(cs_blob, C.c_size_t(len(cs_blob)), bad, C.c_size_t(len(bad)), info.ctypes.data_as(C.c_void_p),
                                 first.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == -1
    # a full circuit blob is not accepted where the constraint system alone is expected, and vice versa
    assert hl.hl_pk_file(circ.blob, C.c_size_t(len(circ.blob)), pk_bin, C.c_size_t(len(pk_bin)), info.ctypes.data_as(C.c_void_p),
                         first.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == -1
    i2 = np.zeros(10, dtype=np.uint64)
    assert hl.hl_cs_info(cs_blob, C.c_size_t(len(cs_blob)), i2.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == 0
    assert int(i2[2]) == shape.degree and int(i2[9]) == shape.proof_len


def test_corrupt_counts_are_rejected_before_allocation(hl):
    """a blob whose query / gate / copy counts exceed what the blob can hold is refused (no multi-gigabyte resize)"""
    shape = circuits.Shape("tiny")
    blob = bytearray(circuits.Circuit(shape, O.OracleBackend, seed=1).blob)
    info = np.zeros(10, dtype=np.uint64)
    err = C.create_string_buffer(256)
    blob[20:24] = (0xFFFFFFF0).to_bytes(4, "little")      # number of advice queries
    assert hl.hl_cs_info(bytes(blob), C.c_size_t(len(blob)), info.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == -1
    assert b"count exceeds" in err.value or b"truncated" in err.value


def test_degree_without_permutation_columns(hl):
    """ConstraintSystem::degree() starts from the permutation argument's 3 even when no column is copy-enabled"""
    import struct
    s = circuits.Shape("tiny")
    blob = [struct.pack("<5I", circuits.MAGIC, s.k, 1, 1, 1)]
    blob.append(struct.pack("<IIi", 1, 0, 0)); blob.append(struct.pack("<IIi", 1, 0, 0)); blob.append(struct.pack("<IIi", 1, 0, 0))   # queries
    blob.append(struct.pack("<I", 0))                                                  # constants
    gate = [(circuits.OP_FIXED, 0), (circuits.OP_ADVICE, 0), (circuits.OP_MUL, 0)]
    blob.append(struct.pack("<II", 1, len(gate))); blob.append(np.array(gate, dtype=np.uint32).tobytes())
    blob.append(struct.pack("<I", 0)); blob.append(struct.pack("<I", 0))                # no permutation columns, no lookups
    blob.append(np.zeros((s.n, 4), dtype=np.uint64).tobytes()); blob.append(struct.pack("<I", 0))
    blob = b"".join(blob)
    info = np.zeros(10, dtype=np.uint64)
    err = C.create_string_buffer(256)
    assert hl.hl_cs_info(blob, C.c_size_t(len(blob)), info.ctypes.data_as(C.c_void_p), err, C.c_size_t(256)) == 0, err.value
    assert int(info[2]) == 3 and int(info[6]) == 2 and int(info[4]) == 1          # degree 3, two quotient pieces, chunk_len 1


@pytest.mark.parametrize("c", [9, 12, 13, 14])
def test_msm_digit_scalar_recodes_to_the_same_value(hl, c):
    """csrc/msm_digits.cuh: the signed digits the sort emits, with the sign flip of values within 2^224 below r, sum back to s_i
    (plain) or s_i - s_{i+1} (difference mode) mod r; small negative values and downward steps get a single non-zero digit."""
    r = P.R_MOD
    rng = np.random.default_rng(c)
    W = 254 // c + 1
    edge = [0, 1, 2, r - 1, r - 2, (r - 1) // 2, (r + 1) // 2, (r + 3) // 2, 1 << (c - 1), (1 << (c - 1)) + 1, (1 << c) - 1, 1 << c, r - (1 << (c - 1)),
            (1 << 253) - 1, 1 << 253, 255, r - 255, int("1" * 253, 2)]
    s = edge * len(edge) + [int.from_bytes(rng.bytes(40), "little") % r for _ in range(4000)]
    nxt = [e for e in edge for _ in edge] + [int.from_bytes(rng.bytes(40), "little") % r for _ in range(3000)] + s[len(edge) ** 2 + 3000:]
    assert len(s) == len(nxt)
    sm = P.int_to_limbs([v * P.MONT_R % r for v in s]).view(np.uint32)
    nm = P.int_to_limbs([v * P.MONT_R % r for v in nxt]).view(np.uint32)
    for diff in (0, 1):
        digits = np.zeros((len(s), W), dtype=np.int32)
        flip = np.zeros(len(s), dtype=np.uint8)
        hl.hl_msm_digits(sm.ctypes.data_as(C.c_void_p), nm.ctypes.data_as(C.c_void_p), C.c_size_t(len(s)), diff, c, W,
                         digits.ctypes.data_as(C.c_void_p), flip.ctypes.data_as(C.c_void_p))
        assert np.abs(digits).max() <= 1 << (c - 1)
        for i in range(len(s)):
            want = (s[i] - nxt[i]) % r if diff else s[i]
            got = sum(int(d) << (c * w) for w, d in enumerate(digits[i]))
            assert got % r == want, (diff, i)
            assert bool(flip[i]) == (want != 0 and r - want < 1 << 224)     # values just below r are taken as small negatives
            assert (got < 0) == bool(flip[i]) and -(1 << 224) < got < r
            if want in (1, r - 1, 255, r - 255):
                assert np.count_nonzero(digits[i]) == 1
            if want == 0:
                assert not digits[i].any()
