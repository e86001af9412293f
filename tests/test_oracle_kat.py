"""Pins the CPU oracle (oracle/*.hpp) against every known answer the reference holds for the hot
path and against an independent pure-Python big-int model (tests/pyref.py).

Fixtures: tests/golden/ppot_0080_11_raw.bin is a byte copy of the reference's
resources/ppot_0080_11_raw (halo2 `ParamsKZG` RawBytes at k=11) — its g_lagrange block was written by
halo2's own `g_to_lagrange`, i.e. 2048 full-width MSM known answers and one 2^11 G1-iFFT known answer.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O
import pyref as P

REF_PTAU = os.path.join(O.GOLDEN, "ppot_0080_11.ptau")   # copy of the reference's resources/ppot_0080_11.ptau
rng = np.random.default_rng(1234)


def rand_canon(n, p):
    return [int.from_bytes(rng.bytes(40), "little") % p for _ in range(n)]


EDGE_R = [0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (1 << 253), (1 << 64) - 1, (1 << 128) - 1, (1 << 192) + 12345]
EDGE_Q = [0, 1, 2, P.Q_MOD - 1, P.Q_MOD - 2, (1 << 253), (1 << 64) - 1, (1 << 128) - 1, (1 << 192) + 12345]


@pytest.mark.parametrize("field,p,edge", [(0, P.R_MOD, EDGE_R), (1, P.Q_MOD, EDGE_Q)])
def test_field_ops_vs_bigint(field, p, edge):
    a = edge + rand_canon(200, p)
    b = list(reversed(edge)) + rand_canon(200, p)
    am = O.to_mont(field, P.int_to_limbs(a))
    bm = O.to_mont(field, P.int_to_limbs(b))
    # Montgomery form is value * 2^256 mod p (the Rust memory layout)
    assert P.limbs_to_int(am) == [P.to_mont(x, p) for x in a]
    assert P.limbs_to_int(O.from_mont(field, am)) == a
    assert P.limbs_to_int(O.from_mont(field, O.field_op(field, 0, am, bm))) == [x * y % p for x, y in zip(a, b)]
    assert P.limbs_to_int(O.from_mont(field, O.field_op(field, 1, am, bm))) == [(x + y) % p for x, y in zip(a, b)]
    assert P.limbs_to_int(O.from_mont(field, O.field_op(field, 2, am, bm))) == [(x - y) % p for x, y in zip(a, b)]
    assert P.limbs_to_int(O.from_mont(field, O.field_op(field, 4, am))) == [x * x % p for x in a]
    assert P.limbs_to_int(O.from_mont(field, O.field_op(field, 5, am))) == [(-x) % p for x in a]
    inv = P.limbs_to_int(O.from_mont(field, O.field_op(field, 3, am)))
    assert inv == [pow(x, -1, p) if x else 0 for x in a]


def test_fr_from_u512():
    wide = rng.integers(0, 2**64, size=(50, 8), dtype=np.uint64)
    wide[0, :] = 0xFFFFFFFFFFFFFFFF
    got = P.limbs_to_int(O.from_mont(0, O.fr_from_u512(wide)))
    exp = [sum(int(w[i]) << (64 * i) for i in range(8)) % P.R_MOD for w in wide]
    assert got == exp


def test_fr_constants():
    """ROOT_OF_UNITY = 7^((r-1)/2^28); DELTA = 7^(2^28) equals the constant the verifier
    template hard-codes (Halo2Verifier.sol:475); ZETA is a primitive cube root of unity."""
    root = P.limbs_to_int(O.from_mont(0, O.fr_const(0)))[0]
    delta = P.limbs_to_int(O.from_mont(0, O.fr_const(1)))[0]
    zeta = P.limbs_to_int(O.from_mont(0, O.fr_const(2)))[0]
    assert root == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
    assert pow(root, 1 << 28, P.R_MOD) == 1 and pow(root, 1 << 27, P.R_MOD) != 1
    assert delta == 4131629893567559867359510883348571134090853742863529169391034518566172092834
    assert zeta == P.ZETA and pow(zeta, 3, P.R_MOD) == 1 and zeta != 1
    assert P.limbs_to_int(O.fr_const(4))[0] == (1 << 256) % P.R_MOD
    # moduli as printed in Halo2Verifier.sol:222-223
    assert P.Q_MOD == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    assert P.R_MOD == 21888242871839275222246405745257275088548364400416034343698204186575808495617


def test_keccak_known_answers():
    """crates/shielder-account/src/secrets.rs:75-106"""
    m1 = (15).to_bytes(32, "big") + b"nullifier" + (0xFF).to_bytes(4, "big")
    assert O.keccak256(m1).hex() == "375a07a9503d15a291307e33ad0c297c9768fea4712947172ad09f2df34d8015"
    m2 = (16).to_bytes(32, "big") + b"id" + (26).to_bytes(8, "big") + (45).to_bytes(4, "big")
    assert O.keccak256(m2).hex() == "f4b3b097dfb3da737872bdf8b59a3b3723345dc147a0b8229608db69cfef6499"
    assert O.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    # multi-block input (rate 136)
    assert O.keccak256(b"a" * 135) != O.keccak256(b"a" * 136) != O.keccak256(b"a" * 137)
    # length-extension sanity against sha3's keccak permutation: sha3_256 differs only by padding
    assert O.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert hashlib.sha3_256(b"abc").hexdigest() != O.keccak256(b"abc").hex()


def test_rng_known_answers():
    # xoshiro256++ from SplitMix64(0): SplitMix64 reference outputs for seed 0
    s = O.smallrng(0, 4)
    # state after seeding with SplitMix64(0) is (e220a8397b1dcdaf, 6e789e6aa1b965f4, 06c45d188009454f, f88bb8a8724c81ec)
    s0, s3 = 0xE220A8397B1DCDAF, 0xF88BB8A8724C81EC
    rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & (2**64 - 1)
    assert int(s[0]) == (rotl((s0 + s3) & (2**64 - 1), 23) + s0) & (2**64 - 1)
    # ChaCha20, all-zero key, block 0: the classic keystream
    ks = O.chacha20(bytes(32), 8).tobytes().hex()
    assert ks == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                  "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


@pytest.fixture(scope="module")
def raw11():
    return O.srs_read(O.RAW11, 0)


def test_srs_raw_parse(raw11):
    """layout k ‖ g ‖ g_lagrange ‖ g2 ‖ s_g2; all points on curve; g[0] = (1,2)."""
    assert raw11["k"] == 11 and os.path.getsize(O.RAW11) == 4 + 2 * 2048 * 64 + 256
    assert O.g1_on_curve(raw11["g"]) and O.g1_on_curve(raw11["g_lagrange"])
    g0 = P.limbs_to_int(O.from_mont(1, raw11["g"][0]))
    assert g0 == [1, 2]
    assert O.g2_on_curve(raw11["g2"]) and O.g2_on_curve(raw11["s_g2"])
    g2 = P.limbs_to_int(O.from_mont(1, raw11["g2"]))
    assert g2[0] == 0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED  # BN254 G2 generator x.c0


def test_raw_equals_perpetual(raw11):
    """crates/powers-of-tau/lib.rs:267-281"""
    pt = O.srs_read(REF_PTAU, 1)
    assert pt["k"] == raw11["k"]
    assert np.array_equal(pt["g"], raw11["g"])
    assert np.array_equal(pt["g2"], raw11["g2"]) and np.array_equal(pt["s_g2"], raw11["s_g2"])


def test_g1_ops_vs_bigint(raw11):
    g = raw11["g"]
    pts = [tuple(P.limbs_to_int(O.from_mont(1, g[i]))) for i in range(4)]
    s = P.limbs_to_int(O.from_mont(1, O.g1_op(0, g[1], g[2])))
    assert tuple(s) == P.ec_add(pts[1], pts[2])
    d = P.limbs_to_int(O.from_mont(1, O.g1_op(1, g[3])))
    assert tuple(d) == P.ec_add(pts[3], pts[3])
    k = rand_canon(1, P.R_MOD)[0]
    m = P.limbs_to_int(O.from_mont(1, O.g1_op(2, g[1], O.to_mont(0, P.int_to_limbs([k]))[0])))
    assert tuple(m) == P.ec_mul(pts[1], k)
    # P + (-P) = identity = (0,0); P + identity = P
    neg = g[1].copy()
    neg[4:] = O.field_op(1, 5, g[1][4:].reshape(1, 4)).reshape(4)
    assert not O.g1_op(0, g[1], neg).any()
    assert np.array_equal(O.g1_op(0, g[1], np.zeros(8, dtype=np.uint64)), g[1])


def test_msm_small_vs_bigint(raw11):
    for n in (1, 3, 5, 31, 33, 100):
        sc = rand_canon(n, P.R_MOD)
        if n > 3:
            sc[1] = 0
            sc[2] = P.R_MOD - 1
        scm = O.to_mont(0, P.int_to_limbs(sc))
        pts = [tuple(P.limbs_to_int(O.from_mont(1, raw11["g"][i]))) for i in range(n)]
        exp = P.ec_msm(sc, pts)
        for threads in (1, 3):
            got = P.limbs_to_int(O.from_mont(1, O.msm(scm, raw11["g"][:n], threads)))
            assert tuple(got) == exp


def test_msm_known_answers_from_g_lagrange(raw11):
    """g_lagrange[i] = sum_j (omega^{-ij}/n) g[j] — halo2's own output, full-width scalars."""
    n, k = 2048, 11
    w_inv = pow(P.omega_for(k), -1, P.R_MOD)
    n_inv = pow(n, -1, P.R_MOD)
    for i in (0, 1, 2, 777, 2047):
        sc = [pow(w_inv, i * j, P.R_MOD) * n_inv % P.R_MOD for j in range(n)]
        got = O.msm(O.to_mont(0, P.int_to_limbs(sc)), raw11["g"], threads=4)
        assert np.array_equal(got, raw11["g_lagrange"][i]), i


def test_g1_ifft_known_answer(raw11):
    """K6 / a8: g_to_lagrange(g) reproduces the fixture's whole g_lagrange block."""
    got = O.g_to_lagrange(raw11["g"], 11, threads=8)
    assert np.array_equal(got, raw11["g_lagrange"])


def test_commit_lagrange_property(raw11):
    """crates/powers-of-tau/lib.rs:248-264: commit(lagrange_to_coeff(a)) == commit_lagrange(a), a[i] = i."""
    n = 2048
    a = O.to_mont(0, P.int_to_limbs(list(range(n))))
    b = O.domain_op(1, 11, 0, a)
    lhs = O.msm(b, raw11["g"], threads=4)
    rhs = O.msm(a, raw11["g_lagrange"], threads=4)
    assert np.array_equal(lhs, rhs)


def test_pairing_pinned_by_srs(raw11):
    """e(g[1], g2) == e(g[0], s_g2)  (g[1] = tau*G, s_g2 = tau*G2); and it is not trivially true."""
    g = raw11["g"]
    neg_g0 = g[0].copy()
    neg_g0[4:] = O.field_op(1, 5, g[0][4:].reshape(1, 4)).reshape(4)
    assert O.pairing_check(g[1], raw11["g2"], neg_g0, raw11["s_g2"])
    assert not O.pairing_check(g[2], raw11["g2"], neg_g0, raw11["s_g2"])
    assert not O.pairing_check(g[1], raw11["g2"], g[0], raw11["s_g2"])
    # bilinearity with larger powers: e(g[5], g2) == e(g[4], s_g2)
    neg_g4 = g[4].copy()
    neg_g4[4:] = O.field_op(1, 5, g[4][4:].reshape(1, 4)).reshape(4)
    assert O.pairing_check(g[5], raw11["g2"], neg_g4, raw11["s_g2"])


@pytest.mark.parametrize("log_n", [1, 2, 5, 8, 10])
def test_fft_vs_bigint(log_n):
    n = 1 << log_n
    a = rand_canon(n, P.R_MOD)
    w = P.omega_for(log_n)
    wm = O.to_mont(0, P.int_to_limbs([w]))[0]
    got = P.limbs_to_int(O.from_mont(0, O.fft(O.to_mont(0, P.int_to_limbs(a)), wm, log_n)))
    assert got == P.ntt_fast(a, w)
    if log_n <= 5:
        assert got == P.ntt_naive(a, w)
    got_mt = P.limbs_to_int(O.from_mont(0, O.fft(O.to_mont(0, P.int_to_limbs(a)), wm, log_n, threads=4)))
    assert got_mt == got


@pytest.mark.parametrize("j,k", [(4, 6), (5, 7), (7, 8)])
def test_evaluation_domain(j, k):
    n = 1 << k
    ek, om = O.domain(j, k)
    assert (1 << ek) >= n * (j - 1) and ((1 << (ek - 1)) < n * (j - 1) or ek == k)
    omega = P.limbs_to_int(O.from_mont(0, om[0]))[0]
    ext_omega = P.limbs_to_int(O.from_mont(0, om[2]))[0]
    assert omega == P.omega_for(k) and ext_omega == P.omega_for(ek)
    a = rand_canon(n, P.R_MOD)
    am = O.to_mont(0, P.int_to_limbs(a))
    # lagrange_to_coeff then coeff_to_lagrange is the identity
    c = O.domain_op(j, k, 0, am)
    assert np.array_equal(O.domain_op(j, k, 1, c), am)
    # coeff_to_extended == evaluations at zeta * ext_omega^i
    ext = P.limbs_to_int(O.from_mont(0, O.domain_op(j, k, 2, am)))
    for i in (0, 1, 5, (1 << ek) - 1):
        x = P.ZETA * pow(ext_omega, i, P.R_MOD) % P.R_MOD
        assert ext[i] == sum(a[t] * pow(x, t, P.R_MOD) for t in range(n)) % P.R_MOD
    # extended_to_coeff inverts it (poly of degree < n zero-padded to n*(j-1))
    back = P.limbs_to_int(O.from_mont(0, O.domain_op(j, k, 3, O.domain_op(j, k, 2, am))))
    assert back[:n] == a and not any(back[n:])
    # divide_by_vanishing_poly multiplies by 1/((zeta w^i)^n - 1)
    ones = O.to_mont(0, P.int_to_limbs([1] * (1 << ek)))
    tinv = P.limbs_to_int(O.from_mont(0, O.domain_op(j, k, 4, ones)))
    for i in (0, 1, 2, 3, (1 << ek) - 1):
        x = P.ZETA * pow(ext_omega, i, P.R_MOD) % P.R_MOD
        assert tinv[i] * (pow(x, n, P.R_MOD) - 1) % P.R_MOD == 1


def test_eval_polynomial():
    a = rand_canon(33, P.R_MOD)
    x = rand_canon(1, P.R_MOD)[0]
    got = P.limbs_to_int(O.from_mont(0, O.eval_polynomial(O.to_mont(0, P.int_to_limbs(a)), O.to_mont(0, P.int_to_limbs([x]))[0])))
    assert got[0] == sum(c * pow(x, i, P.R_MOD) for i, c in enumerate(a)) % P.R_MOD


def test_summation_by_parts_identity_of_difference_mode(raw11):
    """What csrc/msm.cu's difference mode relies on, stated with the oracle's own group operations on the fixture's g_lagrange:
    sum_i s_i G_i == sum_i (s_i - s_{i+1}) P_i with P_i = G_0 + ... + G_i and s_n = 0 -- for a random column, a piecewise-constant
    one (whose differences are almost all zero) and a prefix of the bases."""
    n = 96
    G = raw11["g_lagrange"][:n]
    Pfx = np.empty_like(G)
    acc = np.zeros(8, dtype=np.uint64)
    for i in range(n):
        acc = O.g1_op(0, acc, G[i])
        Pfx[i] = acc
    r = P.R_MOD
    rand = rand_canon(n, r)
    steps = [rand[0]] * 40 + [rand[1]] * 30 + [0] * 6 + [r - 1] * 20
    for col in (rand, steps, [7] * n, rand[:50]):
        m = len(col)
        diffs = [(col[i] - (col[i + 1] if i + 1 < m else 0)) % r for i in range(m)]
        lhs = O.msm(O.to_mont(0, P.int_to_limbs(col)), G[:m])
        rhs = O.msm(O.to_mont(0, P.int_to_limbs(diffs)), Pfx[:m])
        assert np.array_equal(lhs, rhs)
    assert sum(1 for i in range(n - 1) if steps[i] != steps[i + 1]) == 3
