"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through libzkgpu's C ABI;
the expected values come from the CPU oracle on the same seeded inputs, bit-exact.
Golden fixture: tests/golden/ppot_0080_11_raw.bin (halo2's own g / g_lagrange at k=11)."""
import numpy as np
import pytest

import oracle_lib as O
import pyref as P
import zkgpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raw11():
    return O.srs_read(O.RAW11, 0)


@pytest.fixture(scope="module")
def params11(raw11):
    zkgpu.init(0)
    return zkgpu.ParamsKZG(11, raw11["g"], raw11["g_lagrange"])


def omega(log_n):
    return O.to_mont(0, P.int_to_limbs([P.omega_for(log_n)]))[0]


def edge_scalars(n, seed):
    s = O.random_fr(seed, n)
    if n >= 8:
        s[0] = 0                                                   # zero scalar
        s[1] = O.to_mont(0, P.int_to_limbs([P.R_MOD - 1]))[0]      # r - 1
        s[2] = O.to_mont(0, P.int_to_limbs([1]))[0]
        s[3] = O.to_mont(0, P.int_to_limbs([(1 << 253) + 5]))[0]
        s[4] = O.to_mont(0, P.int_to_limbs([(1 << 64) - 1]))[0]    # 64-bit small
    return s


# ---- NTT ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 10, 11, 12, 13, 14, 15, 16, 18])
def test_ntt_matches_oracle(log_n):
    zkgpu.init(0)
    a = O.random_fr(100 + log_n, 1 << log_n)
    w = omega(log_n)
    assert np.array_equal(zkgpu.best_fft(a, w, log_n).reshape(-1, 4), O.fft(a, w, log_n, threads=8).reshape(-1, 4))


@pytest.mark.parametrize("log_n,batch", [(11, 7), (12, 3), (13, 5), (16, 2)])
def test_ntt_batch_matches_oracle(log_n, batch):
    zkgpu.init(0)
    n = 1 << log_n
    a = O.random_fr(7, n * batch).reshape(batch, n, 4)
    a[0, :5] = 0
    w = omega(log_n)
    got = zkgpu.best_fft(a, w, log_n, batch=batch).reshape(batch, n, 4)
    for b in range(batch):
        assert np.array_equal(got[b], O.fft(a[b], w, log_n, threads=8).reshape(n, 4)), b


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_large_roundtrip_and_linearity(log_n):
    """Full-size sweep points checked through size-independent properties: iNTT(NTT(a)) = n*a ... and
    spot values against the direct sum."""
    zkgpu.init(0)
    n = 1 << log_n
    a = O.random_fr(3, n)
    w = P.omega_for(log_n)
    wm = O.to_mont(0, P.int_to_limbs([w]))[0]
    f = zkgpu.best_fft(a, wm, log_n).reshape(n, 4)
    winv = O.to_mont(0, P.int_to_limbs([pow(w, -1, P.R_MOD)]))[0]
    back = zkgpu.best_fft(f, winv, log_n).reshape(n, 4)
    ninv = O.to_mont(0, P.int_to_limbs([pow(n, -1, P.R_MOD)]))
    assert np.array_equal(O.field_op(0, 0, back, np.repeat(ninv, n, axis=0)).reshape(n, 4), a)
    # f[0] = sum a, f[n/2] = sum (-1)^i a_i: check f[0] + f[n/2] = 2 * sum_even a via the oracle on a folded vector
    ai = P.limbs_to_int(O.from_mont(0, a[: 1 << 12]))  # spot-check one output by Horner on a sparse input instead
    sp = np.zeros((n, 4), dtype=np.uint64)
    sp[: 1 << 12] = a[: 1 << 12]
    fs = zkgpu.best_fft(sp, wm, log_n).reshape(n, 4)
    for idx in (1, 12345, n - 1):
        x = pow(w, idx, P.R_MOD)
        exp = 0
        for c in reversed(ai):
            exp = (exp * x + c) % P.R_MOD
        assert P.limbs_to_int(O.from_mont(0, fs[idx]))[0] == exp


@pytest.mark.parametrize("j,k", [(4, 11), (5, 11), (5, 12), (5, 13), (7, 13), (9, 13)])
def test_evaluation_domain_matches_oracle(j, k):
    zkgpu.init(0)
    d = zkgpu.EvaluationDomain(j, k)
    ek, _ = O.domain(j, k)
    assert d.extended_k == ek
    n = 1 << k
    a = O.random_fr(k * 31 + j, n)
    assert np.array_equal(d.lagrange_to_coeff(a).reshape(n, 4), O.domain_op(j, k, 0, a))
    assert np.array_equal(d.coeff_to_lagrange(a).reshape(n, 4), O.domain_op(j, k, 1, a))
    ext = d.coeff_to_extended(a)
    assert np.array_equal(ext, O.domain_op(j, k, 2, a, threads=8))
    e = O.random_fr(5, 1 << ek)
    assert np.array_equal(d.extended_to_coeff(e), O.domain_op(j, k, 3, e, threads=8))
    assert np.array_equal(d.extended_to_coeff(ext)[:n], a)


# ---- MSM ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 100, 1000, 2048])
def test_msm_matches_oracle(raw11, n):
    zkgpu.init(0)
    s = edge_scalars(n, 50 + n)
    got = zkgpu.best_multiexp(s, raw11["g"][:n])
    assert np.array_equal(got, O.msm(s, raw11["g"][:n], threads=8))


def test_msm_edge_cases(raw11):
    zkgpu.init(0)
    g = raw11["g"]
    n = 64
    one = O.to_mont(0, P.int_to_limbs([1]))[0]
    # all-zero scalars -> identity (0,0)
    assert not zkgpu.best_multiexp(np.zeros((n, 4), dtype=np.uint64), g[:n]).any()
    # all bases equal, all scalars 1 -> n * G  (exercises the doubling branch of the mixed add)
    same = np.repeat(g[3:4], n, axis=0)
    ones = np.repeat(one.reshape(1, 4), n, axis=0)
    assert np.array_equal(zkgpu.best_multiexp(ones, same), O.msm(ones, same))
    # P + (-P): scalars (1, r-1) on the same base -> identity
    two = np.repeat(g[7:8], 2, axis=0)
    sc = np.stack([one, O.to_mont(0, P.int_to_limbs([P.R_MOD - 1]))[0]])
    assert not zkgpu.best_multiexp(sc, two).any()
    # identity bases are skipped
    bases = g[:n].copy(); bases[5] = 0; bases[6] = 0
    s = edge_scalars(n, 9)
    assert np.array_equal(zkgpu.best_multiexp(s, bases), O.msm(s, bases))
    # all-(r-1) scalars and sparse (90% zero) scalars
    allmax = np.repeat(O.to_mont(0, P.int_to_limbs([P.R_MOD - 1])), 512, axis=0)
    assert np.array_equal(zkgpu.best_multiexp(allmax, g[:512]), O.msm(allmax, g[:512], threads=8))
    sp = O.random_fr(77, 2048)
    mask = np.random.default_rng(5).random(2048) < 0.9
    sp[mask] = 0
    assert np.array_equal(zkgpu.best_multiexp(sp, g), O.msm(sp, g, threads=8))
    small = O.to_mont(0, P.int_to_limbs([int(x) for x in np.random.default_rng(6).integers(0, 2**63, 2048)]))
    assert np.array_equal(zkgpu.best_multiexp(small, g), O.msm(small, g, threads=8))


def test_msm_known_answers_g_lagrange(raw11, params11):
    """halo2's own known answers: g_lagrange[i] = sum_j (omega^{-ij}/n) g[j] (fixture), via both the
    plain path and the fixed-base SRS path."""
    n, k = 2048, 11
    w_inv = pow(P.omega_for(k), -1, P.R_MOD)
    n_inv = pow(n, -1, P.R_MOD)
    for i in (0, 1, 1000, 2047):
        sc = O.to_mont(0, P.int_to_limbs([pow(w_inv, i * j, P.R_MOD) * n_inv % P.R_MOD for j in range(n)]))
        assert np.array_equal(zkgpu.best_multiexp(sc, raw11["g"]), raw11["g_lagrange"][i]), i
        assert np.array_equal(params11.commit(sc), raw11["g_lagrange"][i]), i


def test_commit_lagrange_property(raw11, params11):
    """crates/powers-of-tau/lib.rs:248-264 on the GPU: commit(lagrange_to_coeff(a)) == commit_lagrange(a)."""
    n = 2048
    a = O.to_mont(0, P.int_to_limbs(list(range(n))))
    b = zkgpu.EvaluationDomain(1, 11).lagrange_to_coeff(a)
    lhs, rhs = params11.commit(b), params11.commit_lagrange(a)
    assert np.array_equal(lhs, rhs)
    assert np.array_equal(lhs, O.msm(a, raw11["g_lagrange"], threads=8))


def test_srs_batch_commit(raw11, params11):
    n, m = 2048, 9
    s = O.random_fr(21, n * m).reshape(m, n, 4)
    s[1] = 0
    s[2, ::2] = 0
    s[3] = O.to_mont(0, P.int_to_limbs([1]))[0]
    for basis, key in ((0, "g"), (1, "g_lagrange")):
        got = params11.commit_batch(basis, s, n)
        for i in range(m):
            assert np.array_equal(got[i], O.msm(s[i], raw11[key], threads=8)), (basis, i)
    # shorter vectors than the SRS (best_multiexp over a prefix of the bases)
    got = params11.commit_batch(0, s[:, :1000].copy(), 1000)
    assert np.array_equal(got[0], O.msm(s[0, :1000], raw11["g"][:1000], threads=8))


def test_srs_batch_commit_structured_columns(raw11, params11):
    """More than 32 commitments per call take the throughput path, whose digit sort may commit a column through its differences
    against the prefix sums of the bases (csrc/msm.cuh MsmPlan::diff_offset): constant stretches, sorted columns, small negative
    values, columns for which the plain half stays cheaper, prefixes of the SRS — all equal to the oracle's best_multiexp."""
    n, m = 2048, 40
    r = P.R_MOD
    rng = np.random.default_rng(77)
    mont = lambda vals: O.to_mont(0, P.int_to_limbs([v % r for v in vals]))
    big = lambda: int.from_bytes(rng.bytes(40), "little") % r
    s = O.random_fr(23, n * m).reshape(m, n, 4)
    v = big()
    s[0] = mont([v])[0]                                         # constant, full width
    s[1] = mont([1])[0]                                         # constant one (a grand product without copies)
    steps = sorted(rng.choice(n, 5, replace=False).tolist()) + [n]
    col, lo = [], 0
    for hi in steps:
        col += [big()] * (hi - lo); lo = hi
    s[2] = mont(col)                                            # piecewise constant
    s[3] = mont(sorted(int(x) for x in rng.integers(0, 256, n)))   # a sorted lookup column
    s[4] = mont([r - 1])[0]                                     # -1 everywhere
    s[5] = mont([(r - int(x)) % r for x in rng.integers(0, 4, n)])   # small negative values
    s[6, 1::2] = 0                                              # alternating: twice as many changes as non-zero terms
    s[7] = 0
    s[7, 700:1500] = mont([big()])[0]                           # one stretch between zeros
    s[8] = 0
    s[9, :-1] = mont([big()])[0]                                # constant up to a different last element
    s[10] = mont(list(range(n)))                                # every difference is -1
    s[11] = mont([big()] * 5 + [0] * (n - 5))
    for basis, key in ((0, "g"), (1, "g_lagrange")):
        got = params11.commit_batch(basis, s, n)
        for i in range(14):
            assert np.array_equal(got[i], O.msm(s[i], raw11[key], threads=8)), (basis, i)
        for i in (20, 39):
            assert np.array_equal(got[i], O.msm(s[i], raw11[key], threads=8)), (basis, i)
    # prefixes of the SRS: the column ends before the bases do (the last difference is the last scalar itself)
    for cut in (1000, 1, 33):
        got = params11.commit_batch(1, s[:, :cut].copy(), cut)
        for i in (0, 1, 2, 3, 4, 9, 10, 12):
            assert np.array_equal(got[i], O.msm(s[i, :cut], raw11["g_lagrange"][:cut], threads=8)), (cut, i)


def test_g_to_lagrange_known_answer(raw11):
    """K6: reproduces the fixture's g_lagrange block (written by halo2's g_to_lagrange)."""
    zkgpu.init(0)
    assert np.array_equal(zkgpu.g_to_lagrange(raw11["g"], 11), raw11["g_lagrange"])


def test_fft_g1_small(raw11):
    zkgpu.init(0)
    log_n = 4
    n = 1 << log_n
    one = O.to_mont(1, P.int_to_limbs([1]))[0]
    jac = np.zeros((n, 12), dtype=np.uint64)
    jac[:, :8] = raw11["g"][:n]
    jac[:, 8:] = one
    w = P.omega_for(log_n)
    out = zkgpu.fft_g1(jac, O.to_mont(0, P.int_to_limbs([w]))[0], log_n).reshape(n, 12)
    for i in (0, 1, 7, 15):
        sc = O.to_mont(0, P.int_to_limbs([pow(w, i * j, P.R_MOD) for j in range(n)]))
        assert np.array_equal(out[i, :8], O.msm(sc, raw11["g"][:n])), i
        assert np.array_equal(out[i, 8:], one)
