"""BASELINE configs[1] and [2] at their full sizes (-m gpu): the standalone G1 MSM sweep 2^11..2^24 and the
size-independent checks of the NTT sweep.  Above the sizes the CPU oracle finishes in seconds the expected
value comes from a domain property instead: with the setup SRS g[i] = s^i * G,
    sum_i c_i * g[i] == [ p(s) ] * G ,   p(X) = sum_i c_i X^i
so one Horner evaluation and one scalar multiplication on the CPU pin an MSM of any size bit-exactly."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import pyref as P
import zkgpu
from zkgpu import multi
from zkgpu.gpu_backend import GpuBackend

pytestmark = pytest.mark.gpu
SEED = 7


@pytest.fixture(scope="module")
def big_bases():
    zkgpu.init(0)
    cache = {}

    def get(k):
        if k not in cache:
            cache.clear()
            cache[k] = zkgpu.params_setup(k, SEED, lagrange=False)[0]
        return cache[k]
    return get


def _expected(scalars):
    s = GpuBackend.random(SEED, 1)[0]               # the setup's toxic scalar: first Fr::random of SmallRng(SEED)
    ps = O.eval_polynomial(scalars, s)
    G = O.to_mont(1, P.int_to_limbs([1, 2])).reshape(8)
    return O.g1_op(2, G, ps)


def _scalars(kind, n):
    if kind == "uniform":
        return GpuBackend.random(11, n)
    if kind == "r_minus_1":
        return np.tile(GpuBackend.const(-1), (n, 1))
    if kind == "sparse":
        a = GpuBackend.random(12, n)
        a[np.random.default_rng(1).random(n) < 0.9] = 0
        return a
    if kind == "small64":
        a = np.zeros((n, 4), dtype=np.uint64)
        a[:, 0] = np.random.default_rng(2).integers(0, 1 << 63, n, dtype=np.uint64)
        out = np.empty_like(a)
        zkgpu._chk(zkgpu.lib().zkgpu_fr_to_mont(zkgpu._p(a), zkgpu._p(out), n))
        return out
    raise ValueError(kind)


@pytest.mark.parametrize("log_n", [11, 14, 17, 20, 22, 24])
def test_msm_sweep_known_answer(big_bases, log_n):
    n = 1 << log_n
    g = big_bases(log_n)
    c = _scalars("uniform", n)
    assert np.array_equal(zkgpu.best_multiexp(c, g), _expected(c)), log_n


@pytest.mark.parametrize("kind", ["r_minus_1", "sparse", "small64"])
def test_msm_scalar_families(big_bases, kind):
    n = 1 << 16
    g = big_bases(16)
    c = _scalars(kind, n)
    assert np.array_equal(zkgpu.best_multiexp(c, g), _expected(c)), kind


def test_msm_linearity_and_sharding(big_bases):
    """MSM(a + b) == MSM(a) + MSM(b); point-sharded partial results sum to the full MSM (2/4/8 shards)."""
    n = 1 << 18
    g = big_bases(18)
    a, b = GpuBackend.random(21, n), GpuBackend.random(22, n)
    full = zkgpu.best_multiexp(a, g)
    assert np.array_equal(zkgpu.best_multiexp(GpuBackend.add(a, b), g), zkgpu.g1_sum(np.stack([full, zkgpu.best_multiexp(b, g)])))
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = multi.shard_bounds(n, r, world)
            parts.append(zkgpu.best_multiexp(a[lo:hi], g[lo:hi]))
        assert np.array_equal(multi.combine_partials(np.stack(parts)), full), world


def test_msm_all_zero_and_identity_bases(big_bases):
    n = 1 << 12
    g = big_bases(12).copy()
    z = np.zeros((n, 4), dtype=np.uint64)
    assert not zkgpu.best_multiexp(z, g).any()                       # all-zero scalars -> identity (0,0)
    c = GpuBackend.random(3, n)
    g2 = g.copy(); g2[::3] = 0                                       # identity bases are skipped
    c2 = c.copy(); c2[::3] = 0
    assert np.array_equal(zkgpu.best_multiexp(c, g2), zkgpu.best_multiexp(c2, g))
    with pytest.raises(zkgpu.ZkGpuError):
        zkgpu.best_multiexp(c[:-1], g)                               # upstream assert_eq!(coeffs.len(), bases.len())


def test_prove_batch_empty(big_bases):
    from zkgpu import circuits
    shape = circuits.Shape("tiny")
    circ = circuits.Circuit(shape, GpuBackend, seed=1)
    g, gl = zkgpu.params_setup(shape.k, 42)
    params = zkgpu.ParamsKZG(shape.k, g, gl)
    pk = zkgpu.ProvingKey(params, circ.blob)
    assert pk.prove_batch(np.zeros((0, shape.num_advice, shape.n, 4), dtype=np.uint64), np.zeros((0, 3, 4), dtype=np.uint64), []) == []
    pk.release(); params.release()


def test_eval_polynomial_matches_oracle():
    """zkgpu_eval_polynomial = halo2 eval_polynomial: empty, length 1, ragged (non power of two), long"""
    zkgpu.init(0)
    x = GpuBackend.random(31, 1)[0]
    assert not zkgpu.eval_polynomial(np.zeros((0, 4), dtype=np.uint64), x).any()
    for n in (1, 2, 7, 1000, 8192, 70001, (1 << 18) + 3):
        c = GpuBackend.random(32 + n % 7, n)
        assert np.array_equal(zkgpu.eval_polynomial(c, x), O.eval_polynomial(c, x)), n


def test_setup_powers_is_a_slice_of_the_setup_srs(big_bases):
    g = big_bases(14)
    assert np.array_equal(zkgpu.setup_powers(SEED, 0, 100), g[:100])
    assert np.array_equal(zkgpu.setup_powers(SEED, 5000, 3000), g[5000:8000])


@pytest.mark.parametrize("log_n", [11, 16, 20])
def test_resident_bases_msm(big_bases, log_n):
    """zkgpu_msm_g1_bases (bases resident, scalars streamed or resident) == best_multiexp == [p(s)] G"""
    n = 1 << log_n
    g = big_bases(log_n)
    c = _scalars("uniform", n)
    bases = zkgpu.Bases(g)
    try:
        want = _expected(c)
        assert np.array_equal(bases.msm(c), want)
        assert np.array_equal(bases.msm(None), want)          # scalars left resident by the previous call
        assert bases.kernel_ms > 0
        c2 = _scalars("sparse", n)
        assert np.array_equal(bases.msm(c2), zkgpu.best_multiexp(c2, g))
        with pytest.raises(zkgpu.ZkGpuError):
            zkgpu._chk(zkgpu.lib().zkgpu_msm_g1_bases(C.c_uint64(bases.handle), zkgpu._p(c), C.c_size_t(n - 1), zkgpu._p(np.zeros(12, np.uint64)), None))
    finally:
        bases.release()
