"""The oracle's rng modes (CPU): the checker must offer the same `rng: &mut impl RngCore` seam as the product
(/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111; running SmallRng in
/root/reference/crates/halo2-verifier/src/generator.rs:117-130)."""
import numpy as np
import pytest

import oracle_lib as O
from zkgpu import circuits

MASK = (1 << 64) - 1


def _rotl(x, n):
    return ((x << n) | (x >> (64 - n))) & MASK


def _xoshiro_next(s):
    """rand 0.8.5 Xoshiro256PlusPlus::next_u64 on a list of four ints (in place)"""
    r = (_rotl((s[0] + s[3]) & MASK, 23) + s[0]) & MASK
    t = (s[1] << 17) & MASK
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
    s[2] ^= t; s[3] = _rotl(s[3], 45)
    return r


@pytest.fixture(scope="module")
def tiny_lookup():
    shape = circuits.Shape("tiny_lookup")
    circ = circuits.Circuit(shape, O.OracleBackend, seed=4)
    po = O.PlonkOracle(circ.blob, O.downsized_srs(shape.k), threads=4)
    return shape, circ, po


def test_seed_state_is_splitmix_and_matches_the_stream():
    st = [int(x) for x in O.smallrng_state(42)]
    assert [_xoshiro_next(st) for _ in range(16)] == [int(x) for x in O.smallrng(42, 16)]


def test_state_mode_continues_the_stream(tiny_lookup):
    shape, circ, po = tiny_lookup
    adv, pi = circ.witness(3)
    state = O.smallrng_state(7)
    proof = po.prove_rng(adv, pi, 1, state)
    assert proof == po.prove(adv, pi, seed=7) and po.verify(proof, pi)
    # number of u64 the prover consumed: Fr::random = 8; draws in create_proof's order (advice rows + Blinds, lookup permuted
    # rows + Blinds, permutation z rows + Blind, lookup z rows + Blind, one 32-byte ChaCha seed, Blind, quotient-piece Blinds)
    A, L, P, Q, bf = shape.num_advice, shape.n_lookup, shape.num_perm_sets, shape.num_quotients, shape.blinding_factors
    wides = A * (bf + 1) + A + L * (2 * (bf + 1) + 2) + P * (bf + 1) + L * (bf + 1) + 1 + Q
    st = [int(x) for x in O.smallrng_state(7)]
    for _ in range(8 * wides + 4):
        _xoshiro_next(st)
    assert st == [int(x) for x in state], "state after the proof is not the stream position create_proof reaches"
    # the next proof from the running state differs and verifies
    second = po.prove_rng(adv, pi, 1, state)
    assert second != proof and po.verify(second, pi)


def test_chacha_seed_mode(tiny_lookup):
    shape, circ, po = tiny_lookup
    adv, pi = circ.witness(4)
    seed = np.arange(32, dtype=np.uint8)
    a = po.prove_rng(adv, pi, 2, seed.copy())
    b = po.prove_rng(adv, pi, 2, seed.copy())
    seed2 = seed.copy(); seed2[31] ^= 1
    c = po.prove_rng(adv, pi, 2, seed2)
    assert a == b and a != c and po.verify(a, pi) and po.verify(c, pi)
    assert np.array_equal(seed, np.arange(32, dtype=np.uint8))      # a seed is not a state: left untouched


def test_setup_and_field_draws_advance_the_state():
    st = O.smallrng_state(42)
    srs = O.params_setup_rng(5, st, threads=2)
    ref = O.params_setup(5, 42, threads=2)
    assert np.array_equal(srs["g"], ref["g"]) and np.array_equal(srs["g_lagrange"], ref["g_lagrange"])
    want = [int(x) for x in O.smallrng_state(42)]
    for _ in range(8):
        _xoshiro_next(want)
    assert [int(x) for x in st] == want
    drawn = O.random_fr_rng(st, 3)
    assert np.array_equal(drawn, O.fr_from_u512(np.array([_xoshiro_next(want) for _ in range(24)], dtype=np.uint64)))
    assert [int(x) for x in st] == want
