"""The prover's rng seam, per-proof failure isolation and the coalescing single-proof call (-m gpu).

`generate_proof(.., rng: &mut impl RngCore)` (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111) is
called with a RUNNING SmallRng that has already produced the SRS and the witness in the reference's seeded tests
(/root/reference/crates/halo2-verifier/src/generator.rs:117-130) and with OsRng / thread_rng in production
(crates/shielder-account/src/call_data.rs:499, crates/shielder_bindings/src/circuits/deposit.rs:108); a request that
makes the prover fail fails alone (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:189-190)."""
import threading

import numpy as np
import pytest

import oracle_lib as O
import zkgpu
from zkgpu import circuits

pytestmark = pytest.mark.gpu


def _setup(name, srs=None, seed=1, **kw):
    shape = circuits.Shape(name, **kw)
    circ = circuits.Circuit(shape, O.OracleBackend, seed=seed)
    srs = srs or O.downsized_srs(shape.k)
    po = O.PlonkOracle(circ.blob, srs, threads=8)
    params = zkgpu.ParamsKZG(shape.k, srs["g"], srs["g_lagrange"])
    pk = zkgpu.ProvingKey(params, circ.blob)
    return shape, circ, po, params, pk


@pytest.fixture(scope="module")
def tiny():
    zkgpu.init(0)
    s = _setup("tiny")
    yield s
    s[4].release(); s[3].release()


@pytest.fixture(scope="module")
def tiny_lookup():
    zkgpu.init(0)
    s = _setup("tiny_lookup", seed=4)
    yield s
    s[4].release(); s[3].release()


def test_running_smallrng_state_mode_matches_oracle(tiny_lookup):
    """mode 1: the proof continues the caller's xoshiro256++ stream and hands the advanced state back"""
    shape, circ, po, params, pk = tiny_lookup
    wits = [circ.witness(70 + i) for i in range(5)]
    adv = np.stack([w[0] for w in wits]); inst = np.stack([w[1] for w in wits])
    states = np.stack([O.smallrng_state(1000 + i) for i in range(5)])
    want_states = states.copy()
    want = [po.prove_rng(adv[i], inst[i], 1, want_states[i]) for i in range(5)]
    proofs, status = pk.prove_batch_rng(adv, inst, pk.RNG_XOSHIRO_STATE, states)
    assert status.tolist() == [0] * 5
    assert proofs == want
    assert np.array_equal(states, want_states), "rng state handed back differs from the CPU prover's"
    # a fresh seed_from_u64 is the same stream as its state
    assert proofs[0] == po.prove(adv[0], inst[0], seed=1000)
    # and proving again with the advanced state gives a different, valid proof (the stream moved on)
    again, _ = pk.prove_batch_rng(adv, inst, pk.RNG_XOSHIRO_STATE, states)
    assert again[0] != proofs[0] and po.verify(again[0], inst[0])
    assert again[0] == po.prove_rng(adv[0], inst[0], 1, want_states[0])


def test_chacha20_seed_mode_matches_oracle(tiny_lookup):
    """mode 2: 32 bytes of caller entropy per proof, expanded by ChaCha20Rng::from_seed"""
    shape, circ, po, params, pk = tiny_lookup
    wits = [circ.witness(80 + i) for i in range(3)]
    adv = np.stack([w[0] for w in wits]); inst = np.stack([w[1] for w in wits])
    seeds = np.random.default_rng(5).integers(0, 256, (3, 32), dtype=np.uint8)
    proofs, status = pk.prove_batch_rng(adv, inst, pk.RNG_CHACHA20_SEED, seeds.copy())
    assert status.tolist() == [0, 0, 0]
    for i in range(3):
        assert proofs[i] == po.prove_rng(adv[i], inst[i], 2, seeds[i].copy()), i
        assert po.verify(proofs[i], inst[i])
    assert len(set(proofs)) == 3


def test_replay_of_the_reference_seeded_flow_from_one_stream():
    """generator.rs:117-130: `let mut rng = rng();` (SmallRng seed 42) -> generate_setup_params(k, &mut rng) -> the witness drawn
    from the same rng -> generate_proof(.., &mut rng).  One running state goes through setup, witness randomness and the prover on
    the GPU and, independently, through the CPU restatement: same SRS, same witness seed, same proof bytes, same final state."""
    zkgpu.init(0)
    k = 6
    st_gpu, st_cpu = O.smallrng_state(42), O.smallrng_state(42)
    g, gl = zkgpu.params_setup_rng(k, st_gpu)
    srs = O.params_setup_rng(k, st_cpu)
    assert np.array_equal(g, srs["g"]) and np.array_equal(gl, srs["g_lagrange"]) and np.array_equal(st_gpu, st_cpu)
    # "random_correct_example(&mut rng)": the witness is derived from the next field element of the stream
    w_gpu = zkgpu.fr_random_rng(st_gpu, 1)
    w_cpu = O.random_fr_rng(st_cpu, 1)
    assert np.array_equal(w_gpu, w_cpu) and np.array_equal(st_gpu, st_cpu)
    shape, circ, po, params, pk = _setup("tiny", srs=srs)
    try:
        adv, pi = circ.witness(int(w_gpu[0, 0] & np.uint64(0xFFFFFF)))
        proofs, status = pk.prove_batch_rng(adv[None], pi[None], pk.RNG_XOSHIRO_STATE, st_gpu.reshape(1, 4))
        want = po.prove_rng(adv, pi, 1, st_cpu)
        assert status[0] == 0 and proofs[0] == want and po.verify(want, pi)
        assert np.array_equal(st_gpu, st_cpu)
        assert proofs[0] != po.prove(adv, pi, seed=42)      # not a fresh seed-42 stream: the rng had moved on
    finally:
        pk.release(); params.release()


def test_one_bad_witness_fails_alone(tiny_lookup):
    """128 proofs, one of them with a lookup input outside its table: 127 good proofs come back, byte-identical to the
    CPU prover's, and the bad one carries a status instead of failing the call"""
    shape, circ, po, params, pk = tiny_lookup
    m, bad = 128, 77
    wits = [circ.witness(200 + i) for i in range(m)]
    adv = np.stack([w[0] for w in wits]); inst = np.stack([w[1] for w in wits])
    adv[bad, shape.lv[0], 3] = O.OracleBackend.const(shape.table_size + 5)
    seeds = np.arange(m, dtype=np.uint64) + 5000
    proofs, status = pk.prove_batch_rng(adv, inst, pk.RNG_SEED_U64, seeds)
    assert status[bad] == pk.PROOF_LOOKUP_FAILED and proofs[bad] == b""
    assert int((status == 0).sum()) == m - 1
    with pytest.raises(RuntimeError):      # the CPU prover refuses the same witness
        po.prove(adv[bad], inst[bad], seed=int(seeds[bad]))
    for i in list(range(0, m, 9)) + [bad - 1, bad + 1]:
        assert proofs[i] == po.prove(adv[i], inst[i], seed=int(seeds[i])), i
    good = [i for i in range(m) if i != bad]
    assert po.verify_batch([proofs[i] for i in good], inst[good], threads=8) == (True, 0)
    # the all-or-nothing test form reports the failure as an error
    with pytest.raises(zkgpu.ZkGpuError):
        pk.prove_batch(adv[bad - 1:bad + 1], inst[bad - 1:bad + 1], seeds[bad - 1:bad + 1])


def test_concurrent_single_proof_callers_are_coalesced(tiny_lookup):
    """100 threads call zkgpu_prove at once (the reference's prover server: one task per client, max 100 in flight): every
    caller gets the proof the CPU prover produces for ITS witness and rng, the requests were served in far fewer batches,
    and the one malformed request fails alone."""
    shape, circ, po, params, pk = tiny_lookup
    m, bad = 100, 31
    wits = [circ.witness(400 + i) for i in range(m)]
    wits[bad][0][shape.lv[0], 5] = O.OracleBackend.const(shape.table_size + 9)
    seeds = np.random.default_rng(9).integers(0, 256, (m, 32), dtype=np.uint8)
    got, errs = {}, {}
    before = pk.prove_stats()
    gate = threading.Barrier(m)

    def client(i):
        gate.wait()
        try:
            got[i] = pk.prove_one(wits[i][0], wits[i][1], pk.RNG_CHACHA20_SEED, seeds[i].copy())
        except zkgpu.ZkGpuError as e:
            errs[i] = str(e)
    ts = [threading.Thread(target=client, args=(i,)) for i in range(m)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    after = pk.prove_stats()
    assert list(errs) == [bad] and "error -5" in errs[bad]
    assert len(got) == m - 1
    for i in range(0, m, 7):
        if i != bad:
            assert got[i] == po.prove_rng(wits[i][0], wits[i][1], 2, seeds[i].copy()), i
    good = [i for i in range(m) if i != bad]
    assert po.verify_batch([got[i] for i in good], np.stack([wits[i][1] for i in good]), threads=8) == (True, 0)
    assert after["requests"] - before["requests"] == m
    assert after["batches"] - before["batches"] < m // 2, "requests were not coalesced"
