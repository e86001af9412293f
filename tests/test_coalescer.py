"""Host logic of the request coalescer (no GPU): concurrent single-proof calls are served in batches, each caller
gets its own proof, a failing batch fails only its requests."""
import threading

import numpy as np
import pytest

from zkgpu.coalescer import ProofCoalescer


def fake_prove_batch(adv, inst, seeds):
    assert adv.shape[0] == inst.shape[0] == seeds.shape[0]
    if (seeds == 666).any():
        raise ValueError("bad witness")
    return [b"proof-%d-%d" % (int(s), int(a.sum())) for a, s in zip(adv, seeds)]


def test_concurrent_requests_are_batched():
    co = ProofCoalescer(fake_prove_batch, max_batch=16, max_wait_ms=50)
    out = {}

    def client(i):
        out[i] = co.prove(np.full((2, 4, 4), i, dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64), 100 + i)
    ts = [threading.Thread(target=client, args=(i,)) for i in range(40)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    co.close()
    assert out == {i: b"proof-%d-%d" % (100 + i, 32 * i) for i in range(40)}
    assert sum(co.batches) == 40 and max(co.batches) <= 16 and len(co.batches) < 40      # batched, never above max_batch


def test_failed_batch_fails_only_its_requests():
    co = ProofCoalescer(fake_prove_batch, max_batch=4, max_wait_ms=1)
    z = np.zeros((1, 1, 4), dtype=np.uint64)
    with pytest.raises(ValueError):
        co.prove(z, z[0], 666)
    assert co.prove(z, z[0], 7) == b"proof-7-0"
    co.close()
    with pytest.raises(RuntimeError):
        co.prove(z, z[0], 8)
