"""N > 1 host logic on CPU: world_size-2 gloo run of the point-sharded MSM (gather of one partial point per
rank + host-side sum) and of the request sharding helpers.  The per-rank shard MSM is computed by the CPU
oracle here (there is no GPU in this container); on the GPU box the same `msm_sharded` is driven with
`zkgpu.best_multiexp` (tests/test_gpu_sweeps.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret_dir):
    for p in (ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle_lib as O
    from zkgpu import multi
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    raw = O.srs_read(O.RAW11, 0)
    scalars = O.random_fr(5, n)
    bases = raw["g"][:n]
    lo, hi = multi.shard_bounds(n, rank, world)
    got = multi.msm_sharded(scalars[lo:hi], bases[lo:hi], dist, lambda c, b: O.msm(c, b))   # a rank is handed its shard only
    want = O.msm(scalars, bases, threads=2)
    np.save(os.path.join(ret_dir, "r%d.npy" % rank), np.concatenate([got, want, np.array([lo, hi], dtype=np.uint64)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 5, 300])
def test_sharded_msm_gloo_world2(tmp_path, n):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() + n) % 300
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(str(tmp_path / ("r%d.npy" % i))) for i in range(2)]
    for x in r:
        assert np.array_equal(x[:8], x[8:16])            # sharded result == full MSM, on every rank
    assert np.array_equal(r[0][:8], r[1][:8])
    assert int(r[0][16]) == 0 and int(r[0][17]) == int(r[1][16]) and int(r[1][17]) == n   # shards tile [0, n)


def test_shard_helpers():
    from zkgpu import multi
    for n in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 4, 8):
            spans = [multi.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            rr = sorted(i for r in range(world) for i in multi.shard_round_robin(n, r, world))
            assert rr == list(range(n))


def test_host_point_sum_matches_oracle():
    """zkgpu_g1_sum_affine runs on the host (no GPU): identity handling, P + P, P + (-P)."""
    import oracle_lib as O
    import zkgpu
    raw = O.srs_read(O.RAW11, 0)
    g = raw["g"]
    ident = np.zeros(8, dtype=np.uint64)
    neg = g[3].copy()
    neg[4:] = O.field_op(1, 2, np.zeros((1, 4), dtype=np.uint64), g[3][4:][None]).reshape(4)
    cases = [g[:1], g[:7], np.stack([g[2], g[2]]), np.stack([g[3], neg]), np.stack([ident, g[5], ident]), np.zeros((0, 8), dtype=np.uint64)]
    for pts in cases:
        want = ident.copy()
        for p in pts:
            want = O.g1_op(0, want, p)
        assert np.array_equal(zkgpu.g1_sum(pts), want)
