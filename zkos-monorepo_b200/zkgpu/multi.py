"""Multi-GPU host logic: one process per GPU, no collective on the proof path.

* Proof batches / request streams shard by rank (`shard_bounds`, `shard_round_robin`): proofs are independent,
  the SRS and the proving key are replicated.
* The standalone large MSM splits its points into contiguous shards; every rank HOLDS ONLY ITS SHARD, reduces it to ONE
  point, and the partial results (64 bytes each) are gathered and summed (`msm_sharded`).  The gather is the only
  exchange step of the whole path; its payload is O(100 B) per GPU, so a plain all_gather is used.  (With several
  devices in ONE process the library does the same with peer copies over NVLink: `zkgpu.Bases`, csrc/msm_sharded.cu.)
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n items for `rank` of `world` (first n % world shards one longer)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_round_robin(n, rank, world):
    """indices of a request stream served by `rank`: request i -> GPU i mod world"""
    return list(range(rank, n, world))


def combine_partials(partials):
    """sum of per-rank partial MSM results, (world, 8) affine -> (8,) affine"""
    from . import g1_sum
    return g1_sum(np.ascontiguousarray(partials, dtype=np.uint64))


def gather_points(local_point, dist, device=None):
    """all_gather of one affine point (8 u64) per rank -> (world, 8)"""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local_point, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


def msm_sharded(shard_coeffs, shard_bases, dist, local_msm, device=None):
    """best_multiexp over points split across ranks.  `shard_coeffs` / `shard_bases` are THIS RANK'S contiguous shard only
    (`shard_bounds(n, rank, world)` of the whole problem; a rank never holds the other shards); `local_msm(coeffs, bases)`
    reduces the shard to one affine point on this rank's GPU.  The 64-byte partial results are all-gathered and summed;
    every rank returns the full result."""
    c = np.asarray(shard_coeffs).reshape(-1, 4)
    part = local_msm(c, np.asarray(shard_bases).reshape(-1, 8)) if c.shape[0] else np.zeros(8, dtype=np.uint64)
    return combine_partials(gather_points(part, dist, device))
