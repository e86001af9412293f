#!/usr/bin/env python3
"""BASELINE configs[4]: a mixed new_account / deposit / withdraw proof stream sharded across 1/2/4/8 B200.

4096 requests (default), types drawn 1:2:2 from seed 7, request i is served by GPU
i mod world (round-robin, no collective); each rank groups its requests by circuit and proves them in batches
through zkgpu_prove_batch.  Run under torchrun for N > 1.  Prints one JSON line on rank 0.
  python tools/mixed_stream.py [--requests 4096] [--check 2]   # --check: verify that many proofs per type with the CPU oracle
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

TYPES = ["new_account", "deposit", "withdraw"]
WEIGHTS = [1, 2, 2]


def request_stream(n, seed=7):
    rng = np.random.default_rng(seed)
    return rng.choice(len(TYPES), size=n, p=np.array(WEIGHTS) / sum(WEIGHTS))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=4096)
    ap.add_argument("--distinct", type=int, default=32, help="distinct witnesses generated per circuit (cycled; seeds stay distinct)")
    ap.add_argument("--check", type=int, default=0)
    args = ap.parse_args()
    import torch
    import zkgpu
    from zkgpu import circuits, multi
    from zkgpu.gpu_backend import GpuBackend

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()                 # NCCL builds its communicator lazily on the first collective: do that here, not inside
        torch.cuda.synchronize()       # the timed region (the data path itself has no collective)
    zkgpu.init(local)
    g13, gl13 = zkgpu.params_setup(13, 42)
    params = {13: zkgpu.ParamsKZG(13, g13, gl13)}
    g12 = g13[:4096].copy()                                   # ParamsKZG::downsize(12)
    params[12] = zkgpu.ParamsKZG(12, g12, zkgpu.g_to_lagrange(g12, 12))
    pks, circs, wits = {}, {}, {}
    for t in TYPES:
        shape = circuits.Shape(t)
        circs[t] = circuits.Circuit(shape, GpuBackend, seed=3)
        pks[t] = zkgpu.ProvingKey(params[shape.k], circs[t].blob)
        wits[t] = [circs[t].witness(500 + i) for i in range(args.distinct)]
    stream = request_stream(args.requests)
    mine = multi.shard_round_robin(args.requests, rank, world)
    by_type = {t: [i for i in mine if TYPES[stream[i]] == t] for t in TYPES}
    batches, pinned = {}, {}
    for t, idx in by_type.items():
        if not idx:
            batches[t] = (None, None, None)
            continue
        # request payloads staged in pinned host memory, as a serving host would hold them (bench.py's e2e leg does the same)
        shape = circuits.Shape(t)
        pinned[t] = torch.empty((len(idx), shape.num_advice, shape.n, 4), dtype=torch.int64, pin_memory=True)
        adv = pinned[t].numpy().view(np.uint64)
        for j, i in enumerate(idx):
            adv[j] = wits[t][i % args.distinct][0]
        inst = np.stack([wits[t][i % args.distinct][1] for i in idx])
        batches[t] = (adv, inst, np.array(idx, dtype=np.uint64) + 1)
    for t in TYPES:                                            # warm-up: enough sub-batches per circuit that every pipeline
        if by_type[t]:                                         # worker's workspace and prefetch buffer exist before the timed region
            w = min(len(by_type[t]), 7 * pks[t].sub_batch)   # 7: all three workers run, each with a background advice prefetch
            pks[t].prove_batch(batches[t][0][:w], batches[t][1][:w], batches[t][2][:w])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    proofs, by_type_s = {}, {}
    for t in TYPES:
        if by_type[t]:
            tt0 = time.perf_counter()
            proofs[t] = pks[t].prove_batch(*batches[t])
            by_type_s[t] = round(time.perf_counter() - tt0, 3)
    torch.cuda.synchronize()
    print("rank %d: %s" % (rank, {t: (len(by_type[t]), by_type_s.get(t)) for t in TYPES}), file=sys.stderr, flush=True)
    dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    checked = 0
    if args.check and rank == 0:
        import oracle_lib as O
        srs13 = O.params_setup(13, 42)
        for t in TYPES:
            shape = circuits.Shape(t)
            srs = srs13 if shape.k == 13 else O.downsized_srs(12, srs13)
            po = O.PlonkOracle(circs[t].blob, srs, threads=os.cpu_count() or 1)
            for j in range(min(args.check, len(by_type[t]))):
                assert po.verify(proofs[t][j], batches[t][1][j]), (t, j)
                assert proofs[t][j] == po.prove(batches[t][0][j], batches[t][1][j], seed=int(batches[t][2][j])), (t, j)
                checked += 1
    if rank == 0:
        print(json.dumps({"workload": "mixed new_account/deposit/withdraw stream (BASELINE configs[4])", "requests": args.requests, "n_gpus": world,
                          "mix": {t: int((stream == i).sum()) for i, t in enumerate(TYPES)}, "seconds": dt, "proofs_per_s": args.requests / dt,
                          "verified_and_byte_checked": checked}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
