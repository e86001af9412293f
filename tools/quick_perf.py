"""Quick kernel-level timing of the device-resident NTT and fixed-base MSM paths (CUDA events on the
launching stream).  Development aid; bench.py is the contract benchmark."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_lib as O
import pyref as P
import zkgpu


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    zkgpu.init(0)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for log_n, m in ((11, 4096), (12, 2048), (13, 2048), (15, 512), (16, 256), (20, 16), (22, 4)):
        n = 1 << log_n
        a = torch.from_numpy(O.random_fr(1, 1 << 12).view(np.int64)).cuda().repeat((n * m) >> 12, 1).contiguous()
        scratch = torch.empty_like(a)
        w = O.to_mont(0, P.int_to_limbs([P.omega_for(log_n)]))[0]
        best, med = timeit(lambda: zkgpu.ntt_batch_dev(a.data_ptr(), w, log_n, m, scratch.data_ptr(), st))
        res["ntt_2^%d_x%d" % (log_n, m)] = dict(ms=best, GBps=64.0 * n * m / best / 1e6, Mmul_per_s=n / 2 * log_n * m / best / 1e3)
    raw = O.srs_read(O.RAW11, 0)
    # k=13 synthetic SRS: reuse the 2048 real points 4x (perf only)
    g = np.tile(raw["g"], (4, 1)); gl = np.tile(raw["g_lagrange"], (4, 1))
    params = zkgpu.ParamsKZG(13, g, gl)
    n = 8192
    for m in (1, 32, 256, 1024):
        s = torch.from_numpy(O.random_fr(2, n * 8).view(np.int64)).cuda().repeat((m + 7) // 8, 1)[: n * m].contiguous()
        out = torch.empty((m, 8), dtype=torch.int64, device="cuda")
        best, med = timeit(lambda: params.commit_batch_dev(1, s.data_ptr(), n, m, out.data_ptr(), st), iters=3, warm=1)
        res["msm_srs_2^13_x%d" % m] = dict(ms=best, msm_per_s=m / best * 1e3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
