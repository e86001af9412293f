"""Where does the host-buffer path lose time against the device-resident one?  (diagnostic, not the benchmark)"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zkos-monorepo_b200"))
import zkgpu
from zkgpu import circuits
from zkgpu.gpu_backend import GpuBackend

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
zkgpu.init(0)
L = zkgpu.lib()
shape = circuits.Shape("withdraw")
g, gl = zkgpu.params_setup(shape.k, 42)
params = zkgpu.ParamsKZG(shape.k, g, gl)
circ = circuits.Circuit(shape, GpuBackend, seed=3)
pk = zkgpu.ProvingKey(params, circ.blob)
A, n = shape.num_advice, shape.n
h = torch.empty((M, A, n, 4), dtype=torch.int64, pin_memory=True)
hv = h.numpy().view(np.uint64)
a1, p1 = circ.witness(1)
hv[:] = a1
inst = np.broadcast_to(p1, (M,) + p1.shape).copy()
seeds = np.arange(M, dtype=np.uint64) + 1
d = h.cuda()
torch.cuda.synchronize()
for _ in range(2):
    t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("raw H2D pinned: %.1f GB in %.3f s = %.1f GB/s" % (h.numel() * 8 / 1e9, dt, h.numel() * 8 / 1e9 / dt))
proofs = np.zeros(M * pk.proof_len, dtype=np.uint8)
steps = (C.c_double * 8)()


def run(kind):
    L.zkgpu_prover_step_seconds(steps, 1)
    t = time.perf_counter()
    if kind == "dev":
        pk.prove_batch_dev(d.data_ptr(), inst, seeds, out=proofs)
    else:
        zkgpu._chk(L.zkgpu_prove_batch(C.c_uint64(pk.handle), C.c_void_p(h.data_ptr()), inst.ctypes.data_as(C.c_void_p), C.c_size_t(shape.num_pi),
                                       C.c_size_t(M), seeds.ctypes.data_as(C.c_void_p), proofs.ctypes.data_as(C.c_void_p), C.c_size_t(pk.proof_len)))
    dt = time.perf_counter() - t
    L.zkgpu_prover_step_seconds(steps, 1)
    return dt, " ".join("%.3f" % x for x in steps[:7])


for workers in ("2", "1"):
    os.environ["ZKGPU_PROVER_WORKERS"] = workers
    for kind in ("dev", "host", "dev", "host", "dev", "host"):
        dt, st = run(kind)
        print("workers=%s %-4s %.3f s  %.1f proofs/s   step seconds (summed over workers): %s" % (workers, kind, dt, M / dt, st))
