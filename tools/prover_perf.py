"""Quick GPU timing of the batched prover (not the benchmark): withdraw shape at k=13, per-step seconds."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zkos-monorepo_b200"))
import zkgpu
from zkgpu import circuits
from zkgpu.gpu_backend import GpuBackend

m = int(sys.argv[1]) if len(sys.argv) > 1 else 128
name = sys.argv[2] if len(sys.argv) > 2 else "withdraw"
zkgpu.init(0)
shape = circuits.Shape(name)
t = time.time()
g, gl = zkgpu.params_setup(shape.k, 42)
params = zkgpu.ParamsKZG(shape.k, g, gl)
print("srs setup+register %.2fs" % (time.time() - t))
t = time.time()
circ = circuits.Circuit(shape, GpuBackend, seed=3)
pk = zkgpu.ProvingKey(params, circ.blob)
print("circuit+keygen %.2fs; sub_batch=%d proof_len=%d msm/proof=%d" % (time.time() - t, pk.sub_batch, pk.proof_len, shape.num_msm))
adv1, pi1 = circ.witness(1)
adv = np.broadcast_to(adv1, (m,) + adv1.shape).copy()
pi = np.broadcast_to(pi1, (m,) + pi1.shape).copy()
seeds = np.arange(m, dtype=np.uint64) + 1
steps = (C.c_double * 8)()
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    zkgpu.lib().zkgpu_prover_step_seconds(steps, 1)
    t = time.time()
    proofs = pk.prove_batch(adv, pi, seeds)
    dt = time.time() - t
    zkgpu.lib().zkgpu_prover_step_seconds(steps, 1)
    print("iter %d: %d proofs in %.3fs = %.1f proofs/s; steps(s): %s" % (it, m, dt, m / dt, " ".join("%.0f" % (1e6 * x) for x in steps[:8]) + " us"))
