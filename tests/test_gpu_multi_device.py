"""Several devices in ONE process (-m gpu, needs >= 2 GPUs; skipped otherwise): zkgpu_init(device_mask) replicates the key material,
zkgpu_prove_batch_rng shards a batch across the devices, zkgpu_prove's dispatchers run on all of them, and the resident-bases MSM
splits its points and sums the partial points over the peer link (SURVEY.md 8b `zkgpu_init(device_mask)`, 8e).
Runs in a child process because the device mask is fixed per process and the other GPU tests select device 0 only."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys, threading
import numpy as np
sys.path[:0] = [%(root)r, %(root)r + "/zkos-monorepo_b200", %(root)r + "/tests"]
import oracle_lib as O
import zkgpu
from zkgpu import circuits
from zkgpu.gpu_backend import GpuBackend

G = %(gpus)d
zkgpu.init(mask=(1 << G) - 1)
assert zkgpu.device_count() == G
out = {}
shape = circuits.Shape("small_lookup")
circ = circuits.Circuit(shape, O.OracleBackend, seed=4)
srs = O.downsized_srs(shape.k)
po = O.PlonkOracle(circ.blob, srs, threads=8)
params = zkgpu.ParamsKZG(shape.k, srs["g"], srs["g_lagrange"])
pk = zkgpu.ProvingKey(params, circ.blob)
assert pk.replicas == G
m = 4 * G + 3                                   # ragged shards
wits = [circ.witness(50 + i) for i in range(m)]
adv = np.stack([w[0] for w in wits]); inst = np.stack([w[1] for w in wits])
bad = m - 2
adv[bad, shape.lv[0], 3] = O.OracleBackend.const(shape.table_size + 5)
seeds = np.random.default_rng(3).integers(0, 256, (m, 32), dtype=np.uint8)
proofs, status = pk.prove_batch_rng(adv, inst, pk.RNG_CHACHA20_SEED, seeds.copy())
out["batch_ok"] = all(proofs[i] == po.prove_rng(adv[i], inst[i], 2, seeds[i].copy()) for i in range(m) if i != bad)
out["bad_alone"] = bool(status[bad] == 1 and proofs[bad] == b"" and int((status == 0).sum()) == m - 1)
# concurrent single-proof callers: dispatchers of every device take work
got = {}
def client(i):
    got[i] = pk.prove_one(wits[i][0] if i != bad else wits[0][0], wits[i][1] if i != bad else wits[0][1], pk.RNG_CHACHA20_SEED, seeds[i].copy())
ts = [threading.Thread(target=client, args=(i,)) for i in range(m)]
[t.start() for t in ts]; [t.join() for t in ts]
out["prove_one_ok"] = all(got[i] == proofs[i] for i in range(m) if i != bad)
out["dispatchers"] = pk.prove_stats()["dispatchers"]
# point-sharded MSM with resident bases: shard g on device g, partial points summed on device 0
n = 1 << 16
g = zkgpu.params_setup(16, 7, lagrange=False)[0]
c = GpuBackend.random(11, n)
bases = zkgpu.Bases(g)
want = zkgpu.best_multiexp(c, g)
out["msm_ok"] = bool(np.array_equal(bases.msm(c), want) and np.array_equal(bases.msm(None), want))
s = GpuBackend.random(7, 1)[0]
gen = zkgpu.setup_powers(7, 0, 1)
out["msm_identity_ok"] = bool(np.array_equal(want, zkgpu.best_multiexp(zkgpu.eval_polynomial(c, s)[None], gen)))
print(json.dumps(out))
'''


def test_two_devices_one_process():
    import torch
    gpus = torch.cuda.device_count()
    if gpus < 2:
        pytest.skip("needs at least two GPUs")
    gpus = min(gpus, 8)
    res = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "gpus": gpus}], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["batch_ok"] and out["bad_alone"] and out["prove_one_ok"] and out["msm_ok"] and out["msm_identity_ok"], out
    assert out["dispatchers"] == 3 * gpus
