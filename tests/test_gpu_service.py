"""The library as a prover SERVICE (-m gpu): a plain C program with 100 pthreads calls zkgpu_prove concurrently — the call shape of
the reference's prover server (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:157-195, at most 100 requests in flight,
tee/crates/shielder-prover-server/src/command_line_args.rs:24-27).  The coalescer inside libzkgpu must turn them into batches:
every proof verifies, equals the batched call's proof for the same request, and the concurrent throughput stays within 20 % of
the batched throughput."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import zkgpu
from zkgpu import circuits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_client(tmp_path):
    exe = str(tmp_path / "coalesce_bench")
    libdir = os.path.dirname(zkgpu.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "coalesce_bench.c"),
                           "-o", exe, "-L", libdir, "-lzkgpu", "-lpthread", "-Wl,-rpath," + libdir])
    return exe


def write_input(path, shape, circ, srs, wits):
    with open(path, "wb") as f:
        f.write(struct.pack("<IQ", shape.k, len(circ.blob)))
        f.write(circ.blob)
        f.write(np.ascontiguousarray(srs["g"]).tobytes())
        f.write(np.ascontiguousarray(srs["g_lagrange"]).tobytes())
        f.write(struct.pack("<III", len(wits), shape.num_advice, shape.num_pi))
        for adv, pi in wits:
            f.write(np.ascontiguousarray(adv).tobytes())
            f.write(np.ascontiguousarray(pi).tobytes())


@pytest.mark.parametrize("name,total,min_ratio", [("small", 600, None), ("withdraw", 1200, 0.8)])
def test_hundred_concurrent_callers(tmp_path, name, total, min_ratio):
    shape = circuits.Shape(name)
    circ = circuits.Circuit(shape, O.OracleBackend, seed=3)
    srs = O.params_setup(shape.k, 42) if shape.k > 11 else O.downsized_srs(shape.k)
    wits = [circ.witness(900 + i) for i in range(8)]
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "proofs.bin")
    write_input(inp, shape, circ, srs, wits)
    exe = build_client(tmp_path)
    res = subprocess.run([exe, inp, "100", str(total), "1", outp], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    print(line)
    assert line["failures"] == 0 and line["identical_to_batched"]
    assert line["coalesced_batches"] < line["coalesced_requests"] / 4, "requests were not coalesced"
    po = O.PlonkOracle(circ.blob, srs, threads=8)
    raw = open(outp, "rb").read()
    inst = np.stack([wits[j % 8][1] for j in range(total)])
    assert po.verify_batch(raw, inst, threads=8) == (True, 0)
    if min_ratio is not None:
        assert line["ratio"] >= min_ratio, "concurrent callers reach only %.0f %% of the batched throughput" % (100 * line["ratio"])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "r02_service_100_callers.json"), "w") as f:
            json.dump(line, f)
