#!/usr/bin/env python3
"""bench.py — batched Shielder-shaped halo2 proofs/sec on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[3]): a batch of 1024 withdraw-shaped proofs per GPU (k = 13, KZG/BN254,
SHPLONK, Keccak transcript), independent seeded witnesses, seeded `ParamsKZG::setup` SRS (the real
ppot_0080_13 file is not in the reference tree).  One "step" = one pass of `zkgpu_prove_batch` over the batch.

  value : proofs/s with the advice columns already resident in HBM (zkgpu_prove_batch_dev)
  e2e   : proofs/s through the reference-facing C-ABI call with HOST (pinned) buffers — host->device copy of
          the advice columns and device->host reads of commitments / evaluations inside the timed region
  roofline : the dominant kernel class, timed live with CUDA events on the library's stream
  cpu_baseline : the CPU oracle prover (restated halo2 create_proof) on all host cores, bounded sample;
          the same sample's GPU proofs are compared byte-for-byte and verified (checker role only)

`--impl reference` times the CPU prover alone.  N > 1: one process per GPU (torchrun), the batch is sharded
by replication of the key material — every rank proves its own 1024 proofs, no data-path collective
("scaling": "weak"); NCCL is used only for the barrier and the max-over-ranks timing.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "Shielder halo2 proofs/sec (batched)"
UNIT = "proofs/s"
SRS_SEED = 42            # SHIELDER_RNG_SEED default (crates/shielder-setup/lib.rs:19)
CIRCUIT_SEED = 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="proofs per GPU per step")
    ap.add_argument("--shape", default="withdraw")
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_prover(shape_name, blob_builder):
    """The CPU oracle prover over the same SRS / circuit (checker + CPU baseline; never on the product path)."""
    import oracle_lib as O
    from zkgpu import circuits
    shape = circuits.Shape(shape_name)
    circ = blob_builder(shape, O.OracleBackend)
    srs = O.params_setup(shape.k, SRS_SEED, threads=os.cpu_count() or 1)
    po = O.PlonkOracle(circ.blob, srs, threads=os.cpu_count() or 1)
    return shape, circ, po


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference prover is Rust with
    un-vendored git dependencies and there is no cargo here, so this is the CPU oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from zkgpu import circuits
    cores = os.cpu_count() or 1
    shape, circ, po = cpu_prover(args.shape, lambda s, be: circuits.Circuit(s, be, seed=CIRCUIT_SEED))
    adv, pi = circ.witness(1)
    t = time.perf_counter(); po.prove(adv, pi, seed=1); first = time.perf_counter() - t
    per_step = max(1, min(8, int(args.cpu_sample_seconds / max(first, 1e-3) / max(args.steps + args.warmup, 1))))
    wits = [circ.witness(100 + i) for i in range(per_step)]
    for _ in range(args.warmup):
        for i, (a, p) in enumerate(wits):
            po.prove(a, p, seed=i + 1)
    t = time.perf_counter()
    for _ in range(args.steps):
        for i, (a, p) in enumerate(wits):
            po.prove(a, p, seed=i + 1)
    dt = time.perf_counter() - t
    value = args.steps * per_step / dt
    sample = "%d %s-shaped proofs (k=%d) per step, %d steps, CPU oracle prover on %d threads" % (per_step, args.shape, shape.k, args.steps, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery)",
        "data": "synthetic", "config": {"workload": "batch of withdraw-shaped proofs (BASELINE configs[3]), bounded CPU sample", "shape": args.shape,
                                         "k": shape.k, "proofs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import zkgpu
    from zkgpu import circuits
    from zkgpu.gpu_backend import GpuBackend

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    zkgpu.init(local)           # raises if libzkgpu.so or a GPU is missing: there is no CPU fallback
    L = zkgpu.lib()
    L.zkgpu_stream.restype = C.c_void_p

    # ---- key material: SRS, circuit, proving key (untimed; the reference loads params.bin / pk.bin once) ----
    shape = circuits.Shape(args.shape)
    g, gl = zkgpu.params_setup(shape.k, SRS_SEED)
    params = zkgpu.ParamsKZG(shape.k, g, gl)
    circ = circuits.Circuit(shape, GpuBackend, seed=CIRCUIT_SEED)
    pk = zkgpu.ProvingKey(params, circ.blob)
    M = args.batch

    # ---- batch of M independent seeded witnesses in pinned host memory ----
    A, n = shape.num_advice, shape.n
    h_adv_t = torch.empty((M, A, n, 4), dtype=torch.int64, pin_memory=True)
    h_adv = h_adv_t.numpy().view(np.uint64)
    inst = np.empty((M, shape.num_pi, 4), dtype=np.uint64)
    base = rank * M
    for i in range(M):
        a, p = circ.witness(1000 + base + i)
        h_adv[i] = a
        inst[i] = p
    seeds = (np.arange(M, dtype=np.uint64) + np.uint64(1 + base))
    d_adv = h_adv_t.cuda()
    proofs = np.zeros(M * pk.proof_len, dtype=np.uint8)
    stream = torch.cuda.ExternalStream(L.zkgpu_stream())

    def step_dev():
        pk.prove_batch_dev(d_adv.data_ptr(), inst, seeds, out=proofs)

    def step_host():
        _chk = zkgpu._chk
        _chk(L.zkgpu_prove_batch(C.c_uint64(pk.handle), C.c_void_p(h_adv_t.data_ptr()), inst.ctypes.data_as(C.c_void_p), C.c_size_t(shape.num_pi),
                                 C.c_size_t(M), seeds.ctypes.data_as(C.c_void_p), proofs.ctypes.data_as(C.c_void_p), C.c_size_t(pk.proof_len)))

    def timed(fn, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_dev()

    # ---- value: inputs resident in HBM --------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = zkgpu.launch_count()
    ms_value = timed(step_dev, args.steps)
    gpu_launches = zkgpu.launch_count() - launches0
    clocks = sampler.summary()
    value = world * M * args.steps / (ms_value / 1e3)
    proofs_value = proofs.copy()

    # ---- per-kernel-class device time: one more step with CUDA events around every launch group on the
    # library's streams.  The library then runs a single pipeline worker, so launches do not share the GPU and
    # the durations are per-kernel (the headline passes overlap two workers).
    L.zkgpu_kernel_timing(1)
    for s in range(8):
        L.zkgpu_kernel_times(s, None, None, 1)
    ms_ktimed = timed(step_dev, 1)
    ktimes = {}
    names = ["msm_bucket_accumulate", "msm_digit_sort", "msm_bucket_reduce", "ntt_tile", "quotient_eval_h", "permutation_product", "poly_algebra"]
    for s, nm in enumerate(names):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        L.zkgpu_kernel_times(s, C.byref(ms), C.byref(cnt), 1)
        ktimes[nm] = (ms.value, cnt.value)
    L.zkgpu_kernel_timing(0)
    ksteps = 1

    # ---- single-proof latency (the metric's second half): m = 1 through the same call, p50 of 64 -----
    lat = []
    for i in range(68):
        t = time.perf_counter()
        pk.prove_batch_dev(d_adv.data_ptr(), inst[:1], seeds[:1], out=proofs[:pk.proof_len])
        lat.append(1e3 * (time.perf_counter() - t))
    lat = sorted(lat[4:])
    p50_ms = lat[len(lat) // 2]
    latency = {"calls": len(lat), "p10_ms": lat[len(lat) // 10], "p50_ms": p50_ms, "p90_ms": lat[(9 * len(lat)) // 10]}

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------
    step_host()
    ms_e2e = timed(step_host, args.steps)
    e2e_value = world * M * args.steps / (ms_e2e / 1e3)
    assert np.array_equal(proofs, proofs_value), "host-buffer and device-resident paths produced different proofs"
    bf = shape.blinding_factors
    h2d = M * (A * n * 32 + shape.num_pi * 32 + (A * (bf + 1) + shape.num_perm_sets * bf) * 64 + 32)
    d2h = M * (64 * (A + shape.num_perm_sets + 1 + shape.num_quotients + 2) + 32 * (shape.num_evals + 1))

    # ---- roofline of the dominant kernel class ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    imad, traffic = {}, {}
    try:
        imad = json.load(open(os.path.join(ROOT, "profiles", "imad_peak.json")))
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")))   # ncu --set full, per launch
    except Exception:
        pass
    tr_of = lambda kname: (traffic[kname]["dram_bytes_read"] + traffic[kname]["dram_bytes_write"]) if kname in traffic else None
    # algorithmic work per step (DESIGN.md "Kernels"): fixed-base MSM = n*W mixed additions of 10 Fq muls;
    # NTT = 64 bytes per point per transform (32 B for zero-padded inputs)
    W_win, c_win = 254 // 13 + 1, 13
    msm_per_proof = shape.num_msm
    fmul_bucket = ksteps * M * msm_per_proof * n * W_win * 10
    # the quotient is evaluated on num_quotients cosets of the size-n subgroup (not on halo2's whole 2^extended_k domain):
    # every extended column = one read of n coefficients + Qc size-n transforms written; h = Qc in-place size-n inverse transforms
    cn = shape.num_quotients * n
    ntt_bytes_proof = shape.num_ntt * 64 * n + (shape.num_ext_ntt - 1) * (32 * n + 32 * cn) + 64 * cn
    ntt_bytes = ksteps * M * ntt_bytes_proof
    total_k_ms = sum(v[0] for v in ktimes.values()) or 1.0
    dom = max(ktimes, key=lambda k_: ktimes[k_][0])
    fmul_peak = imad.get("imad_wide_Gops", 11360.0) / 136.0     # Montgomery product = 136 32x32->64 multiply-adds
    roof_imad = {"kernel": "k_msm_buckets", "bound": "imad", "achieved": fmul_bucket / (ktimes["msm_bucket_accumulate"][0] / 1e3) / 1e9 if ktimes["msm_bucket_accumulate"][0] else None,
                 "peak": fmul_peak, "unit": "Gfieldmul/s", "traffic": tr_of("k_msm_buckets"),
                 "traffic_note": "DRAM bytes of one 1024-MSM launch (ncu --set full, profiles/r01_kernel_traffic.json); algorithmic bytes of that launch: 1.24e9",
                 "peak_source": "IMAD.WIDE issue rate measured with tools/imad_peak.cu on this pool / 136 multiply-adds per 254-bit Montgomery product",
                 "launches": ktimes["msm_bucket_accumulate"][1], "ms_total": ktimes["msm_bucket_accumulate"][0],
                 "share_of_kernel_time": ktimes["msm_bucket_accumulate"][0] / total_k_ms}
    roof_imad["frac"] = roof_imad["achieved"] / roof_imad["peak"] if roof_imad["achieved"] else None
    roof_hbm = {"kernel": "k_ntt_tile", "bound": "hbm", "achieved": ntt_bytes / (ktimes["ntt_tile"][0] / 1e3) / 1e9 if ktimes["ntt_tile"][0] else None,
                "peak": hbm_peak, "unit": "GB/s", "traffic": tr_of("k_ntt_tile"), "peak_source": hbm_src,
                "launches": ktimes["ntt_tile"][1], "ms_total": ktimes["ntt_tile"][0], "share_of_kernel_time": ktimes["ntt_tile"][0] / total_k_ms}
    roof_hbm["frac"] = roof_hbm["achieved"] / roof_hbm["peak"] if roof_hbm["achieved"] else None
    roofline = dict(roof_hbm if dom == "ntt_tile" else roof_imad)
    roofline["dominant_class"] = dom
    # ---- whole-proof bound: field products of the MSMs and transforms at the IMAD ceiling vs transform bytes at HBM peak
    nb = 1 << (c_win - 1)
    k_log = shape.k
    mul_msm = msm_per_proof * (n * W_win * 10 + 2 * nb * 14)
    mul_ntt = shape.num_ntt * (n // 2) * k_log + (shape.num_ext_ntt - 1) * shape.num_quotients * ((n // 2) * k_log + n) + shape.num_quotients * (n // 2) * k_log
    imad_bound = fmul_peak * 1e9 / (mul_msm + mul_ntt)
    hbm_bound = hbm_peak * 1e9 / ntt_bytes_proof
    proof_bound = {"msm_fieldmul_per_proof": mul_msm, "ntt_fieldmul_per_proof": mul_ntt, "ntt_bytes_per_proof": ntt_bytes_proof,
                   "imad_bound_proofs_per_s": imad_bound, "hbm_bound_proofs_per_s": hbm_bound, "bound_proofs_per_s": min(imad_bound, hbm_bound),
                   "achieved_frac_per_gpu": (value / world) / min(imad_bound, hbm_bound),
                   "note": "MSM (bucket additions + bucket reduction) and NTT products only; quotient evaluation, grand products and openings are extra work the bound ignores"}

    # ---- CPU baseline + byte parity of a sample (rank 0, N = 1 only) ----------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        _, ocirc, po = cpu_prover(args.shape, lambda s, be: circuits.Circuit(s, be, seed=CIRCUIT_SEED))
        setup_s = time.perf_counter() - t0
        assert ocirc.blob == circ.blob, "CPU- and GPU-built circuits differ"
        done, t_prove, ok = 0, 0.0, True
        while done < min(M, 16) and (t_prove < args.cpu_sample_seconds or done < 2):
            t = time.perf_counter()
            want = po.prove(h_adv[done], inst[done], seed=int(seeds[done]))
            t_prove += time.perf_counter() - t
            got = proofs_value[done * pk.proof_len:(done + 1) * pk.proof_len].tobytes()
            ok = ok and got == want and po.verify(got, inst[done])
            done += 1
        assert ok, "GPU proofs differ from the CPU prover or fail verification"
        # BASELINE configs[3]: the WHOLE batch is checked by the restated halo2-verifier (random linear combination of the
        # per-proof pairing inputs, one pairing product)
        t = time.perf_counter()
        all_ok, malformed = po.verify_batch(proofs_value.tobytes(), inst, threads=cores)
        t_verify = time.perf_counter() - t
        assert all_ok and malformed == 0, "batch verification of the GPU proofs failed"
        cpu_baseline = {"value": done / t_prove, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "first %d proofs of the batch, CPU oracle prover (restated halo2 create_proof) on %d threads, %.1f s; "
                                  "GPU proofs byte-identical and accepted by the verifier restatement" % (done, cores, t_prove),
                        "keygen_and_srs_seconds": setup_s,
                        "batch_verified": "all %d proofs of the timed batch accepted by the verifier restatement (batched pairing check, %.1f s)" % (M, t_verify)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (254-bit Montgomery)", "data": "synthetic",
            "config": {"workload": "batch of %d withdraw-shaped proofs per GPU (BASELINE configs[3])" % M, "shape": args.shape, "k": shape.k,
                       "extended_k": shape.extended_k, "advice_columns": A, "msm_per_proof": shape.num_msm, "ntt_per_proof": shape.num_ntt,
                       "ext_ntt_per_proof": shape.num_ext_ntt, "quotient_cosets": shape.num_quotients, "proof_bytes": pk.proof_len, "sub_batch": pk.sub_batch, "msm_window_bits": c_win,
                       "srs": "ParamsKZG::setup seed 42 (ppot_0080_13 absent from the reference tree)",
                       "l2": "inputs larger than L2 (%.1f GB of advice per step)" % (M * A * n * 32 / 1e9), "parallelism": "dp%d, no collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "roofline_hbm": roof_hbm, "roofline_imad": roof_imad,
            "proof_bound": proof_bound, "kernel_ms_per_step": {k_: round(v[0], 3) for k_, v in ktimes.items()}, "kernel_timed_step_ms": ms_ktimed,
            "single_proof_p50_ms": p50_ms, "single_proof_latency": latency, "cpu_baseline": cpu_baseline,
        }), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
