#!/usr/bin/env python3
"""bench.py — batched Shielder-shaped halo2 proofs/sec on B200 (BASELINE.json metric).

Default workload (BASELINE.json configs[3]): a batch of 1024 withdraw-shaped proofs per GPU (k = 13, KZG/BN254,
SHPLONK, Keccak transcript), independent seeded witnesses, seeded `ParamsKZG::setup` SRS (the real
ppot_0080_13 file is not in the reference tree).  One "step" = one pass of `zkgpu_prove_batch` over the batch.

  value : proofs/s with the advice columns already resident in HBM (zkgpu_prove_batch_dev)
  e2e   : proofs/s through the reference-facing C-ABI call with HOST (pinned) buffers — host->device copy of
          the advice columns and device->host reads of commitments / evaluations inside the timed region
  roofline : the dominant kernel class, timed live with CUDA events on the library's stream
  cpu_baseline : the CPU oracle prover (restated halo2 create_proof) on the host cores, bounded sample;
          the same sample's GPU proofs are compared byte-for-byte and verified (checker role only)

The same JSON line also carries three more measurements so the driver's records cover them (each one can be run alone
with --workload, in which case it IS the line's `value` / `e2e`):
  withdraw_lookup : the same batch with two lookup arguments in the circuit (range-check style), value + e2e + kernel classes
  mixed_stream    : BASELINE configs[4] — 4096 new_account / deposit / withdraw requests (1:2:2, seed 7), request i served by
                    GPU i mod N, the three circuits proved concurrently, end to end from pinned host buffers
  msm24           : BASELINE configs[1] at its largest size — one 2^24-point G1 MSM with the points split across the N GPUs,
                    bases resident per shard, partial points summed (peer copies inside one process, a 64-byte all-gather
                    between processes), checked against [p(s)] G

`--impl reference` times the CPU prover alone.  N > 1: one process per GPU (torchrun), the batch is sharded
by replication of the key material — every rank proves its own 1024 proofs, no data-path collective
("scaling": "weak"); NCCL is used only for the barrier and the max-over-ranks timing.  `--single-process` drives the
N GPUs from ONE process instead (zkgpu_init(device_mask); the library shards the batch across its devices).
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
MSM_SEED = 7
DTYPE = "u32x8 (254-bit Montgomery)"
KT_NAMES = ["msm_bucket_accumulate", "msm_digit_sort", "msm_bucket_reduce", "ntt_tile", "quotient_eval_h", "permutation_product",
            "poly_algebra", "lookup_permute", "misc_blind_chacha_normalize", "host_fiat_shamir_gap"]
MIX_TYPES = ["new_account", "deposit", "withdraw"]
MIX_WEIGHTS = [1, 2, 2]
STRICT = False           # --strict: a timing-class coverage below 0.97 is an error instead of a field of the line


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="proofs", choices=["proofs", "mixed", "msm24"])
    ap.add_argument("--batch", type=int, default=1024, help="proofs per GPU per step")
    ap.add_argument("--shape", default="withdraw")
    ap.add_argument("--requests", type=int, default=4096, help="mixed stream: requests per step (whole job)")
    ap.add_argument("--msm-log-n", type=int, default=24)
    ap.add_argument("--single-process", action="store_true", help="drive --gpus devices from this one process (no torchrun)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="default workload only: skip withdraw_lookup / mixed_stream / msm24")
    ap.add_argument("--strict", action="store_true", help="fail if the kernel-class timers cover less than 97 %% of the timed step")
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle prover (checker + CPU baseline; never on the product path)
# ---------------------------------------------------------------------------------------------------------------------
def physical_cores():
    """one logical CPU per physical core (lscpu), so the CPU prover's threads do not share cores"""
    try:
        out = subprocess.run(["lscpu", "-p=CPU,CORE,SOCKET"], capture_output=True, text=True, timeout=5).stdout
        seen, cpus = set(), []
        for ln in out.splitlines():
            if ln.startswith("#") or not ln.strip():
                continue
            cpu, core, sock = ln.split(",")[:3]
            if (core, sock) not in seen:
                seen.add((core, sock))
                cpus.append(int(cpu))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        return cpus or sorted(allowed)
    except Exception:
        return sorted(os.sched_getaffinity(0))


def cpu_prover(shape_name, threads):
    """The CPU oracle prover over the same SRS / circuit, its threads pinned to distinct physical cores."""
    import oracle_lib as O
    from zkgpu import circuits
    shape = circuits.Shape(shape_name)
    circ = circuits.Circuit(shape, O.OracleBackend, seed=CIRCUIT_SEED)
    srs = O.params_setup(shape.k, SRS_SEED, threads=threads)
    po = O.PlonkOracle(circ.blob, srs, threads=threads)
    return shape, circ, po


def cpu_phase_times(shape_name, threads):
    """per-phase milliseconds of ONE CPU proof (the oracle's ORACLE_TIMING trace, printed to stderr by the C++ code): run in a
    child process so the trace can be captured"""
    code = ("import sys; sys.path[:0] = %r\n"
            "import bench\n"
            "shape, circ, po = bench.cpu_prover(%r, %d)\n"
            "adv, pi = circ.witness(1)\n"
            "po.prove(adv, pi, seed=1)\n" % ([ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")], shape_name, threads))
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, ORACLE_TIMING="1"))
        phases = {}
        for ln in out.stderr.splitlines():
            if ln.startswith("[oracle]") and ln.rstrip().endswith("ms"):
                name, ms = ln[len("[oracle]"):].rsplit(None, 2)[0].strip(), float(ln.split()[-2])
                phases[name] = round(phases.get(name, 0.0) + ms, 1)
        return phases
    except Exception as e:      # the phase table is context, never a reason to lose the benchmark line
        return {"error": str(e)[:200]}


def cpu_best_threads(shape_name, candidates):
    """the thread count at which the CPU prover is fastest on this host (round 1: 16 threads beat 32)"""
    from zkgpu import circuits  # noqa: F401
    best = None
    for t in candidates:
        shape, circ, po = cpu_prover(shape_name, t)
        adv, pi = circ.witness(1)
        po.prove(adv, pi, seed=1)
        t0 = time.perf_counter(); po.prove(adv, pi, seed=2); dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (t, dt, shape, circ, po)
    return best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference prover is Rust with
    un-vendored git dependencies and there is no cargo here, so this is the CPU oracle port, on the host cores at the
    thread count where it runs fastest, threads pinned to physical cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    phys = physical_cores()
    os.sched_setaffinity(0, set(phys))
    logical = os.cpu_count() or 1
    cands = sorted({len(phys), max(1, len(phys) // 2)}, reverse=True)
    threads, first, shape, circ, po = cpu_best_threads(args.shape, cands)
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
    sample = "%d %s-shaped proofs (k=%d) per step, %d steps, CPU oracle prover on %d threads pinned to %d physical cores (%d logical CPUs)" % (
        per_step, args.shape, shape.k, args.steps, threads, len(phys), logical)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": {"workload": "batch of withdraw-shaped proofs (BASELINE configs[3]), bounded CPU sample", "shape": args.shape,
                                         "k": shape.k, "proofs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "thread_counts_tried": cands, "phase_ms_one_proof": cpu_phase_times(args.shape, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
class Env:
    """process / device set-up shared by the workloads"""

    def __init__(self, args):
        import torch
        import zkgpu
        self.torch, self.zkgpu, self.args = torch, zkgpu, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.devices = 1                      # devices driven by THIS process
        if args.single_process:
            assert self.world == 1, "--single-process is not launched under torchrun"
            self.devices = max(1, args.gpus)
            self.local = 0
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            dist.barrier()                    # NCCL builds its communicator on the first collective: not inside a timed region
            torch.cuda.synchronize()
        # raises if libzkgpu.so or a GPU is missing: there is no CPU fallback
        zkgpu.init(mask=(1 << self.devices) - 1) if args.single_process else zkgpu.init(self.local)
        self.L = zkgpu.lib()
        self.L.zkgpu_stream.restype = C.c_void_p
        self.stream = torch.cuda.ExternalStream(self.L.zkgpu_stream())
        self.total_gpus = self.world * self.devices
        self._params = {}

    def params(self, k):
        """ParamsKZG::setup seed 42 at k = 13, downsized for smaller circuits (ParamsKZG::downsize)"""
        z = self.zkgpu
        if 13 not in self._params:
            g13, gl13 = z.params_setup(13, SRS_SEED)
            self._params[13] = z.ParamsKZG(13, g13, gl13)
            self._g13 = g13
        if k not in self._params:
            g = self._g13[: 1 << k].copy()
            self._params[k] = z.ParamsKZG(k, g, z.g_to_lagrange(g, k))
        return self._params[k]

    def timed(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize, CUDA events on the library's stream, max over ranks"""
        torch, dist = self.torch, self.dist
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), 0.0)
        if self.devices > 1:      # several devices in one process: the call returns when ALL of them are done; events see only device 0
            ms = max(ms, 1e3 * (time.perf_counter() - t0))
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def ktimes(self, fn):
        """one more call with CUDA events around every launch group of a kernel class (the library then runs one pipeline
        worker, so launches do not share the GPU and the durations are per kernel)"""
        L = self.L
        L.zkgpu_kernel_timing(1)
        for s in range(len(KT_NAMES)):
            L.zkgpu_kernel_times(s, None, None, 1)
        L.zkgpu_msm_additions(None, 1)
        ms_step = self.timed(fn, 1)
        adds = C.c_uint64(0)
        L.zkgpu_msm_additions(C.byref(adds), 1)
        self.last_msm_additions = adds.value
        out = {}
        for s, nm in enumerate(KT_NAMES):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            L.zkgpu_kernel_times(s, C.byref(ms), C.byref(cnt), 1)
            out[nm] = (ms.value, cnt.value)
        L.zkgpu_kernel_timing(0)
        return out, ms_step


def peaks():
    pk = {}
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = (pk["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in pk else (6650.0, "fallback (B200_PROFILING.md)")
    imad = {}
    try:
        imad = json.load(open(os.path.join(ROOT, "profiles", "imad_peak.json")))
    except Exception:
        pass
    return hbm[0], hbm[1], imad.get("imad_wide_Gops", 11360.0) / 136.0     # Montgomery product = 136 32x32->64 multiply-adds


def kernel_traffic():
    for name in ("r02_kernel_traffic.json", "r01_kernel_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))), name
        except Exception:
            pass
    return {}, None


def run_proofs(env, shape_name, M, steps, warmup, full):
    """1024 (M) proofs of one circuit per step and device.  full: the headline pass (latency, roofline, CPU check)."""
    from zkgpu import circuits
    from zkgpu.gpu_backend import GpuBackend
    torch, zkgpu, L = env.torch, env.zkgpu, env.L
    shape = circuits.Shape(shape_name)
    params = env.params(shape.k)
    circ = circuits.Circuit(shape, GpuBackend, seed=CIRCUIT_SEED)
    pk = zkgpu.ProvingKey(params, circ.blob)
    A, n = shape.num_advice, shape.n
    Mtot = M * env.devices                      # proofs this PROCESS proves per step
    distinct = M if full else min(M, 64)        # the light passes cycle 64 witnesses (seeds stay distinct)
    h_adv_t = torch.empty((Mtot, A, n, 4), dtype=torch.int64, pin_memory=True)
    h_adv = h_adv_t.numpy().view(np.uint64)
    inst = np.empty((Mtot, shape.num_pi, 4), dtype=np.uint64)
    base = env.rank * Mtot
    for i in range(distinct):
        a, p = circ.witness(1000 + base + i)
        h_adv[i], inst[i] = a, p
    for i in range(distinct, Mtot):
        h_adv[i], inst[i] = h_adv[i % distinct], inst[i % distinct]
    seeds = (np.arange(Mtot, dtype=np.uint64) + np.uint64(1 + base))
    proofs = np.zeros(Mtot * pk.proof_len, dtype=np.uint8)
    d_adv = h_adv_t[:M].cuda() if env.devices == 1 else None

    def step_dev():
        pk.prove_batch_dev(d_adv.data_ptr(), inst, seeds, out=proofs)

    def step_host():
        zkgpu._chk(L.zkgpu_prove_batch(C.c_uint64(pk.handle), C.c_void_p(h_adv_t.data_ptr()), inst.ctypes.data_as(C.c_void_p), C.c_size_t(shape.num_pi),
                                       C.c_size_t(Mtot), seeds.ctypes.data_as(C.c_void_p), proofs.ctypes.data_as(C.c_void_p), C.c_size_t(pk.proof_len)))

    res = {"shape": shape, "pk": pk, "circ": circ, "h_adv": h_adv, "inst": inst, "seeds": seeds, "M": M, "Mtot": Mtot}
    total = env.world * Mtot
    if d_adv is not None:
        for _ in range(max(warmup, 3) if full else 1):
            step_dev()
        sampler = ClockSampler(env.local)
        sampler.start()
        launches0 = zkgpu.launch_count()
        ms_value = env.timed(step_dev, steps)
        res["gpu_launches"] = zkgpu.launch_count() - launches0
        res["clocks"] = sampler.summary()
        res["value"], res["ms_value"] = total * steps / (ms_value / 1e3), ms_value
        res["proofs_value"] = proofs.copy()
        res["ktimes"], res["ms_ktimed"] = env.ktimes(step_dev)
        res["msm_additions"] = env.last_msm_additions
        if full:   # single-proof latency (the metric's second half): m = 1 through the same call, p50 of 64
            lat = []
            for i in range(68):
                t = time.perf_counter()
                pk.prove_batch_dev(d_adv.data_ptr(), inst[:1], seeds[:1], out=proofs[:pk.proof_len])
                lat.append(1e3 * (time.perf_counter() - t))
            lat = sorted(lat[4:])
            res["latency"] = {"calls": len(lat), "p10_ms": lat[len(lat) // 10], "p50_ms": lat[len(lat) // 2], "p90_ms": lat[(9 * len(lat)) // 10]}
    # e2e: host buffers through the C ABI
    step_host()
    if d_adv is None:
        for _ in range(max(warmup, 3) - 1):
            step_host()
        sampler = ClockSampler(env.local)
        sampler.start()
        launches0 = zkgpu.launch_count()
    ms_e2e = env.timed(step_host, steps)
    if d_adv is None:
        res["gpu_launches"] = zkgpu.launch_count() - launches0
        res["clocks"] = sampler.summary()
        res["value"], res["ms_value"], res["proofs_value"] = None, None, proofs.copy()
        res["ktimes"], res["ms_ktimed"] = env.ktimes(step_host)
        res["msm_additions"] = env.last_msm_additions
    else:
        assert np.array_equal(proofs, res["proofs_value"]), "host-buffer and device-resident paths produced different proofs"
    res["e2e"], res["ms_e2e"] = total * steps / (ms_e2e / 1e3), ms_e2e
    bf = shape.blinding_factors
    res["h2d"] = Mtot * (A * n * 32 + shape.num_pi * 32 + (A * (bf + 1) + shape.num_perm_sets * bf) * 64 + 32)
    res["d2h"] = Mtot * (64 * (A + shape.num_perm_sets + 1 + shape.num_quotients + 2 + 3 * shape.n_lookup) + 32 * (shape.num_evals + 1))
    return res


def proofs_rooflines(env, r, steps):
    """roofline of the dominant kernel class + the whole-proof bound (DESIGN.md 'Kernels')"""
    shape, M, kt = r["shape"], r["Mtot"], r["ktimes"]
    hbm_peak, hbm_src, fmul_peak = peaks()
    traffic, traffic_file = kernel_traffic()
    tr_of = lambda kname: (traffic[kname]["dram_bytes_read"] + traffic[kname]["dram_bytes_write"]) if kname in traffic else None
    n = shape.n
    # algorithmic work of ONE timed step (the class timers cover one step): fixed-base MSM = n*W mixed additions of 10 Fq products;
    # NTT = 64 bytes per point per in-place transform, 32 B in + 32 B out per coset of an extension
    c_win = 13 if shape.k >= 12 else 12
    W_win = 254 // c_win + 1
    nominal_adds = M * shape.num_msm * n * W_win
    # additions actually queued for the bucket kernels in the class-timed step (zero digits are skipped; a column that is constant
    # over stretches is committed through its differences): the work the achieved rate is computed from
    adds = r.get("msm_additions") or nominal_adds
    fmul_bucket = adds * 10
    cn = shape.num_quotients * n
    ntt_bytes_proof = shape.num_ntt * 64 * n + (shape.num_ext_ntt - 1) * (32 * n + 32 * cn) + 64 * cn
    total_k_ms = sum(v[0] for k_, v in kt.items() if k_ != "host_fiat_shamir_gap") or 1.0
    dom = max((k_ for k_ in kt if k_ != "host_fiat_shamir_gap"), key=lambda k_: kt[k_][0])
    b_ms, n_ms = kt["msm_bucket_accumulate"], kt["ntt_tile"]
    roof_imad = {"kernel": "k_msm_buckets", "bound": "imad", "achieved": fmul_bucket / (b_ms[0] / 1e3) / 1e9 if b_ms[0] else None,
                 "peak": fmul_peak, "unit": "Gfieldmul/s", "traffic": tr_of("k_msm_buckets"),
                 "traffic_note": "DRAM bytes of one 1024-MSM launch (ncu --set full, profiles/%s); algorithmic bytes of that launch: 1.24e9" % traffic_file,
                 "peak_source": "IMAD.WIDE issue rate measured with tools/imad_peak.cu on this pool / 136 multiply-adds per 254-bit Montgomery product",
                 "algorithmic_note": "mixed additions counted by the digit sort (zkgpu_msm_additions) x 10 products (8M + 2S); the two products of Y3 "
                                     "share one reduction (fe_mul_add2), so the kernel issues fewer multiply-adds than this count assumes",
                 "additions": adds, "additions_if_every_digit_were_nonzero": nominal_adds,
                 "launches": b_ms[1], "ms_total": b_ms[0], "share_of_kernel_time": b_ms[0] / total_k_ms}
    roof_imad["frac"] = roof_imad["achieved"] / roof_imad["peak"] if roof_imad["achieved"] else None
    roof_hbm = {"kernel": "k_ntt_cluster8 (k_ntt_tile below 592 transforms per call)", "bound": "hbm", "achieved": M * ntt_bytes_proof / (n_ms[0] / 1e3) / 1e9 if n_ms[0] else None,
                "peak": hbm_peak, "unit": "GB/s", "traffic": tr_of("k_ntt_cluster8") or tr_of("k_ntt_tile"), "peak_source": hbm_src,
                "launches": n_ms[1], "ms_total": n_ms[0], "share_of_kernel_time": n_ms[0] / total_k_ms}
    roof_hbm["frac"] = roof_hbm["achieved"] / roof_hbm["peak"] if roof_hbm["achieved"] else None
    roofline = dict(roof_hbm if dom == "ntt_tile" else roof_imad)
    roofline["dominant_class"] = dom
    nb = 1 << (c_win - 1)
    mul_msm = adds * 10 // M + shape.num_msm * 2 * nb * 14     # counted additions of this witness set + the bucket reductions
    mul_ntt = shape.num_ntt * (n // 2) * shape.k + (shape.num_ext_ntt - 1) * shape.num_quotients * ((n // 2) * shape.k + n) + shape.num_quotients * (n // 2) * shape.k
    imad_bound = fmul_peak * 1e9 / (mul_msm + mul_ntt)
    hbm_bound = hbm_peak * 1e9 / ntt_bytes_proof
    per_gpu = (r["value"] if r["value"] else r["e2e"]) / env.total_gpus
    bound = {"msm_fieldmul_per_proof": mul_msm, "ntt_fieldmul_per_proof": mul_ntt, "ntt_bytes_per_proof": ntt_bytes_proof,
             "imad_bound_proofs_per_s": imad_bound, "hbm_bound_proofs_per_s": hbm_bound, "bound_proofs_per_s": min(imad_bound, hbm_bound),
             "achieved_frac_per_gpu": per_gpu / min(imad_bound, hbm_bound),
             "note": "MSM (the bucket additions this witness set needs + bucket reduction) and NTT products only; quotient evaluation, grand products "
                     "and openings are extra work the bound ignores"}
    return roofline, roof_hbm, roof_imad, bound


def timing_coverage(kt, ms_ktimed, devices=1):
    """share of the single-worker timed step accounted for: kernel classes + the device idle time at the Fiat-Shamir round trips
    (`host_fiat_shamir_gap`: with ONE pipeline worker the GPU waits while the host hashes the transcript; the headline passes hide it
    behind the other workers' kernels).  What is left is launch gaps and the small host->device uploads."""
    s = sum(v[0] for v in kt.values()) / devices      # the class timers add up over the devices of a single-process run
    k = s - kt.get("host_fiat_shamir_gap", (0.0, 0))[0] / devices
    cov = s / ms_ktimed if ms_ktimed else None
    ok = cov is None or cov >= 0.97 or devices > 1     # several devices in one process: the wall time also holds their imbalance
    if STRICT:
        assert ok, "kernel-class timers cover only %.1f %% of the timed step: a kernel is missing from the classes" % (100 * cov)
    elif not ok:
        print("WARNING: kernel-class timers cover only %.1f %% of the timed step" % (100 * cov), file=sys.stderr)
    return {"sum_of_kernel_classes_ms": round(k, 3), "host_gap_ms": round(s - k, 3), "timed_step_ms": round(ms_ktimed, 3),
            "covered": round(cov, 4) if cov is not None else None, "untimed_share": round(1 - cov, 4) if cov is not None else None,
            "at_least_0.97": bool(ok)}


def request_stream(n, seed=7):
    rng = np.random.default_rng(seed)
    return rng.choice(len(MIX_TYPES), size=n, p=np.array(MIX_WEIGHTS) / sum(MIX_WEIGHTS))


def assign_requests(stream, world, policy=None):
    """Which GPU serves which request.  SURVEY 8e: "proof i -> GPU (i mod G) (or work-stealing queue for the mixed stream)".  The three
    circuits cost differently (advice columns x rows), so plain round-robin leaves the ranks 2-3 % apart at 8 GPUs; the default here is
    the deterministic form of a work queue: every request, in arrival order, goes to the GPU with the least work assigned so far (ties
    to the lowest index), computed identically by every rank.  ZKGPU_MIXED_POLICY=round_robin selects i mod G."""
    from zkgpu import circuits
    policy = policy or os.environ.get("ZKGPU_MIXED_POLICY", "least_loaded")
    if policy == "round_robin":
        return [list(range(r, len(stream), world)) for r in range(world)]
    cost = []
    for t in MIX_TYPES:
        sh = circuits.Shape(t)
        cost.append(sh.num_msm * sh.n)
    load, out = [0] * world, [[] for _ in range(world)]
    for i, t in enumerate(stream):
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += cost[int(t)]
    return out


def run_mixed(env, requests, steps, warmup, distinct=32):
    """BASELINE configs[4]: `requests` requests of three circuits, request i served by GPU i mod N (one process per GPU) or
    sharded by the library (single process), the three circuits proved CONCURRENTLY (one host thread each: the library has no
    process-wide lock), end to end from pinned host buffers."""
    from zkgpu import circuits, multi
    from zkgpu.gpu_backend import GpuBackend
    torch, zkgpu = env.torch, env.zkgpu
    stream = request_stream(requests)
    mine = assign_requests(stream, env.world)[env.rank]
    jobs = {}
    for ti, t in enumerate(MIX_TYPES):
        idx = [i for i in mine if stream[i] == ti]
        shape = circuits.Shape(t)
        circ = circuits.Circuit(shape, GpuBackend, seed=CIRCUIT_SEED)
        pk = zkgpu.ProvingKey(env.params(shape.k), circ.blob)
        if not idx:
            continue
        wits = [circ.witness(500 + i) for i in range(min(distinct, len(idx)))]
        # request payloads staged in pinned host memory, as a serving host would hold them
        pinned = torch.empty((len(idx), shape.num_advice, shape.n, 4), dtype=torch.int64, pin_memory=True)
        adv = pinned.numpy().view(np.uint64)
        for j in range(len(idx)):
            adv[j] = wits[j % len(wits)][0]
        inst = np.stack([wits[j % len(wits)][1] for j in range(len(idx))])
        # production rng mode: 32 bytes of entropy per request, expanded by ChaCha20 (drawn here from a seeded generator: synthetic)
        seeds = np.random.default_rng(1000 + env.rank * 7 + ti).integers(0, 256, (len(idx), 32), dtype=np.uint8)
        jobs[t] = dict(shape=shape, circ=circ, pk=pk, pinned=pinned, adv=adv, inst=inst, seeds=seeds, idx=idx,
                       out=np.zeros(len(idx) * pk.proof_len, dtype=np.uint8), status=np.zeros(len(idx), dtype=np.int32))
    L = env.L

    def prove_type(j):
        pk = j["pk"]
        zkgpu._chk(L.zkgpu_prove_batch_rng(C.c_uint64(pk.handle), C.c_void_p(j["pinned"].data_ptr()), j["inst"].ctypes.data_as(C.c_void_p),
                                           C.c_size_t(j["shape"].num_pi), C.c_size_t(len(j["idx"])), 2, j["seeds"].ctypes.data_as(C.c_void_p),
                                           j["out"].ctypes.data_as(C.c_void_p), C.c_size_t(pk.proof_len), j["status"].ctypes.data_as(C.c_void_p)))

    def step():
        ths = [threading.Thread(target=prove_type, args=(j,)) for j in jobs.values()]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    for _ in range(max(1, warmup)):
        step()
    ms = env.timed(step, steps)
    for j in jobs.values():
        assert not j["status"].any(), "a proof of the mixed stream failed"
    bytes_in = sum(j["adv"].nbytes + j["inst"].nbytes + j["seeds"].nbytes for j in jobs.values())
    bytes_out = sum(j["out"].nbytes for j in jobs.values())
    return {"jobs": jobs, "value": requests * steps / (ms / 1e3), "ms_per_step": ms / steps, "requests": requests,
            "mix": {t: int((stream == i).sum()) for i, t in enumerate(MIX_TYPES)}, "h2d": bytes_in, "d2h": bytes_out,
            "per_rank": {t: len(j["idx"]) for t, j in jobs.items()}}


def fr_pow(F, base, e):
    """base^e for one field element (Montgomery limbs) with the library's vectorised multiply"""
    acc, b = F.const(1)[None], base[None]
    while e:
        if e & 1:
            acc = F.mul(acc, b)
        b = F.mul(b, b)
        e >>= 1
    return acc[0]


def run_msm(env, log_n, steps, warmup):
    """one 2^log_n-point G1 MSM, points split across the GPUs (a rank / device only holds its shard), bases resident"""
    from zkgpu import multi
    from zkgpu.gpu_backend import GpuBackend as F
    torch, zkgpu, dist = env.torch, env.zkgpu, env.dist
    n = 1 << log_n
    lo, hi = multi.shard_bounds(n, env.rank, env.world)
    cnt = hi - lo
    g = np.empty((cnt, 8), dtype=np.uint64)
    for off in range(0, cnt, 1 << 22):      # setup SRS slice g[i] = s^i G, i in [lo, hi)
        c = min(1 << 22, cnt - off)
        g[off:off + c] = zkgpu.setup_powers(MSM_SEED, lo + off, c)
    bases = zkgpu.Bases(g)
    del g
    sc_t = torch.empty((cnt, 4), dtype=torch.int64, pin_memory=True)
    sc = sc_t.numpy().view(np.uint64)
    sc[:] = F.random(11 + env.rank, cnt)
    dev = torch.device("cuda", env.local)

    def combine(part):
        return multi.combine_partials(multi.gather_points(part, dist, dev)) if dist is not None else part

    out = {}

    def step_e2e():
        out["e2e"] = combine(bases.msm(sc))

    def step_dev():
        out["dev"] = combine(bases.msm(None))

    for _ in range(max(1, warmup)):
        step_e2e()
    ms_e2e = env.timed(step_e2e, steps)
    step_dev()
    ms_dev = env.timed(step_dev, steps)
    kernel_ms = bases.kernel_ms
    # expected value by an independent route: sum_i c_i g[i] = [p(s)] G with p(s) = sum_r s^lo_r * p_r(s) (Horner on the GPU)
    s = F.random(MSM_SEED, 1)[0]
    mine = F.mul(zkgpu.eval_polynomial(sc, s)[None], fr_pow(F, s, lo)[None])[0]
    if dist is not None:
        t = torch.from_numpy(mine.view(np.int64).copy()).to(dev)
        parts = [torch.empty_like(t) for _ in range(env.world)]
        dist.all_gather(parts, t)
        ps = parts[0].cpu().numpy().view(np.uint64)
        for q in parts[1:]:
            ps = F.add(ps[None], q.cpu().numpy().view(np.uint64)[None])[0]
    else:
        ps = mine
    gen = zkgpu.setup_powers(MSM_SEED, 0, 1)           # s^0 * G = the generator (1, 2)
    want = zkgpu.best_multiexp(ps[None], gen)
    ok = bool(np.array_equal(out["e2e"], want) and np.array_equal(out["dev"], want))
    assert ok, "sharded MSM result differs from [p(s)] G"
    # algorithmic products of plain Pippenger on one shard (the window the library picks for that size)
    best = None
    for c in range(2, 17):
        W = 254 // c + 1
        cost = W * (cnt * 10 + (1 << (c - 1)) * 28)
        if best is None or cost < best[0]:
            best = (cost, c, W)
    _, _, fmul_peak = peaks()
    fmul = best[0] + 254 * 9
    bases.release()
    return {"log_n": log_n, "points_per_gpu": cnt, "ms_per_msm_e2e": ms_e2e / steps, "ms_per_msm_resident": ms_dev / steps, "kernel_ms": kernel_ms,
            "h2d_ms_estimate": max(ms_e2e - ms_dev, 0.0) / steps, "window_bits": best[1], "windows": best[2],
            "value": n * steps / (ms_dev / 1e3), "e2e": n * steps / (ms_e2e / 1e3), "h2d": cnt * 32, "d2h": 64,
            "roofline": {"kernel": "k_msm_buckets (plain mode)", "bound": "imad", "achieved": fmul / (kernel_ms / 1e3) / 1e9 if kernel_ms else None,
                         "peak": fmul_peak, "unit": "Gfieldmul/s", "frac": (fmul / (kernel_ms / 1e3) / 1e9 / fmul_peak) if kernel_ms else None,
                         "note": "all MSM kernels of one shard (digit sort, buckets, reduction, window Horner) against the algorithmic products"},
            "checked": "result == [p(s)] G (Horner evaluation of the scalars on the GPU + one scalar multiplication)"}


def cpu_check_proofs(env, r, args):
    """CPU baseline + byte parity of a sample + batch verification (rank 0, N = 1 only; checker role)"""
    phys = physical_cores()
    os.sched_setaffinity(0, set(phys))
    threads = len(phys)
    shape, M = r["shape"], r["Mtot"]
    t0 = time.perf_counter()
    _, ocirc, po = cpu_prover(shape.name, threads)
    setup_s = time.perf_counter() - t0
    assert ocirc.blob == r["circ"].blob, "CPU- and GPU-built circuits differ"
    pl = r["pk"].proof_len
    done, t_prove, ok = 0, 0.0, True
    while done < min(M, 16) and (t_prove < args.cpu_sample_seconds or done < 2):
        t = time.perf_counter()
        want = po.prove(r["h_adv"][done], r["inst"][done], seed=int(r["seeds"][done]))
        t_prove += time.perf_counter() - t
        got = r["proofs_value"][done * pl:(done + 1) * pl].tobytes()
        ok = ok and got == want and po.verify(got, r["inst"][done])
        done += 1
    assert ok, "GPU proofs differ from the CPU prover or fail verification"
    # BASELINE configs[3]: the WHOLE batch is checked by the restated halo2-verifier (random linear combination of the
    # per-proof pairing inputs, one pairing product)
    t = time.perf_counter()
    all_ok, malformed = po.verify_batch(r["proofs_value"].tobytes(), r["inst"], threads=threads)
    t_verify = time.perf_counter() - t
    assert all_ok and malformed == 0, "batch verification of the GPU proofs failed"
    os.sched_setaffinity(0, set(range(os.cpu_count() or 1)))
    return {"value": done / t_prove, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d proofs of the batch, CPU oracle prover (restated halo2 create_proof) on %d threads pinned to physical cores, %.1f s; "
                      "GPU proofs byte-identical and accepted by the verifier restatement" % (done, threads, t_prove),
            "keygen_and_srs_seconds": setup_s,
            "batch_verified": "all %d proofs of the timed batch accepted by the verifier restatement (batched pairing check, %.1f s)" % (M, t_verify)}


def cpu_check_mixed(env, mx):
    """every proof of the mixed stream through the verifier restatement (rank 0, N = 1 only; checker role)"""
    import oracle_lib as O
    threads = len(physical_cores())
    srs13 = O.params_setup(13, SRS_SEED, threads=threads)
    n_ok = 0
    for t, j in mx["jobs"].items():
        srs = srs13 if j["shape"].k == 13 else O.downsized_srs(j["shape"].k, srs13)
        po = O.PlonkOracle(j["circ"].blob, srs, threads=threads)
        ok, bad = po.verify_batch(j["out"].tobytes(), j["inst"], threads=threads)
        assert ok and bad == 0, "mixed stream: %s proofs rejected" % t
        n_ok += len(j["idx"])
    return "all %d proofs of the stream accepted by the verifier restatement (one batched pairing check per circuit)" % n_ok


def main():
    global STRICT
    args = parse()
    STRICT = args.strict
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args)
    N, par = env.total_gpus, ("one process, %d devices" % env.devices if args.single_process else "dp%d, no collective" % env.world)
    cpu_ok = env.rank == 0 and N == 1 and not args.no_cpu_baseline
    common = {"n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic"}
    line = None

    if args.workload == "proofs":
        r = run_proofs(env, args.shape, args.batch, args.steps, args.warmup, full=True)
        shape, pk = r["shape"], r["pk"]
        roofline, roof_hbm, roof_imad, bound = proofs_rooflines(env, r, args.steps)
        head = r["value"] if r["value"] is not None else r["e2e"]
        line = dict(common, metric=METRIC, value=head, unit=UNIT, ms_per_step=(r["ms_value"] or r["ms_e2e"]) / args.steps, scaling="weak",
                    config={"workload": "batch of %d withdraw-shaped proofs per GPU (BASELINE configs[3])" % args.batch, "shape": args.shape, "k": shape.k,
                            "extended_k": shape.extended_k, "advice_columns": shape.num_advice, "msm_per_proof": shape.num_msm, "ntt_per_proof": shape.num_ntt,
                            "ext_ntt_per_proof": shape.num_ext_ntt, "quotient_cosets": shape.num_quotients, "proof_bytes": pk.proof_len, "sub_batch": pk.sub_batch,
                            "msm_window_bits": 13, "srs": "ParamsKZG::setup seed 42 (ppot_0080_13 absent from the reference tree)",
                            "l2": "inputs larger than L2 (%.1f GB of advice per step and GPU)" % (args.batch * shape.num_advice * shape.n * 32 / 1e9),
                            "parallelism": par},
                    e2e={"value": r["e2e"], "unit": UNIT, "ms_per_step": r["ms_e2e"] / args.steps, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                    gpu_launches=r["gpu_launches"], clocks=r["clocks"], roofline=roofline, roofline_hbm=roof_hbm, roofline_imad=roof_imad, proof_bound=bound,
                    kernel_ms_per_step={k_: round(v[0], 3) for k_, v in r["ktimes"].items()}, kernel_timed_step_ms=r["ms_ktimed"],
                    kernel_timing_coverage=timing_coverage(r["ktimes"], r["ms_ktimed"], env.devices))
        if "latency" in r:
            line["single_proof_p50_ms"] = r["latency"]["p50_ms"]
            line["single_proof_latency"] = r["latency"]
        if r["value"] is None:
            line["value_note"] = "single-process multi-GPU mode: the batch starts in host memory, value = e2e"
        line["cpu_baseline"] = cpu_check_proofs(env, r, args) if cpu_ok else None
        r["pk"].release()
        del r
        if not args.no_extras:
            # second headline: the same batch with two lookup arguments in the circuit
            lk = run_proofs(env, "withdraw_lookup", args.batch, 1, 1, full=False)
            line["withdraw_lookup"] = {"value": lk["value"], "unit": UNIT, "e2e": lk["e2e"], "proofs_per_step": lk["Mtot"] * env.world,
                                       "msm_per_proof": lk["shape"].num_msm, "proof_bytes": lk["pk"].proof_len,
                                       "kernel_ms_per_step": {k_: round(v[0], 3) for k_, v in lk["ktimes"].items()},
                                       "kernel_timing_coverage": timing_coverage(lk["ktimes"], lk["ms_ktimed"], env.devices),
                                       "note": "64 distinct witnesses cycled over the batch (seeds distinct), 1 warm-up + 1 timed step"}
            lk["pk"].release()
            del lk
            mx = run_mixed(env, args.requests, 1, 1)
            line["mixed_stream"] = {"workload": "BASELINE configs[4]: %d requests new_account/deposit/withdraw 1:2:2 (seed 7), least-loaded assignment over %d GPU(s), "
                                                "three circuits proved concurrently, pinned host buffers -> proofs on the host, ChaCha20 rng seeds" % (args.requests, N),
                                    "value": mx["value"], "unit": UNIT, "ms_per_step": mx["ms_per_step"], "mix": mx["mix"], "scaling": "strong",
                                    "steps": 1, "warmup": 1,
                                    "h2d_bytes_per_step": mx["h2d"], "d2h_bytes_per_step": mx["d2h"],
                                    "verified": cpu_check_mixed(env, mx) if cpu_ok else "N = 1 run only (CPU verifier)"}
            for j in mx["jobs"].values():
                j["pk"].release()
            del mx
            ms = run_msm(env, args.msm_log_n, 2, 1)
            line["msm24"] = {"workload": "BASELINE configs[1] at 2^%d: one G1 MSM, points split over %d GPU(s), bases resident per shard" % (args.msm_log_n, N),
                             "points_per_s_resident": ms["value"], "points_per_s_e2e": ms["e2e"], "scaling": "strong",
                             **{k_: ms[k_] for k_ in ("points_per_gpu", "ms_per_msm_e2e", "ms_per_msm_resident", "kernel_ms", "h2d_ms_estimate", "window_bits",
                                                      "windows", "roofline", "checked")}}
    elif args.workload == "mixed":
        mx = run_mixed(env, args.requests, args.steps, max(args.warmup, 3))
        line = dict(common, metric="Shielder halo2 proofs/sec (mixed new_account/deposit/withdraw stream)", value=mx["value"], unit=UNIT,
                    ms_per_step=mx["ms_per_step"], scaling="strong",
                    config={"workload": "BASELINE configs[4]: %d requests 1:2:2 (seed 7), each request to the least-loaded GPU (deterministic work queue), three circuits proved concurrently" % args.requests,
                            "mix": mx["mix"], "requests_of_rank0": mx["per_rank"], "parallelism": par, "rng": "ChaCha20 seed per request",
                            "l2": "inputs larger than L2"},
                    e2e={"value": mx["value"], "unit": UNIT, "ms_per_step": mx["ms_per_step"], "h2d_bytes_per_step": mx["h2d"], "d2h_bytes_per_step": mx["d2h"]},
                    value_note="the stream starts in pinned host memory: value = e2e", gpu_launches=env.zkgpu.launch_count(), roofline=None,
                    verified=cpu_check_mixed(env, mx) if cpu_ok else "N = 1 run only (CPU verifier)", cpu_baseline=None)
    else:
        ms = run_msm(env, args.msm_log_n, args.steps, max(args.warmup, 3))
        line = dict(common, metric="BN254 G1 MSM points/sec (2^%d points, split across GPUs)" % args.msm_log_n, value=ms["value"], unit="points/s",
                    ms_per_step=ms["ms_per_msm_resident"], scaling="strong",
                    config={"workload": "BASELINE configs[1] at 2^%d: one G1 MSM, points split over %d GPU(s), bases resident per shard, partial points summed"
                                        % (args.msm_log_n, N), "points_per_gpu": ms["points_per_gpu"], "window_bits": ms["window_bits"], "parallelism": par,
                            "l2": "inputs larger than L2 (%.0f MB of scalars + %.0f MB of bases per GPU)" % (ms["points_per_gpu"] * 32 / 1e6, ms["points_per_gpu"] * 64 / 1e6)},
                    e2e={"value": ms["e2e"], "unit": "points/s", "ms_per_step": ms["ms_per_msm_e2e"], "h2d_bytes_per_step": ms["h2d"], "d2h_bytes_per_step": ms["d2h"]},
                    kernel_ms=ms["kernel_ms"], h2d_ms_estimate=ms["h2d_ms_estimate"], roofline=ms["roofline"], checked=ms["checked"],
                    gpu_launches=env.zkgpu.launch_count(), cpu_baseline=None)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
