#!/usr/bin/env python3
"""BASELINE configs[1] and [2] as measurements: standalone G1 MSM sweep 2^11..2^24 and Fr NTT sweep 2^11..2^22 through the
drop-in C ABI (host buffers), with the kernel-only device time from the library's CUDA-event timers and the CPU oracle
(restated halo2curves best_multiexp / best_fft, all host threads) beside it.  Writes a markdown table to stdout."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "zkos-monorepo_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_lib as O
import pyref as P
import zkgpu
from zkgpu.gpu_backend import GpuBackend

zkgpu.init(0)
L = zkgpu.lib()
cores = os.cpu_count() or 1


def ktime(slots):
    tot = 0.0
    for s in slots:
        ms, cnt = C.c_double(0), C.c_uint64(0)
        L.zkgpu_kernel_times(s, C.byref(ms), C.byref(cnt), 1)
        tot += ms.value
    return tot


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


print("## G1 MSM sweep (`zkgpu_msm_g1` = best_multiexp; bases = setup SRS g[i] = s^i G, uniform scalars)\n")
print("| n | GPU call, host buffers (ms) | device kernels only (ms) | G point-adds/s (kernels) | CPU oracle, %d threads (ms) | speed-up (call) |" % cores)
print("|---|---:|---:|---:|---:|---:|")
for log_n in (11, 12, 13, 14, 16, 18, 20, 22, 24):
    n = 1 << log_n
    g = zkgpu.params_setup(log_n, 7, lagrange=False)[0]
    c = GpuBackend.random(11, n)
    zkgpu.best_multiexp(c, g)
    t_call = best(lambda: zkgpu.best_multiexp(c, g), 3)
    L.zkgpu_kernel_timing(1); ktime(range(3))
    zkgpu.best_multiexp(c, g)
    t_k = ktime(range(3)); L.zkgpu_kernel_timing(0)
    t_cpu = best(lambda: O.msm(c, g, threads=cores), 1) if log_n <= 18 else None
    print("| 2^%d | %.2f | %.2f | %.2f | %s | %s |" % (log_n, 1e3 * t_call, t_k, n * (254 // 13 + 1) / (t_k / 1e3) / 1e9 if t_k else 0,
                                                      "%.1f" % (1e3 * t_cpu) if t_cpu else "—", "%.0fx" % (t_cpu / t_call) if t_cpu else "—"))
    del g, c

print("\n## Fr NTT sweep (`zkgpu_ntt_fr` = best_fft, natural order in and out)\n")
print("| n | GPU call, host buffers (ms) | device kernels only (ms) | kernel GB/s (64 B/point) | kernel G mul/s | CPU oracle, %d threads (ms) | speed-up (call) |" % cores)
print("|---|---:|---:|---:|---:|---:|---:|")
for log_n in (11, 12, 13, 14, 16, 18, 20, 22):
    n = 1 << log_n
    a = GpuBackend.random(5, n)
    w = O.to_mont(0, P.int_to_limbs([P.omega_for(log_n)]))[0]
    zkgpu.best_fft(a, w, log_n)
    t_call = best(lambda: zkgpu.best_fft(a, w, log_n), 3)
    L.zkgpu_kernel_timing(1); ktime([3])
    zkgpu.best_fft(a, w, log_n)
    t_k = ktime([3]); L.zkgpu_kernel_timing(0)
    t_cpu = best(lambda: O.fft(a, w, log_n, threads=cores), 1) if log_n <= 20 else None
    print("| 2^%d | %.3f | %.3f | %.0f | %.1f | %s | %s |" % (log_n, 1e3 * t_call, t_k, 64 * n / (t_k / 1e3) / 1e9, n / 2 * log_n / (t_k / 1e3) / 1e9,
                                                            "%.1f" % (1e3 * t_cpu) if t_cpu else "—", "%.0fx" % (t_cpu / t_call) if t_cpu else "—"))

print("\n## Batched transforms as the prover issues them (device resident)\n")
import torch
print("| transform | batch | ms | GB/s (64 B/point) | G mul/s |")
print("|---|---:|---:|---:|---:|")
for log_n, batch in ((13, 3200), (16, 512)):
    n = 1 << log_n
    x = torch.randint(0, 2**62, (batch, n, 4), dtype=torch.int64, device="cuda")
    x[..., 3] &= 0x0fffffffffffffff
    w = O.to_mont(0, P.int_to_limbs([P.omega_for(log_n)]))[0]
    scratch = torch.empty_like(x)
    for _ in range(2):
        zkgpu.ntt_batch_dev(x.data_ptr(), w, log_n, batch, scratch.data_ptr())
    torch.cuda.synchronize()
    L.zkgpu_kernel_timing(1); ktime([3])
    zkgpu.ntt_batch_dev(x.data_ptr(), w, log_n, batch, scratch.data_ptr())
    torch.cuda.synchronize()
    t_k = ktime([3]); L.zkgpu_kernel_timing(0)
    print("| best_fft 2^%d | %d | %.2f | %.0f | %.1f |" % (log_n, batch, t_k, 64 * n * batch / (t_k / 1e3) / 1e9, batch * n / 2 * log_n / (t_k / 1e3) / 1e9))
