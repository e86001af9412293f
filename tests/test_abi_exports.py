"""CPU-only: libzkgpu.so builds, loads, exports every symbol include/zkgpu.h declares, and refuses
to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import zkgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(zkgpu.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return zkgpu.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zkgpu_\w+)\s*\(", hdr)))


def test_header_declares_boundary():
    syms = declared_symbols()
    for s in ("zkgpu_init", "zkgpu_msm_g1", "zkgpu_srs_register", "zkgpu_msm_g1_srs", "zkgpu_msm_g1_srs_batch",
              "zkgpu_ntt_fr", "zkgpu_ntt_fr_batch", "zkgpu_coset_ntt_fr", "zkgpu_coset_intt_fr", "zkgpu_fft_g1",
              "zkgpu_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    for s in declared_symbols():
        assert hasattr(built, s), "libzkgpu.so does not export " + s
    assert built.zkgpu_abi_version() >= 1


def test_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", zkgpu.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = np.zeros((4, 4), dtype=np.uint64)
    with pytest.raises(zkgpu.ZkGpuError):
        zkgpu.best_fft(a, np.zeros(4, dtype=np.uint64), 2)
    with pytest.raises(zkgpu.ZkGpuError):
        zkgpu.best_multiexp(a, np.zeros((4, 8), dtype=np.uint64))
