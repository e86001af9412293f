"""Field backend for zkgpu.circuits backed by libzkgpu (vectorised Fr ops on the GPU): used by bench.py and
the tools to build synthetic circuits and witnesses without touching the CPU oracle."""
import ctypes as C

import numpy as np

from . import _chk, _p, _u64, lib

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def _vec_op(op, a, b):
    a, b = _u64(a).reshape(-1, 4), _u64(b).reshape(-1, 4)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = np.empty_like(a)
    _chk(lib().zkgpu_fr_vec_op(op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0])))
    return out


class GpuBackend:
    @staticmethod
    def random(seed, count):
        """count uniform field elements: Fr::random over SmallRng::seed_from_u64(seed)"""
        out = np.empty((count, 4), dtype=np.uint64)
        _chk(lib().zkgpu_fr_random(C.c_uint64(seed), _p(out), C.c_size_t(count)))
        return out

    @staticmethod
    def const(v):
        v %= R_MOD
        canon = np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]], dtype=np.uint64)
        out = np.empty_like(canon)
        _chk(lib().zkgpu_fr_to_mont(_p(canon), _p(out), C.c_size_t(1)))
        return out[0]

    @staticmethod
    def to_mont(canon):
        canon = _u64(canon).reshape(-1, 4)
        out = np.empty_like(canon)
        _chk(lib().zkgpu_fr_to_mont(_p(canon), _p(out), C.c_size_t(canon.shape[0])))
        return out

    @staticmethod
    def mul(a, b):
        return _vec_op(0, a, b)

    @staticmethod
    def add(a, b):
        return _vec_op(1, a, b)

    @staticmethod
    def sub(a, b):
        return _vec_op(2, a, b)
