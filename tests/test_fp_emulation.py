"""The GPU's 8x32-bit Montgomery limb algorithm (generated PTX carry chains, csrc/fp_gen.inc) is
emitted a second time as a C emulation of the same instruction list; this test runs that emulation
on the CPU against the oracle, including the 'carry provably zero' assertions inside it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pyref as P

ROOT = O.ROOT
CSRC = os.path.join(ROOT, "zkos-monorepo_b200", "csrc")


@pytest.fixture(scope="module", params=["schoolbook", "karatsuba", "plain_sqr"])
def emu(request, tmp_path_factory):
    """schoolbook = the shipped interleaved product with the triangular squaring (regenerated in place and required to be unchanged);
    karatsuba = the optional ZK_FP_KARATSUBA=1 variant, generated into a scratch directory."""
    d = tmp_path_factory.mktemp("emu_" + request.param)
    if request.param == "schoolbook":
        shipped = open(os.path.join(CSRC, "fp_gen.inc")).read()
        subprocess.check_call(["python3", os.path.join(CSRC, "gen_fp.py"), str(d / "fp_gen.inc")], env=dict(os.environ, ZK_FP_KARATSUBA="0"))
        assert open(str(d / "fp_gen.inc")).read() == shipped, "csrc/fp_gen.inc is stale: run gen_fp.py"
    elif request.param == "plain_sqr":   # squaring emitted as mul(a, a) instead of the triangular product
        subprocess.check_call(["python3", os.path.join(CSRC, "gen_fp.py"), str(d / "fp_gen.inc")], env=dict(os.environ, ZK_FP_KARATSUBA="0", ZK_FP_TRISQR="0"))
    else:
        subprocess.check_call(["python3", os.path.join(CSRC, "gen_fp.py"), str(d / "fp_gen.inc")], env=dict(os.environ, ZK_FP_KARATSUBA="1"))
    so = str(d / "libfpemu.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", str(d), "-I", CSRC,
                           os.path.join(ROOT, "tests", "emu", "fp_emu.cpp"), "-o", so])
    return C.CDLL(so)


def run(emu, field, op, a, b=None):
    a32 = np.ascontiguousarray(a, dtype=np.uint64).view(np.uint32)
    b32 = np.ascontiguousarray(b, dtype=np.uint64).view(np.uint32) if b is not None else None
    out = np.empty_like(a32)
    emu.emu_field_op(field, op, a32.ctypes.data_as(C.c_void_p), b32.ctypes.data_as(C.c_void_p) if b32 is not None else None,
                     out.ctypes.data_as(C.c_void_p), C.c_size_t(a32.size // 8))
    return out.view(np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("field,p", [(0, P.R_MOD), (1, P.Q_MOD)])
def test_limb_arithmetic_matches_oracle(emu, field, p):
    rng = np.random.default_rng(7 + field)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 253) - 1, (1 << 253), (1 << 32) - 1, (1 << 64) - 1, 0xFFFFFFFF << 224 | 5,
            (1 << 128) - 1, 1 << 128, ((1 << 125) << 128) | (1 << 125), ((1 << 128) - 1) << 120, (3 << 128) | 7, (7 << 128) | 3,
            # bit 31 of a limb is the bit the triangular squaring moves between the doubled limbs
            0x7fffffff, 0x80000000, 0x80000000 << 32, (0x80000000 << 192) | 0x80000000, int("80000000" * 7, 16), int("ffffffff" * 7, 16)]
    edge = [e % p for e in edge]
    vals_a = edge * len(edge) + [int.from_bytes(rng.bytes(40), "little") % p for _ in range(20000)]
    vals_b = [e for e in edge for _ in edge] + [int.from_bytes(rng.bytes(40), "little") % p for _ in range(20000)]
    # feed raw limbs (any value < p is a valid Montgomery representative)
    a = P.int_to_limbs(vals_a)
    b = P.int_to_limbs(vals_b)
    for op, oop in ((0, 0), (1, 1), (2, 2)):
        assert np.array_equal(run(emu, field, op, a, b), O.field_op(field, oop, a, b).reshape(-1, 4)), op
    assert np.array_equal(run(emu, field, 3, a), O.field_op(field, 4, a).reshape(-1, 4))



@pytest.mark.parametrize("field,p", [(0, P.R_MOD), (1, P.Q_MOD)])
def test_dual_product_with_one_reduction(emu, field, p):
    """fe_mul_add2(a, b, c, d) = a*b + c*d (the Y-coordinate of every point addition): same value as two products and a modular
    addition, on extreme operands (all four at p - 1 maximise every partial sum of the nine-limb frame) and random ones."""
    rng = np.random.default_rng(17 + field)
    edge = [0, 1, p - 1, p - 2, (1 << 253), (1 << 32) - 1, int("ffffffff" * 7, 16) % p, int("80000000" * 8, 16) % p, (p - 1) // 2]
    quads = [(a, b, c, d) for a in edge for b in edge[:5] for c in edge[2:6] for d in edge]
    quads += [tuple(int.from_bytes(rng.bytes(40), "little") % p for _ in range(4)) for _ in range(20000)]
    cols = [P.int_to_limbs([q[i] for q in quads]) for i in range(4)]
    u32 = [np.ascontiguousarray(c, dtype=np.uint64).view(np.uint32) for c in cols]
    out = np.empty_like(u32[0])
    emu.emu_mul_add2(field, *[x.ctypes.data_as(C.c_void_p) for x in u32], out.ctypes.data_as(C.c_void_p), C.c_size_t(len(quads)))
    want = O.field_op(field, 1, O.field_op(field, 0, cols[0], cols[1]), O.field_op(field, 0, cols[2], cols[3])).reshape(-1, 4)
    assert np.array_equal(out.view(np.uint64).reshape(-1, 4), want)
