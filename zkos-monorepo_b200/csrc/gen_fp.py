#!/usr/bin/env python3
"""Generates csrc/fp_gen.inc — the 8x32-bit-limb Montgomery arithmetic for BN254 Fr and Fq.

One instruction list (a tiny PTX-like IR) is the single source of truth.  It is emitted twice:
  * for the device as inline-PTX carry chains (mad.lo.cc / madc.hi.cc ...; one asm statement per
    carry chain so no flag ever crosses an asm boundary).  ptxas pairs each lo/hi couple that targets
    an aligned register pair into one IMAD.WIDE.U32 with carry-in/out, i.e. 64 wide multiplies for
    the a*b part and 64 for the reduction;
  * for the host as a C emulation of the very same instruction list (explicit carry flag), so the
    limb algorithm is checked on the CPU (tests/test_fp_emulation.py) before it ever runs on a GPU.

Multiplication layout ("even/odd" accumulators): products a[j]*b_i for even j accumulate into
even[j..j+1], for odd j into odd[j-1..j] (odd[] is offset one limb up), so every 64-bit partial
product lands on an aligned register pair.  After adding m*MOD the frame shifts one limb and the
two arrays swap roles.  Fields: reference symbols bn256::Fr / bn256::Fq.
"""
import os
import sys

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
N = 8


def limbs32(v, n=N):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


class Prog:
    """Straight-line program: list of chains; a chain is a list of (op, dst, srcs...)."""

    def __init__(self):
        self.chains = []
        self.cur = None

    def chain(self):
        self.cur = []
        self.chains.append(self.cur)

    def ins(self, op, dst, *srcs):
        self.cur.append((op, dst) + tuple(srcs))


def is_imm(x):
    return isinstance(x, int)


def emit_ptx_chain(chain):
    """One asm volatile statement for the chain."""
    ops = {}  # var -> index
    written = set()
    order = []

    def reg(v):
        if v not in ops:
            ops[v] = len(order)
            order.append(v)
        return "%%%d" % ops[v]

    # collect in order so outputs come first
    for ins in chain:
        written.add(ins[1])
    outs = []
    for ins in chain:
        if ins[1] not in outs:
            outs.append(ins[1])
    for v in outs:
        reg(v)
    lines = []
    for ins in chain:
        op, dst, srcs = ins[0], ins[1], ins[2:]
        args = [reg(dst)]
        for s in srcs:
            args.append(("0x%x" % s) if is_imm(s) else reg(s))
        lines.append("%s.u32 %s;" % (op, ", ".join(args)))
    outs_c = ", ".join('"+r"(%s)' % v for v in order if v in written)
    ins_c = ", ".join('"r"(%s)' % v for v in order if v not in written)
    body = " ".join(lines)
    return '    asm volatile("%s" : %s : %s);' % (body, outs_c, ins_c) if ins_c else \
           '    asm volatile("%s" : %s);' % (body, outs_c)


def emit_c_chain(chain, check_carry_free):
    """Host emulation of the same chain with an explicit carry flag."""
    out = ["    { uint32_t cc = 0; (void)cc;"]

    def val(s):
        return ("0x%xu" % s) if is_imm(s) else s

    for ins in chain:
        op, dst, srcs = ins[0], ins[1], [val(s) for s in ins[2:]]
        base = op.split(".")
        name = base[0]
        cin = name.endswith("c") and name not in ("sub",) and name in ("madc", "addc", "subc")
        cout = op.endswith(".cc")
        if name in ("mul",):
            half = base[1]
            e = "(uint64_t)%s * %s" % (srcs[0], srcs[1])
            out.append("      %s = (uint32_t)((%s)%s);" % (dst, e, " >> 32" if half == "hi" else ""))
        elif name in ("mad", "madc"):
            half = base[1]
            prod = "(uint32_t)(((uint64_t)%s * %s)%s)" % (srcs[0], srcs[1], " >> 32" if half == "hi" else "")
            out.append("      { uint64_t w_ = (uint64_t)%s + %s + %s; %s = (uint32_t)w_; %s }" % (
                prod, srcs[2], "cc" if cin else "0", dst, "cc = (uint32_t)(w_ >> 32);" if cout else ""))
        elif name in ("add", "addc"):
            out.append("      { uint64_t w_ = (uint64_t)%s + %s + %s; %s = (uint32_t)w_; %s }" % (
                srcs[0], srcs[1], "cc" if cin else "0", dst, "cc = (uint32_t)(w_ >> 32);" if cout else ""))
        elif name in ("sub", "subc"):
            out.append("      { uint64_t w_ = (uint64_t)%s - %s - %s; %s = (uint32_t)w_; %s }" % (
                srcs[0], srcs[1], "cc" if cin else "0", dst, "cc = (uint32_t)(w_ >> 63);" if cout else ""))
        else:
            raise ValueError(op)
    if check_carry_free:
        out.append("      ZK_EMU_ASSERT(cc == 0);")
    out.append("    }")
    return "\n".join(out)


def gen_mont_mul(mod, sqr=False):
    """Returns list of (chain, carry_must_be_zero) computing r = a*b/2^256 mod p into r[0..7]."""
    M = limbs32(mod)
    m0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    p = Prog()
    flags = []
    A = ["a[%d]" % i for i in range(N)]
    Bv = ["b[%d]" % i for i in range(N)] if not sqr else A
    ev = ["ev[%d]" % i for i in range(N)]
    od = ["od[%d]" % i for i in range(N)]

    def step(e, o, bi, first):
        if first:
            p.chain(); flags.append(False)
            for j in range(0, N, 2):
                p.ins("mul.lo", o[j], A[j + 1], bi)
                p.ins("mul.hi", o[j + 1], A[j + 1], bi)
            for j in range(0, N, 2):
                p.ins("mul.lo", e[j], A[j], bi)
                p.ins("mul.hi", e[j + 1], A[j], bi)
        else:
            p.chain(); flags.append(False)
            p.ins("add.cc", e[0], e[0], o[1])
            for j in range(0, N - 2, 2):
                p.ins("madc.lo.cc", o[j], A[j + 1], bi, o[j + 2])
                p.ins("madc.hi.cc", o[j + 1], A[j + 1], bi, o[j + 3])
            p.ins("madc.lo.cc", o[N - 2], A[N - 1], bi, 0)
            p.ins("madc.hi", o[N - 1], A[N - 1], bi, 0)
            p.chain(); flags.append(False)
            p.ins("mad.lo.cc", e[0], A[0], bi, e[0])
            p.ins("madc.hi.cc", e[1], A[0], bi, e[1])
            for j in range(2, N, 2):
                p.ins("madc.lo.cc", e[j], A[j], bi, e[j])
                p.ins("madc.hi.cc", e[j + 1], A[j], bi, e[j + 1])
            p.ins("addc", o[N - 1], o[N - 1], 0)
        p.chain(); flags.append(False)
        p.ins("mul.lo", "mi", e[0], m0)
        # odd += MOD[odd limbs] * mi   (top carry is provably zero; the emulation asserts it)
        p.chain(); flags.append(True)
        p.ins("mad.lo.cc", o[0], M[1], "mi", o[0])
        p.ins("madc.hi.cc", o[1], M[1], "mi", o[1])
        for j in range(2, N, 2):
            p.ins("madc.lo.cc", o[j], M[j + 1], "mi", o[j])
            p.ins("madc.hi.cc", o[j + 1], M[j + 1], "mi", o[j + 1])
        p.chain(); flags.append(False)
        p.ins("mad.lo.cc", e[0], M[0], "mi", e[0])
        p.ins("madc.hi.cc", e[1], M[0], "mi", e[1])
        for j in range(2, N, 2):
            p.ins("madc.lo.cc", e[j], M[j], "mi", e[j])
            p.ins("madc.hi.cc", e[j + 1], M[j], "mi", e[j + 1])
        p.ins("addc", o[N - 1], o[N - 1], 0)

    for i in range(0, N, 2):
        step(ev, od, Bv[i], i == 0)
        step(od, ev, Bv[i + 1], False)
    # merge: ev[i] += od[i+1]
    p.chain(); flags.append(False)
    p.ins("add.cc", ev[0], ev[0], od[1])
    for i in range(1, N - 1):
        p.ins("addc.cc", ev[i], ev[i], od[i + 1])
    p.ins("addc", ev[N - 1], ev[N - 1], 0)
    # conditional subtract: t = ev - p ; borrow -> keep ev
    p.chain(); flags.append(False)
    p.ins("sub.cc", "t[0]", ev[0], M[0])
    for i in range(1, N):
        p.ins("subc.cc", "t[%d]" % i, ev[i], M[i])
    p.ins("subc", "bw", 0, 0)
    return list(zip(p.chains, flags))


def gen_add(mod):
    M = limbs32(mod)
    p = Prog()
    p.chain()
    p.ins("add.cc", "s[0]", "a[0]", "b[0]")
    for i in range(1, N - 1):
        p.ins("addc.cc", "s[%d]" % i, "a[%d]" % i, "b[%d]" % i)
    p.ins("addc", "s[%d]" % (N - 1), "a[%d]" % (N - 1), "b[%d]" % (N - 1))
    p.chain()
    p.ins("sub.cc", "t[0]", "s[0]", M[0])
    for i in range(1, N):
        p.ins("subc.cc", "t[%d]" % i, "s[%d]" % i, M[i])
    p.ins("subc", "bw", 0, 0)
    return [(c, False) for c in p.chains]


def gen_sub(mod):
    M = limbs32(mod)
    p = Prog()
    p.chain()
    p.ins("sub.cc", "s[0]", "a[0]", "b[0]")
    for i in range(1, N):
        p.ins("subc.cc", "s[%d]" % i, "a[%d]" % i, "b[%d]" % i)
    p.ins("subc", "bw", 0, 0)
    # t = s + (p & mask) where mask = bw (all ones if borrow)
    p.chain()
    p.ins("add.cc", "t[0]", "s[0]", "pm[0]")
    for i in range(1, N - 1):
        p.ins("addc.cc", "t[%d]" % i, "s[%d]" % i, "pm[%d]" % i)
    p.ins("addc", "t[%d]" % (N - 1), "s[%d]" % (N - 1), "pm[%d]" % (N - 1))
    return [(c, False) for c in p.chains]


def body(chains, device):
    out = []
    for ch, flag in chains:
        out.append(emit_ptx_chain(ch) if device else emit_c_chain(ch, flag))
    return "\n".join(out)


def emit_field(name, mod):
    M = limbs32(mod)
    R = limbs32((1 << 256) % mod)
    R2 = limbs32(pow(1 << 256, 2, mod))
    o = []
    o.append("// ---- %s: p = 0x%x ----" % (name, mod))
    o.append("struct %s_params {" % name)
    o.append("    static constexpr uint32_t MOD[8] = {%s};" % ", ".join("0x%08xu" % x for x in M))
    o.append("    static constexpr uint32_t ONE[8] = {%s};  // R mod p" % ", ".join("0x%08xu" % x for x in R))
    o.append("    static constexpr uint32_t R2[8]  = {%s};  // R^2 mod p" % ", ".join("0x%08xu" % x for x in R2))
    o.append("};")
    for cname, vals in (("one", R), ("r2", R2), ("modm2", limbs32(mod - 2))):
        o.append("ZK_FP_FN void %s_set_%s(uint32_t* r) { %s }" % (name, cname, " ".join("r[%d] = 0x%08xu;" % (i, v) for i, v in enumerate(vals))))
    for fn, gen in (("mul", lambda: gen_mont_mul(mod)), ("sqr", lambda: gen_mont_mul(mod, sqr=True))):
        sig = "ZK_FP_FN void %s_%s(uint32_t* __restrict__ r, const uint32_t* __restrict__ a%s)" % (
            name, fn, ", const uint32_t* __restrict__ b" if fn == "mul" else "")
        chains = gen()
        o.append(sig + " {")
        o.append("    uint32_t ev[8] = {0,0,0,0,0,0,0,0}, od[8] = {0,0,0,0,0,0,0,0}, t[8] = {0,0,0,0,0,0,0,0}, mi = 0, bw = 0;")
        o.append("#ifdef __CUDA_ARCH__")
        o.append(body(chains, True))
        o.append("#else")
        o.append(body(chains, False))
        o.append("#endif")
        o.append("    for (int i = 0; i < 8; ++i) r[i] = bw ? ev[i] : t[i];")
        o.append("}")
    # add
    o.append("ZK_FP_FN void %s_add(uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {" % name)
    o.append("    uint32_t s[8] = {0,0,0,0,0,0,0,0}, t[8] = {0,0,0,0,0,0,0,0}, bw = 0;")
    o.append("#ifdef __CUDA_ARCH__")
    o.append(body(gen_add(mod), True))
    o.append("#else")
    o.append(body(gen_add(mod), False))
    o.append("#endif")
    o.append("    for (int i = 0; i < 8; ++i) r[i] = bw ? s[i] : t[i];")
    o.append("}")
    # sub
    o.append("ZK_FP_FN void %s_sub(uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {" % name)
    o.append("    uint32_t s[8] = {0,0,0,0,0,0,0,0}, t[8] = {0,0,0,0,0,0,0,0}, pm[8], bw = 0;")
    chains = gen_sub(mod)
    for dev in (True, False):
        o.append("#ifdef __CUDA_ARCH__" if dev else "#else")
        o.append(body(chains[:1], dev))
        o.append("    " + " ".join("pm[%d] = 0x%08xu & bw;" % (i, M[i]) for i in range(8)))
        o.append(body(chains[1:], dev))
    o.append("#endif")
    o.append("    for (int i = 0; i < 8; ++i) r[i] = t[i];")
    o.append("}")
    return "\n".join(o)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = ["// GENERATED by gen_fp.py — do not edit.  8x32-bit Montgomery arithmetic for BN254 Fr / Fq.",
           "// Device path: inline-PTX carry chains.  Host path: C emulation of the same instruction list.",
           "#pragma once", "#include <stdint.h>",
           "#ifndef ZK_FP_FN", "#ifdef __CUDACC__", "#define ZK_FP_FN __host__ __device__ __forceinline__",
           "#else", "#define ZK_FP_FN static inline", "#endif", "#endif",
           "#ifndef ZK_EMU_ASSERT", "#define ZK_EMU_ASSERT(x) ((void)0)", "#endif", ""]
    out.append(emit_field("fr", R_MOD))
    out.append("")
    out.append(emit_field("fq", Q_MOD))
    out.append("")
    # Fr domain constants in Montgomery form (bn256::Fr::{ROOT_OF_UNITY, ZETA, DELTA})
    mont = lambda v: limbs32(v * (1 << 256) % R_MOD)
    root = pow(7, (R_MOD - 1) >> 28, R_MOD)
    zeta = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
    assert pow(zeta, 3, R_MOD) == 1 and zeta != 1
    consts = {"ROOT_OF_UNITY": root, "ROOT_OF_UNITY_INV": pow(root, -1, R_MOD), "ZETA": zeta,
              "ZETA_INV": zeta * zeta % R_MOD, "DELTA": pow(7, 1 << 28, R_MOD), "TWO_INV": pow(2, -1, R_MOD)}
    out.append("struct fr_consts {")
    for k, v in consts.items():
        out.append("    static constexpr uint32_t %s[8] = {%s};" % (k, ", ".join("0x%08xu" % x for x in mont(v))))
    out.append("    static constexpr unsigned S = 28;")
    out.append("};")
    with open(os.path.join(here, "fp_gen.inc"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
