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
        lines.append("%s.%s %s;" % (op, "b32" if op in ("xor", "and") else "u32", ", ".join(args)))
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
        elif name in ("xor", "and"):
            out.append("      %s = %s %s %s;" % (dst, srcs[0], "^" if name == "xor" else "&", srcs[1]))
        elif name in ("sub", "subc"):
            out.append("      { uint64_t w_ = (uint64_t)%s - %s - %s; %s = (uint32_t)w_; %s }" % (
                srcs[0], srcs[1], "cc" if cin else "0", dst, "cc = (uint32_t)(w_ >> 63);" if cout else ""))
        else:
            raise ValueError(op)
    if check_carry_free is True:
        out.append("      ZK_EMU_ASSERT(cc == 0);")
    elif check_carry_free:
        out.append("      ZK_EMU_ASSERT(%s == 0);" % check_carry_free)
    out.append("    }")
    return "\n".join(out)


def gen_mont_mul(mod, sqr=False, tri=False, dual=False):
    """Returns list of (chain, carry_must_be_zero) computing r = a*b/2^256 mod p into r[0..7].

    dual: r = (a*b + c*d)/2^256 mod p with ONE interleaved reduction — step i accumulates a*b_i and c*d_i before m_i is
    formed, so the second product costs its 64 wide multiplies but no second reduction (64 + 8 multiplies, a conditional
    subtraction and the modular addition of the two results are saved).  Bounds: every partial sum is below
    2^32 (3p + 1) < 2^288, i.e. fits the nine-limb frame of the even/odd accumulators (the emulation asserts the top
    carries), and the result (ab + cd + m p)/2^256 < p (2p/2^256 + 1) < 2p needs one conditional subtraction.

    tri (squaring only): triangular product.  a^2 = sum_i a_i 2^(32 i) * (a_i 2^(32 i) + sum_{j>i} 2 a_j 2^(32 j)), so step i
    multiplies a_i by the limbs j >= i of (a_i, 2a) only: 36 wide products instead of 64.  d = 2a is one carry chain of eight
    adds; the limbs j > i of 2a are d[j] except that bit 0 of d[i+1] (= bit 31 of a_i) belongs to 2 a_i: dm[j] = d[j] & ~1.
    Skipped products become plain carry-propagating adds (ALU pipe), which still perform the frame shift of the interleaved
    reduction."""
    M = limbs32(mod)
    m0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    p = Prog()
    flags = []
    A = ["a[%d]" % i for i in range(N)]
    Bv = ["b[%d]" % i for i in range(N)] if not sqr else A
    Cv = ["c[%d]" % i for i in range(N)]
    Dv = ["d[%d]" % i for i in range(N)]
    assert not (dual and sqr)
    ev = ["ev[%d]" % i for i in range(N)]
    od = ["od[%d]" % i for i in range(N)]
    tri = tri and sqr
    if tri:
        p.chain(); flags.append(False)
        p.ins("add.cc", "d[0]", A[0], A[0])
        for i in range(1, N - 1):
            p.ins("addc.cc", "d[%d]" % i, A[i], A[i])
        p.ins("addc", "d[%d]" % (N - 1), A[N - 1], A[N - 1])
        p.chain(); flags.append(False)
        for i in range(1, N):
            p.ins("and", "dm[%d]" % i, "d[%d]" % i, 0xFFFFFFFE)

    def mad(op, dst, x, bi, addend=None):
        if x is not None:
            if addend is None:
                p.ins(op, dst, x, bi)
            else:
                p.ins(op, dst, x, bi, addend)
            return
        name, cc = op.split(".")[0], op.endswith(".cc")
        assert name in ("mad", "madc") and addend is not None
        p.ins(("add" if name == "mad" else "addc") + (".cc" if cc else ""), dst, addend, 0)

    def step(e, o, bi, first, i):
        def X(j):
            if not tri:
                return A[j]
            if j < i:
                return None
            return A[j] if j == i else ("dm[%d]" % j if j == i + 1 else "d[%d]" % j)
        if first:
            p.chain(); flags.append(False)
            for j in range(0, N, 2):
                p.ins("mul.lo", o[j], X(j + 1), bi)
                p.ins("mul.hi", o[j + 1], X(j + 1), bi)
            for j in range(0, N, 2):
                p.ins("mul.lo", e[j], X(j), bi)
                p.ins("mul.hi", e[j + 1], X(j), bi)
        else:
            p.chain(); flags.append(False)
            p.ins("add.cc", e[0], e[0], o[1])
            for j in range(0, N - 2, 2):
                mad("madc.lo.cc", o[j], X(j + 1), bi, o[j + 2])
                mad("madc.hi.cc", o[j + 1], X(j + 1), bi, o[j + 3])
            mad("madc.lo.cc", o[N - 2], X(N - 1), bi, 0)
            mad("madc.hi", o[N - 1], X(N - 1), bi, 0)
            p.chain(); flags.append(False)
            j0 = 0
            while tri and j0 < N and X(j0) is None:      # leading skipped products of a chain without carry-in: nothing to do
                j0 += 2
            if j0 < N:
                p.ins("mad.lo.cc", e[j0], X(j0), bi, e[j0])
                p.ins("madc.hi.cc", e[j0 + 1], X(j0), bi, e[j0 + 1])
                for j in range(j0 + 2, N, 2):
                    mad("madc.lo.cc", e[j], X(j), bi, e[j])
                    mad("madc.hi.cc", e[j + 1], X(j), bi, e[j + 1])
                p.ins("addc", o[N - 1], o[N - 1], 0)
        if dual:
            # second product of the step: c * d_i into the same frame (odd limbs, then even limbs + carry into the top)
            di = Dv[i]
            p.chain(); flags.append(True)
            p.ins("mad.lo.cc", o[0], Cv[1], di, o[0])
            p.ins("madc.hi.cc", o[1], Cv[1], di, o[1])
            for j in range(2, N, 2):
                p.ins("madc.lo.cc", o[j], Cv[j + 1], di, o[j])
                p.ins("madc.hi.cc", o[j + 1], Cv[j + 1], di, o[j + 1])
            p.chain(); flags.append(False)
            p.ins("mad.lo.cc", e[0], Cv[0], di, e[0])
            p.ins("madc.hi.cc", e[1], Cv[0], di, e[1])
            for j in range(2, N, 2):
                p.ins("madc.lo.cc", e[j], Cv[j], di, e[j])
                p.ins("madc.hi.cc", e[j + 1], Cv[j], di, e[j + 1])
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
        step(ev, od, Bv[i], i == 0, i)
        step(od, ev, Bv[i + 1], False, i + 1)
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


class Acc:
    """Even/odd accumulator pair for a schoolbook product without reduction: E[k] is limb position k, O[k] limb position
    k+1, so every 64-bit partial product lands on an aligned register pair of one of the two arrays."""

    def __init__(self, p, flags, ename, oname, width):
        self.p, self.flags = p, flags
        self.E = ["%s[%d]" % (ename, i) for i in range(width)]
        self.O = ["%s[%d]" % (oname, i) for i in range(width)]
        self.init = set()

    def chain(self, arr, base, prods):
        """prods: list of (x, y) multiplied into consecutive limb pairs of `arr` starting at index `base`."""
        p = self.p
        p.chain(); self.flags.append(False)
        first = True
        carry_live = False
        for t, (x, y) in enumerate(prods):
            lo, hi = arr[base + 2 * t], arr[base + 2 * t + 1]
            for half, dst in (("lo", lo), ("hi", hi)):
                have = dst in self.init
                if not carry_live and not have:
                    p.ins("mul." + half, dst, x, y)          # fresh limb, no carry in flight: plain product half
                else:
                    op = ("mad." if not carry_live else "madc.") + half + ".cc"
                    p.ins(op, dst, x, y, dst if have else 0)
                    carry_live = True
                self.init.add(dst)
            first = False
        if carry_live:
            nxt = base + 2 * len(prods)
            if nxt < len(arr):
                dst = arr[nxt]
                p.ins("addc", dst, dst if dst in self.init else 0, 0)   # by construction the carry cannot ripple further
                self.init.add(dst)
            else:
                # top of the accumulator: the product fits, the carry is provably zero (asserted by the emulation)
                p.ins("addc", "cz", 0, 0)
                self.flags[-1] = "cz"

    def product(self, X, Y):
        """accumulates X * Y (lists of limb operand names) into E / O"""
        for i, y in enumerate(Y):
            ev = [(j, X[j]) for j in range(len(X)) if (i + j) % 2 == 0]
            odd = [(j, X[j]) for j in range(len(X)) if (i + j) % 2 == 1]
            if ev:
                self.chain(self.E, i + ev[0][0], [(x, y) for _, x in ev])
            if odd:
                self.chain(self.O, i + odd[0][0] - 1, [(x, y) for _, x in odd])

    def merge(self, out, n):
        """out[0..n) = E + (O << 32)"""
        p = self.p
        p.chain(); self.flags.append(False)
        zero = lambda v: v if v in self.init else 0
        p.ins("add.cc", out[0], zero(self.E[0]), 0)
        for k in range(1, n):
            p.ins("addc.cc" if k < n - 1 else "addc", out[k], zero(self.E[k]), zero(self.O[k - 1]))


def gen_mont_mul_karatsuba(mod, sqr=False):
    """r = a*b/2^256 mod p as: one-level (subtractive) Karatsuba for the 512-bit product (48 wide multiplies
    instead of 64), then a separate Montgomery reduction of its low half ('multiply by 1': 64 wide multiplies +
    8 low multiplies), add the high half, one conditional subtraction."""
    M = limbs32(mod)
    m0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    p = Prog()
    flags = []
    A = ["a[%d]" % i for i in range(N)]
    B = ["b[%d]" % i for i in range(N)] if not sqr else A
    H = N // 2
    z0 = ["z0[%d]" % i for i in range(N)]
    z2 = ["z2[%d]" % i for i in range(N)]
    zm = ["zm[%d]" % i for i in range(N)]
    da = ["da[%d]" % i for i in range(H)]
    db = ["db[%d]" % i for i in range(H)]
    mid = ["mid[%d]" % i for i in range(N + 1)]
    tt = ["tt[%d]" % i for i in range(2 * N)]
    # z0 = a_lo * b_lo, z2 = a_hi * b_hi
    for out, X, Y, en, on in ((z0, A[:H], B[:H], "pe0", "po0"), (z2, A[H:], B[H:], "pe2", "po2")):
        acc = Acc(p, flags, en, on, N)
        acc.product(X, Y)
        acc.merge(out, N)
    # da = |a_lo - a_hi|, db = |b_hi - b_lo| with sign masks sa, sb (all ones when negative)
    for d, X, Y, sm in ((da, A[:H], A[H:], "sa"), (db, B[H:], B[:H], "sb")):
        p.chain(); flags.append(False)
        p.ins("sub.cc", d[0], X[0], Y[0])
        for i in range(1, H):
            p.ins("subc.cc", d[i], X[i], Y[i])
        p.ins("subc", sm, 0, 0)
        p.chain(); flags.append(False)
        for i in range(H):
            p.ins("xor", d[i], d[i], sm)
        p.chain(); flags.append(False)
        p.ins("sub.cc", d[0], d[0], sm)
        for i in range(1, H):
            p.ins("subc.cc" if i < H - 1 else "subc", d[i], d[i], sm)
    acc = Acc(p, flags, "pe1", "po1", N)
    acc.product(da, db)
    acc.merge(zm, N)
    # mid = z0 + z2 + sign * zm  (= a_lo*b_hi + a_hi*b_lo >= 0, 9 limbs); sign is negative iff sa ^ sb ... note
    # (a_lo - a_hi)(b_hi - b_lo) = a_lo b_hi + a_hi b_lo - z0 - z2, so mid = z0 + z2 + (a_lo - a_hi)(b_hi - b_lo)
    p.chain(); flags.append(False)
    p.ins("xor", "sn", "sa", "sb")
    p.chain(); flags.append(False)
    p.ins("add.cc", mid[0], z0[0], z2[0])
    for i in range(1, N):
        p.ins("addc.cc", mid[i], z0[i], z2[i])
    p.ins("addc", mid[N], 0, 0)
    p.chain(); flags.append(False)
    for i in range(N):
        p.ins("xor", zm[i], zm[i], "sn")
    p.chain(); flags.append(False)
    p.ins("add.cc", "cz", "sn", 1)                       # carry = 1 iff the middle term is negative (two's complement +1)
    for i in range(N):
        p.ins("addc.cc", mid[i], mid[i], zm[i])
    p.ins("addc", mid[N], mid[N], "sn")                  # sign extension
    # tt = z0 + mid * 2^128 + z2 * 2^256
    p.chain(); flags.append(False)
    p.ins("add.cc", tt[H], z0[H], mid[0])
    for i in range(1, H):
        p.ins("addc.cc", tt[H + i], z0[H + i], mid[i])
    for i in range(H):
        p.ins("addc.cc", tt[N + i], z2[i], mid[H + i])
    p.ins("addc.cc", tt[N + H], z2[H], mid[N])
    for i in range(H + 1, N):
        p.ins("addc.cc" if i < N - 1 else "addc", tt[N + i], z2[i], 0)
    # Montgomery reduction of the low half: ev = tt[0..8) (tt[0..H) = z0[0..H)), eight 'multiply by one' rows
    ev = ["ev[%d]" % i for i in range(N)]
    od = ["od[%d]" % i for i in range(N)]
    p.chain(); flags.append(False)
    for i in range(N):
        p.ins("add", ev[i], z0[i] if i < H else tt[i], 0)
    for row in range(N):
        e, o = (ev, od) if row % 2 == 0 else (od, ev)
        if row == 0:
            p.chain(); flags.append(False)
            p.ins("mul.lo", "mi", e[0], m0)
            for j in range(0, N, 2):
                p.ins("mul.lo", o[j], M[j + 1], "mi")
                p.ins("mul.hi", o[j + 1], M[j + 1], "mi")
        else:
            p.chain(); flags.append(False)
            p.ins("add.cc", e[0], e[0], o[1])
            p.ins("mul.lo", "mi", e[0], m0)
            for j in range(0, N - 2, 2):
                p.ins("madc.lo.cc", o[j], M[j + 1], "mi", o[j + 2])
                p.ins("madc.hi.cc", o[j + 1], M[j + 1], "mi", o[j + 3])
            p.ins("madc.lo.cc", o[N - 2], M[N - 1], "mi", 0)
            p.ins("madc.hi", o[N - 1], M[N - 1], "mi", 0)
        p.chain(); flags.append(False)
        p.ins("mad.lo.cc", e[0], M[0], "mi", e[0])
        p.ins("madc.hi.cc", e[1], M[0], "mi", e[1])
        for j in range(2, N, 2):
            p.ins("madc.lo.cc", e[j], M[j], "mi", e[j])
            p.ins("madc.hi.cc", e[j + 1], M[j], "mi", e[j + 1])
        p.ins("addc", o[N - 1], o[N - 1], 0)
    # merge: ev[i] += od[i+1]; then add the high half of the product
    p.chain(); flags.append(False)
    p.ins("add.cc", ev[0], ev[0], od[1])
    for i in range(1, N - 1):
        p.ins("addc.cc", ev[i], ev[i], od[i + 1])
    p.ins("addc", ev[N - 1], ev[N - 1], 0)
    p.chain(); flags.append(False)
    p.ins("add.cc", ev[0], ev[0], tt[N])
    for i in range(1, N):
        p.ins("addc.cc" if i < N - 1 else "addc", ev[i], ev[i], tt[N + i])
    # conditional subtract
    p.chain(); flags.append(False)
    p.ins("sub.cc", "t[0]", ev[0], M[0])
    for i in range(1, N):
        p.ins("subc.cc", "t[%d]" % i, ev[i], M[i])
    p.ins("subc", "bw", 0, 0)
    return list(zip(p.chains, flags))


# Measured on B200 (tools/imad_peak.cu): the Karatsuba variant reaches 67.4 G mul/s against 68.7 G mul/s for the
# interleaved schoolbook product — ptxas leaves some lo/hi pairs unfused and moves carry adds onto the IMAD pipe
# (IMAD.X / IMAD.MOV), which eats the 16 saved wide multiplies.  Kept as an option (ZK_FP_KARATSUBA=1), off by default.
KARATSUBA = os.environ.get("ZK_FP_KARATSUBA", "0") != "0"
# Triangular squaring (36 instead of 64 wide products in the a*a part; see gen_mont_mul): 92 instead of 120 IMAD.WIDE per squaring in SASS,
# bucket accumulation (8M + 2S per mixed addition) 1008 -> 987 ms per 1024 proofs on B200.  ZK_FP_TRISQR=0 emits sqr as mul(a, a).
TRISQR = os.environ.get("ZK_FP_TRISQR", "1") != "0"


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
    mulgen = gen_mont_mul_karatsuba if KARATSUBA else gen_mont_mul
    sqrgen = (lambda: gen_mont_mul(mod, sqr=True, tri=True)) if (TRISQR and not KARATSUBA) else (lambda: mulgen(mod, sqr=True))
    for fn, gen in (("mul", lambda: mulgen(mod)), ("sqr", sqrgen)):
        sig = "ZK_FP_FN void %s_%s(uint32_t* __restrict__ r, const uint32_t* __restrict__ a%s)" % (
            name, fn, ", const uint32_t* __restrict__ b" if fn == "mul" else "")
        chains = gen()
        o.append(sig + " {")
        o.append("    uint32_t ev[8] = {0,0,0,0,0,0,0,0}, od[8] = {0,0,0,0,0,0,0,0}, t[8] = {0,0,0,0,0,0,0,0}, mi = 0, bw = 0;")
        if fn == "sqr" and TRISQR and not KARATSUBA:
            o.append("    uint32_t d[8] = {0,0,0,0,0,0,0,0}, dm[8] = {0,0,0,0,0,0,0,0};")
        if KARATSUBA:
            o.append("    uint32_t z0[8] = {0}, z2[8] = {0}, zm[8] = {0}, pe0[8] = {0}, po0[8] = {0}, pe1[8] = {0}, po1[8] = {0}, pe2[8] = {0}, po2[8] = {0};")
            o.append("    uint32_t da[4] = {0}, db[4] = {0}, mid[9] = {0}, tt[16] = {0}, sa = 0, sb = 0, sn = 0, cz = 0; (void)cz;")
        o.append("#ifdef __CUDA_ARCH__")
        o.append(body(chains, True))
        o.append("#else")
        o.append(body(chains, False))
        o.append("#endif")
        o.append("    for (int i = 0; i < 8; ++i) r[i] = bw ? ev[i] : t[i];")
        o.append("}")
    # r = a*b + c*d with one reduction
    o.append("ZK_FP_FN void %s_mul_add2(uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, "
             "const uint32_t* __restrict__ c, const uint32_t* __restrict__ d) {" % name)
    o.append("    uint32_t ev[8] = {0,0,0,0,0,0,0,0}, od[8] = {0,0,0,0,0,0,0,0}, t[8] = {0,0,0,0,0,0,0,0}, mi = 0, bw = 0;")
    chains = gen_mont_mul(mod, dual=True)
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
    with open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "fp_gen.inc"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
