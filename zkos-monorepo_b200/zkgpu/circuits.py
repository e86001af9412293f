"""Synthetic "Shielder-shaped" PLONKish circuits.

The real NewAccount / Deposit / Withdraw circuits live in the un-vendored zkOS-circuits repository
and cannot be reproduced here, so configs 1/4/5 run on shape-equivalent synthetic circuits: a family
of arithmetic units (q_m*a*b + q_l*a + q_r*b + q_o*c + q_c + q_pi*PI = 0) chained by copy constraints
and Poseidon-style power-5 units (q_pow * ((x + rc)^5 + y - x_next) = 0, degree 6 => 5 quotient
pieces, extended domain 2^(k+3)), one instance column that takes part in the permutation argument,
and optionally range-check style lookup arguments (`n_lookup`: lookup 0 is q_lk*v in {0..T-1}; lookup 1 compresses two
expressions, (q_lk*v, q_lk*w) in {(i, 3i+1)} u {(0,0)}, so theta is exercised).  This module builds the constraint-system blob both libzkgpu and the CPU oracle parse,
and satisfying witnesses.  Field arithmetic is delegated to an injected backend (vectorised
mul/add/sub over (n,4) uint64 Montgomery arrays): the tests inject the CPU oracle, bench.py injects
the GPU library, so this file never touches the oracle itself.
"""
import struct

import numpy as np

OP_CONST, OP_FIXED, OP_ADVICE, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE = range(8)
COL_ADVICE, COL_FIXED, COL_INSTANCE = 0, 1, 2
MAGIC = 0x5A4B4353
CS_ONLY_MAGIC = 0x5A4B4354   # constraint system only: what the Rust exporter of INTEGRATION.md writes next to a pk.bin

# name -> (k, arithmetic units, pow5 units, public inputs)   (public-input counts: SURVEY §8d config 5;
# /root/reference/contracts/Shielder.sol:504-519, 679-701)
SHAPES = {
    "new_account": dict(k=12, n_arith=4, n_pow=2, num_pi=13),
    "deposit": dict(k=13, n_arith=5, n_pow=3, num_pi=8),
    "withdraw": dict(k=13, n_arith=6, n_pow=3, num_pi=8),
    # small shapes for fast CPU tests
    "tiny": dict(k=6, n_arith=2, n_pow=1, num_pi=3),
    "small": dict(k=9, n_arith=3, n_pow=2, num_pi=8),
    # the same families with lookup arguments
    "tiny_lookup": dict(k=6, n_arith=2, n_pow=1, num_pi=3, n_lookup=2, table_size=16),
    "small_lookup": dict(k=9, n_arith=3, n_pow=0, num_pi=8, n_lookup=1, table_size=64),
    "withdraw_lookup": dict(k=13, n_arith=6, n_pow=3, num_pi=8, n_lookup=2, table_size=256),
}

FIXED_NAMES = ["q_m", "q_l", "q_r", "q_o", "q_c", "q_pi", "q_pow", "rc"]


class Shape:
    def __init__(self, name=None, **kw):
        p = dict(SHAPES[name]) if name else {}
        p.update(kw)
        self.name = name or "custom"
        self.k, self.n_arith, self.n_pow, self.num_pi = p["k"], p["n_arith"], p["n_pow"], p["num_pi"]
        self.n_lookup, self.table_size = p.get("n_lookup", 0), p.get("table_size", 0)
        assert self.n_lookup in (0, 1, 2)
        self.n = 1 << self.k
        self.num_advice = 3 * self.n_arith + 2 * self.n_pow + (0, 1, 3)[self.n_lookup]
        self.fixed_names = FIXED_NAMES + (["q_lk", "t_a", "t_b"] if self.n_lookup else [])
        self.num_fixed = len(self.fixed_names)
        base = 3 * self.n_arith + 2 * self.n_pow
        self.lv = [base, base + 1][: self.n_lookup]      # looked-up value columns
        self.lw = base + 2 if self.n_lookup == 2 else None
        # advice column indices
        self.a = [3 * u for u in range(self.n_arith)]
        self.b = [3 * u + 1 for u in range(self.n_arith)]
        self.c = [3 * u + 2 for u in range(self.n_arith)]
        self.x = [3 * self.n_arith + 2 * v for v in range(self.n_pow)]
        self.y = [3 * self.n_arith + 2 * v + 1 for v in range(self.n_pow)]
        # queries: every advice column at rotation 0 (in column order), then x_v at rotation +1
        self.advice_queries = [(c, 0) for c in range(self.num_advice)] + [(c, 1) for c in self.x]
        self.fixed_queries = [(c, 0) for c in range(self.num_fixed)]
        self.instance_queries = [(0, 0)]
        self.degree = max(6 if self.n_pow else 3, 5 if self.n_lookup else 0)   # lookup: max(4, 2 + deg(q_lk*v) + deg(t))
        max_q = 2 if self.n_pow else 1
        self.blinding_factors = max(3, max_q) + 2
        self.usable = self.n - (self.blinding_factors + 1)
        self.chunk_len = self.degree - 2
        # permutation columns: all advice columns, then the instance column
        self.perm_columns = [(COL_ADVICE, c) for c in range(self.num_advice)] + [(COL_INSTANCE, 0)]
        self.num_perm_sets = -(-len(self.perm_columns) // self.chunk_len)
        self.num_quotients = self.degree - 1
        self.extended_k = self.k
        while (1 << self.extended_k) < self.n * (self.degree - 1):
            self.extended_k += 1
        L = self.n_lookup
        self.num_evals = (len(self.advice_queries) + len(self.fixed_queries) + 1 + len(self.perm_columns)
                          + 3 * self.num_perm_sets - 1 + 5 * L)
        self.proof_len = 64 * (self.num_advice + 3 * L + self.num_perm_sets + 1 + self.num_quotients) + 32 * self.num_evals + 128
        self.num_msm = self.num_advice + 3 * L + self.num_perm_sets + self.num_quotients + 3
        self.num_ntt = 1 + self.num_advice + 3 * L + self.num_perm_sets
        self.num_ext_ntt = self.num_advice + 1 + 3 * L + self.num_perm_sets + 1

    # ---- expressions (postfix) -------------------------------------------------------------
    def lookups(self):
        """[(input expressions, table expressions)] in postfix form"""
        F = {n: i for i, n in enumerate(self.fixed_names)}
        aq = {q: i for i, q in enumerate(self.advice_queries)}
        sel = lambda col: [(OP_FIXED, F["q_lk"]), (OP_ADVICE, aq[(col, 0)]), (OP_MUL, 0)]
        out = []
        if self.n_lookup >= 1:
            out.append(([sel(self.lv[0])], [[(OP_FIXED, F["t_a"])]]))
        if self.n_lookup == 2:
            out.append(([sel(self.lv[1]), sel(self.lw)], [[(OP_FIXED, F["t_a"])], [(OP_FIXED, F["t_b"])]]))
        return out

    def gates(self):
        F = {n: i for i, n in enumerate(self.fixed_names)}
        aq = {q: i for i, q in enumerate(self.advice_queries)}
        gates = []
        for u in range(self.n_arith):
            a, b, c = aq[(self.a[u], 0)], aq[(self.b[u], 0)], aq[(self.c[u], 0)]
            e = [(OP_FIXED, F["q_m"]), (OP_ADVICE, a), (OP_MUL, 0), (OP_ADVICE, b), (OP_MUL, 0),
                 (OP_FIXED, F["q_l"]), (OP_ADVICE, a), (OP_MUL, 0), (OP_ADD, 0),
                 (OP_FIXED, F["q_r"]), (OP_ADVICE, b), (OP_MUL, 0), (OP_ADD, 0),
                 (OP_FIXED, F["q_o"]), (OP_ADVICE, c), (OP_MUL, 0), (OP_ADD, 0),
                 (OP_FIXED, F["q_c"]), (OP_ADD, 0)]
            if u == 0:
                e += [(OP_FIXED, F["q_pi"]), (OP_INSTANCE, 0), (OP_MUL, 0), (OP_ADD, 0)]
            gates.append(e)
        for v in range(self.n_pow):
            x, y, xn = aq[(self.x[v], 0)], aq[(self.y[v], 0)], aq[(self.x[v], 1)]
            t = [(OP_ADVICE, x), (OP_FIXED, F["rc"]), (OP_ADD, 0)]          # t = x + rc
            e = t + t + [(OP_MUL, 0)]                                       # t^2
            e = e + e + [(OP_MUL, 0)] + t + [(OP_MUL, 0)]                   # t^4 * t
            e += [(OP_ADVICE, y), (OP_ADD, 0), (OP_ADVICE, xn), (OP_NEG, 0), (OP_ADD, 0)]
            e = [(OP_FIXED, F["q_pow"])] + e + [(OP_MUL, 0)]
            gates.append(e)
        return gates


class Circuit:
    """Fixed assignment + copy constraints + serialised blob for one shape (seeded)."""

    def __init__(self, shape, backend, seed=1):
        self.shape, self.F = shape, backend
        s, F = shape, backend
        n, usable = s.n, s.usable
        rnd = F.random(seed, 6 * n)
        zero = np.zeros((n, 4), dtype=np.uint64)
        active = np.zeros(n, dtype=bool)
        active[:usable] = True
        fx = {}
        for i, name in enumerate(["q_m", "q_l", "q_r", "q_c", "rc"]):
            col = rnd[i * n:(i + 1) * n].copy()
            col[~active] = 0
            fx[name] = col
        minus_one = F.const(-1)
        one = F.const(1)
        fx["q_o"] = np.where(active[:, None], minus_one[None, :], zero)
        q_pi = zero.copy()
        q_pi[: s.num_pi] = one
        fx["q_pi"] = q_pi
        q_pow = zero.copy()
        even = np.arange(n) % 2 == 0
        q_pow[even & (np.arange(n) + 1 < usable)] = one
        fx["q_pow"] = q_pow
        if s.n_lookup:
            T = s.table_size
            assert T < usable
            fx["q_lk"] = np.where(active[:, None], one[None, :], zero)
            idx = np.zeros((n, 4), dtype=np.uint64)
            idx[:T, 0] = np.arange(T, dtype=np.uint64)
            tb = np.zeros((n, 4), dtype=np.uint64)
            tb[:T, 0] = 3 * np.arange(T, dtype=np.uint64) + 1
            fx["t_a"], fx["t_b"] = F.to_mont(idx), F.to_mont(tb)
        self.fixed = np.stack([fx[nm] for nm in s.fixed_names])  # (F, n, 4)
        # copy constraints (indices into perm_columns: advice column c -> c, instance -> num_advice)
        copies = []
        for u in range(1, s.n_arith):
            copies += [(s.c[u - 1], r, s.a[u], r) for r in range(usable)]
        inst = s.num_advice
        if s.n_pow:
            copies += [(inst, i, s.y[0], i) for i in range(s.num_pi)]
        else:
            copies += [(inst, i, s.b[0], i) for i in range(s.num_pi)]
        self.copies = copies
        self.blob = self._serialize()

    def cs_blob(self, transcript_repr, num_selectors=0):
        """The constraint system alone (no fixed assignment, no copy constraints — a pk.bin holds what is derived from those), followed by
        the number of selector bit-vectors in the pk.bin's verifying-key section and `vk.transcript_repr()` (4 Montgomery limbs)."""
        return self._serialize(cs_only=(np.ascontiguousarray(transcript_repr, dtype=np.uint64), num_selectors))

    def _serialize(self, cs_only=None):
        s = self.shape
        out = [struct.pack("<5I", CS_ONLY_MAGIC if cs_only else MAGIC, s.k, s.num_fixed, s.num_advice, 1)]
        for qs in (s.advice_queries, s.fixed_queries, s.instance_queries):
            out.append(struct.pack("<I", len(qs)))
            for c, r in qs:
                out.append(struct.pack("<Ii", c, r))
        out.append(struct.pack("<I", 0))  # constants
        gates = s.gates()
        out.append(struct.pack("<I", len(gates)))
        for g in gates:
            out.append(struct.pack("<I", len(g)))
            out.append(np.array(g, dtype=np.uint32).tobytes())
        out.append(struct.pack("<I", len(s.perm_columns)))
        for t, i in s.perm_columns:
            out.append(struct.pack("<II", t, i))
        lks = s.lookups()
        out.append(struct.pack("<I", len(lks)))
        for inputs, tables in lks:
            for exprs in (inputs, tables):
                out.append(struct.pack("<I", len(exprs)))
                for e in exprs:
                    out.append(struct.pack("<I", len(e)))
                    out.append(np.array(e, dtype=np.uint32).tobytes())
        if cs_only:
            out.append(struct.pack("<I", cs_only[1]))
            out.append(cs_only[0].tobytes())
            return b"".join(out)
        out.append(np.ascontiguousarray(self.fixed).tobytes())
        out.append(struct.pack("<I", len(self.copies)))
        out.append(np.array(self.copies, dtype=np.uint32).tobytes())
        return b"".join(out)

    # ---- witness ---------------------------------------------------------------------------
    def witness(self, seed):
        """Returns (advice (A, n, 4) uint64 Montgomery, instance (num_pi, 4)) satisfying the circuit."""
        s, F = self.shape, self.F
        n, usable = s.n, s.usable
        fx = {nm: self.fixed[i] for i, nm in enumerate(s.fixed_names)}
        need = s.n_arith + 1 + 2 * s.n_pow
        rnd = F.random(0x9E3779B9 * (seed + 1) & 0xFFFFFFFFFFFF, need * n + s.num_pi)
        take = iter(range(need))
        col = lambda: rnd[next(take) * n:][:n].copy()
        pi = rnd[need * n:][: s.num_pi].copy()
        inst = np.zeros((n, 4), dtype=np.uint64)
        inst[: s.num_pi] = pi
        adv = np.zeros((s.num_advice, n, 4), dtype=np.uint64)
        # the cell tied to the instance column must be set before it is used
        pi_cell_col = s.y[0] if s.n_pow else s.b[0]
        prev_c = None
        for u in range(s.n_arith):
            a = col() if u == 0 else prev_c
            b = col()
            if u == 0 and pi_cell_col == s.b[0]:
                b[: s.num_pi] = pi
            c = F.add(F.add(F.add(F.mul(fx["q_m"], F.mul(a, b)), F.mul(fx["q_l"], a)), F.mul(fx["q_r"], b)), fx["q_c"])
            if u == 0:
                c = F.add(c, F.mul(fx["q_pi"], inst))
            adv[s.a[u]], adv[s.b[u]], adv[s.c[u]] = a, b, c
            prev_c = c
        even = np.arange(n) % 2 == 0
        for v in range(s.n_pow):
            x, y = col(), col()
            if v == 0:
                y[: s.num_pi] = pi
            t = F.add(x, fx["rc"])
            t2 = F.mul(t, t)
            nxt = F.add(F.mul(F.mul(t2, t2), t), y)  # value required at row+1 wherever q_pow = 1
            x[1:][even[:-1]] = nxt[:-1][even[:-1]]
            adv[s.x[v]], adv[s.y[v]] = x, y
        if s.n_lookup:
            lrng = np.random.default_rng(seed * 7919 + 13)
            for l, col in enumerate(s.lv):
                v = np.zeros((n, 4), dtype=np.uint64)
                v[:usable, 0] = lrng.integers(0, s.table_size, usable, dtype=np.uint64)
                adv[col] = F.to_mont(v)
                if l == 1:
                    w = np.zeros((n, 4), dtype=np.uint64)
                    w[:usable, 0] = 3 * v[:usable, 0] + 1
                    adv[s.lw] = F.to_mont(w)
        adv[:, usable:] = 0  # unusable rows are overwritten with blinding by the prover
        return adv, pi
