"""Pure-Python big-integer reference for BN254 — an *independent* cross-check of the C++ oracle
(tests only).  Nothing here is fast; use at small sizes."""
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
MONT_R = 1 << 256


def to_mont(x, p):
    return (x * MONT_R) % p


def from_mont(x, p):
    return (x * pow(MONT_R, -1, p)) % p


def limbs_to_int(arr):
    """numpy uint64[...,4] -> python ints (flat list)"""
    a = arr.reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def int_to_limbs(vals):
    import numpy as np
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


# --- G1 affine arithmetic over Fq with python ints; identity = None -------------------------
def ec_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    return (x3, (lam * (x1 - x3) - y1) % Q_MOD)


def ec_mul(P, k):
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, P)
        P = ec_add(P, P)
        k >>= 1
    return acc


def ec_msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = ec_add(acc, ec_mul(p, s % R_MOD))
    return acc


def ntt_naive(a, omega, p=R_MOD):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, p) for j in range(n)) % p for i in range(n)]


def ntt_fast(a, omega, p=R_MOD):
    n = len(a)
    if n == 1:
        return list(a)
    ev = ntt_fast(a[0::2], omega * omega % p, p)
    od = ntt_fast(a[1::2], omega * omega % p, p)
    out = [0] * n
    w = 1
    for i in range(n // 2):
        t = w * od[i] % p
        out[i] = (ev[i] + t) % p
        out[i + n // 2] = (ev[i] - t) % p
        w = w * omega % p
    return out


ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23


def omega_for(k):
    return pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)


# ---- Poseidon2, t = 8 (plain-integer restatement of /root/reference/poseidon2-solidity/generate_t8.py) ----
def _p2_params():
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poseidon2_t8.json")))
    return d, [int(c, 16) for c in d["round_constants"]], [int(x, 16) for x in d["diag"]]


def poseidon2_t8(inputs, p=R_MOD):
    """hash of 1..7 integers: zero-padded rate part, capacity element len * 2^64, output state[0]"""
    d, Cs, D = _p2_params()
    RF, RP = d["rounds_f"], d["rounds_p"]
    s = [x % p for x in inputs] + [0] * (7 - len(inputs)) + [len(inputs) << 64]

    def ext(s):
        # generate_t8.py:480-516 written as the matrix it implements (M of the generator, :461-468)
        return [sum(m * x for m, x in zip(row, s)) % p for row in d["M"]]

    s = ext(s)
    for r in range(RF + RP):
        if RF // 2 <= r < RF // 2 + RP:
            s[0] = pow(s[0] + Cs[8 * r], 7, p)
            tot = sum(s) % p
            s = [(D[i] * s[i] + tot) % p for i in range(8)]
        else:
            s = ext([pow(s[i] + Cs[8 * r + i], 7, p) for i in range(8)])
    return s[0]
