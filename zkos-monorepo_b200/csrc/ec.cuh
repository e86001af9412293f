// BN254 G1 (y^2 = x^3 + 3 over Fq) on the device.
//   affine  : (x, y) Montgomery, identity = (0,0)  — the halo2curves `G1Affine` memory layout
//             (crates/powers-of-tau/lib.rs:190-231 reads this layout)
//   XYZZ    : (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity has ZZ = 0.
//             Used for bucket accumulation: mixed add 8M+2S, add 12M+2S, double 6M+3S; the two products of every
//             Y3 = R (Q - X3) - Y1 PPP share one Montgomery reduction (fe_mul_add2).
// Every exceptional case (identity operands, P+P, P-P) is handled, so results are exact group
// elements regardless of accumulation order — required for bit-exact parity after normalisation.
#pragma once
#include "fp.cuh"

namespace zk {

struct __align__(16) g1_affine_t {
    fq_t x, y;
    __host__ __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct __align__(16) g1_xyzz_t {
    fq_t x, y, zz, zzz;
    __host__ __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
    __host__ __device__ __forceinline__ static g1_xyzz_t identity() {
        g1_xyzz_t r;
        r.x = fq_t::zero(); r.y = fq_t::zero(); r.zz = fq_t::zero(); r.zzz = fq_t::zero();
        return r;
    }
    __host__ __device__ __forceinline__ static g1_xyzz_t from_affine(const g1_affine_t& a) {
        g1_xyzz_t r;
        if (a.is_identity()) return identity();
        r.x = a.x; r.y = a.y; r.zz = fe_one<FqTag>(); r.zzz = fe_one<FqTag>();
        return r;
    }
};

// dbl-2008-s-1 (a = 0)
__host__ __device__ inline g1_xyzz_t xyzz_dbl(const g1_xyzz_t& p) {
    if (p.is_identity()) return p;
    fq_t U = dbl(p.y), V = sqr(U), W = U * V, S = p.x * V;
    fq_t X2 = sqr(p.x), M = dbl(X2) + X2;
    g1_xyzz_t r;
    r.x = sqr(M) - dbl(S);
    r.y = fe_mul_add2(M, S - r.x, neg(W), p.y);
    r.zz = V * p.zz;
    r.zzz = W * p.zzz;
    return r;
}
// mdbl-2008-s-1: doubling of an affine point
__host__ __device__ inline g1_xyzz_t xyzz_dbl_affine(const g1_affine_t& p) {
    if (p.is_identity()) return g1_xyzz_t::identity();
    fq_t U = dbl(p.y), V = sqr(U), W = U * V, S = p.x * V;
    fq_t X2 = sqr(p.x), M = dbl(X2) + X2;
    g1_xyzz_t r;
    r.x = sqr(M) - dbl(S);
    r.y = fe_mul_add2(M, S - r.x, neg(W), p.y);
    r.zz = V;
    r.zzz = W;
    return r;
}
// madd-2008-s; `negate` adds -q instead
__host__ __device__ inline void xyzz_madd(g1_xyzz_t& acc, const g1_affine_t& q, bool negate) {
    if (q.is_identity()) return;
    fq_t qy = negate ? neg(q.y) : q.y;
    if (acc.is_identity()) {
        acc.x = q.x; acc.y = qy; acc.zz = fe_one<FqTag>(); acc.zzz = fe_one<FqTag>();
        return;
    }
    fq_t U2 = q.x * acc.zz, S2 = qy * acc.zzz;
    fq_t P = U2 - acc.x, R = S2 - acc.y;
    if (P.is_zero()) {
        if (R.is_zero()) { g1_affine_t t; t.x = q.x; t.y = qy; acc = xyzz_dbl_affine(t); }
        else acc = g1_xyzz_t::identity();
        return;
    }
    fq_t PP = sqr(P), PPP = P * PP, Q = acc.x * PP;
    fq_t X3 = sqr(R) - PPP - dbl(Q);
    acc.y = fe_mul_add2(R, Q - X3, neg(acc.y), PPP);   // R (Q - X3) - Y1 PPP, one reduction
    acc.x = X3;
    acc.zz = acc.zz * PP;
    acc.zzz = acc.zzz * PPP;
}
// add-2008-s
__host__ __device__ inline g1_xyzz_t xyzz_add(const g1_xyzz_t& a, const g1_xyzz_t& b) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    fq_t U1 = a.x * b.zz, U2 = b.x * a.zz, S1 = a.y * b.zzz, S2 = b.y * a.zzz;
    fq_t P = U2 - U1, R = S2 - S1;
    if (P.is_zero()) {
        if (R.is_zero()) return xyzz_dbl(a);
        return g1_xyzz_t::identity();
    }
    fq_t PP = sqr(P), PPP = P * PP, Q = U1 * PP;
    g1_xyzz_t r;
    r.x = sqr(R) - PPP - dbl(Q);
    r.y = fe_mul_add2(R, Q - r.x, neg(S1), PPP);
    r.zz = a.zz * b.zz * PP;
    r.zzz = a.zzz * b.zzz * PPP;
    return r;
}
__host__ __device__ inline g1_xyzz_t xyzz_neg(const g1_xyzz_t& a) {
    g1_xyzz_t r = a; r.y = neg(a.y); return r;
}
// k * p for a small unsigned k (double-and-add, MSB first)
__host__ __device__ inline g1_xyzz_t xyzz_mul_small(const g1_xyzz_t& p, uint32_t k) {
    g1_xyzz_t acc = g1_xyzz_t::identity();
    for (int i = 31; i >= 0; --i) {
        acc = xyzz_dbl(acc);
        if ((k >> i) & 1) acc = xyzz_add(acc, p);
    }
    return acc;
}
// full-width scalar (canonical, 8 LE limbs)
__host__ __device__ inline g1_xyzz_t xyzz_mul(const g1_xyzz_t& p, const uint32_t* k) {
    g1_xyzz_t acc = g1_xyzz_t::identity();
    for (int i = 255; i >= 0; --i) {
        acc = xyzz_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) acc = xyzz_add(acc, p);
    }
    return acc;
}
// exact affine normalisation (one inversion): x = X/ZZ, y = Y/ZZZ = Y * ZZ^-1 ... computed from one
// inverse of ZZZ*ZZ: 1/ZZ = ZZZ * t, 1/ZZZ = ZZ * t with t = 1/(ZZ*ZZZ)
__host__ __device__ inline g1_affine_t xyzz_to_affine(const g1_xyzz_t& p) {
    g1_affine_t r;
    if (p.is_identity()) { r.x = fq_t::zero(); r.y = fq_t::zero(); return r; }
    fq_t t = fe_inv(p.zz * p.zzz);
    r.x = p.x * (p.zzz * t);
    r.y = p.y * (p.zz * t);
    return r;
}

}  // namespace zk
