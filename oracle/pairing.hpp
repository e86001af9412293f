// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
//
// BN254 pairing check for the proof-acceptance gate.  The in-repo verifier ends in the EVM pairing
// precompile 0x08 (/root/reference/crates/halo2-verifier/templates/Halo2Verifier.sol:204-219,
// 552-559): accept iff e(LHS, G2) * e(RHS, -s*G2) == 1.  Any bilinear non-degenerate pairing on
// (G1, G2) decides that equation identically, so this oracle uses the plain ate pairing
// f_{t-1,Q}(P)^((q^12-1)/r) (Miller loop over t-1 = 6x^2, affine twist coordinates, no Frobenius
// end-steps) — simple enough to audit, and pinned by the SRS fixture itself:
// e(g[1], g2) == e(g[0], s_g2) on resources/ppot_0080_11_raw (tests/test_oracle_kat.py).
// Tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9+u, Fq12 = Fq6[w]/(w^2 - v).
// G2 points live on the D-twist y^2 = x^3 + 3/xi; untwist (x',y') -> (x' w^2, y' w^3).
#pragma once
#include "bn254.hpp"
#include "misc.hpp"
#include "arith.hpp"
#include "final_exp.inc"

namespace oracle {

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 operator-() const { return {-c0, -c1}; }
    Fq2 operator*(const Fq2& o) const { return {c0 * o.c0 - c1 * o.c1, c0 * o.c1 + c1 * o.c0}; }
    Fq2 scale(const Fq& s) const { return {c0 * s, c1 * s}; }
    Fq2 square() const { return *this * *this; }
    Fq2 mul_xi() const {  // * (9 + u)
        Fq n9 = Fq::from_u64(9);
        return {c0 * n9 - c1, c0 + c1 * n9};
    }
    Fq2 inv() const { Fq t = (c0.square() + c1.square()).inv(); return {c0 * t, -(c1 * t)}; }
};

struct Fq6 {
    Fq2 a0, a1, a2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool operator==(const Fq6& o) const { return a0 == o.a0 && a1 == o.a1 && a2 == o.a2; }
    Fq6 operator+(const Fq6& o) const { return {a0 + o.a0, a1 + o.a1, a2 + o.a2}; }
    Fq6 operator-(const Fq6& o) const { return {a0 - o.a0, a1 - o.a1, a2 - o.a2}; }
    Fq6 operator-() const { return {-a0, -a1, -a2}; }
    Fq6 operator*(const Fq6& o) const {
        Fq2 t0 = a0 * o.a0 + (a1 * o.a2 + a2 * o.a1).mul_xi();
        Fq2 t1 = a0 * o.a1 + a1 * o.a0 + (a2 * o.a2).mul_xi();
        Fq2 t2 = a0 * o.a2 + a1 * o.a1 + a2 * o.a0;
        return {t0, t1, t2};
    }
    Fq6 mul_v() const { return {a2.mul_xi(), a0, a1}; }
    Fq6 inv() const {
        Fq2 c0 = a0.square() - (a1 * a2).mul_xi();
        Fq2 c1 = a2.square().mul_xi() - a0 * a1;
        Fq2 c2 = a1.square() - a0 * a2;
        Fq2 t = (a0 * c0 + (a2 * c1 + a1 * c2).mul_xi()).inv();
        return {c0 * t, c1 * t, c2 * t};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 operator*(const Fq12& o) const {
        return {c0 * o.c0 + (c1 * o.c1).mul_v(), c0 * o.c1 + c1 * o.c0};
    }
    Fq12 square() const { return *this * *this; }
    Fq12 conj() const { return {c0, -c1}; }
    Fq12 inv() const {
        Fq6 t = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * t, -(c1 * t)};
    }
};

static inline Fq2 g2_twist_b() {  // 3 / (9 + u)
    static const Fq2 b = Fq2{Fq::from_u64(9), Fq::one()}.inv().scale(Fq::from_u64(3));
    return b;
}
static inline bool g2_on_curve(const G2AffineRaw& p) {
    Fq2 x{p.x0, p.x1}, y{p.y0, p.y1};
    if (x.is_zero() && y.is_zero()) return true;
    return y.square() == x.square() * x + g2_twist_b();
}

// line through untwisted T (and slope lam' on the twist) evaluated at P:
//   yP - lam' xP w + (lam' xT - yT) w^3      with w^3 = v*w
static inline Fq12 line_eval(const Fq2& lam, const Fq2& xt, const Fq2& yt, const G1Affine& p) {
    Fq12 l;
    l.c0 = {Fq2{p.y, Fq::zero()}, Fq2::zero(), Fq2::zero()};
    l.c1 = {-(lam.scale(p.x)), lam * xt - yt, Fq2::zero()};
    return l;
}

// Miller function f_{t-1,Q}(P)
static inline Fq12 miller_ate(const G1Affine& p, const G2AffineRaw& q) {
    if (p.is_identity()) return Fq12::one();
    Fq2 qx{q.x0, q.x1}, qy{q.y0, q.y1};
    if (qx.is_zero() && qy.is_zero()) return Fq12::one();
    // t - 1 = 6 x^2, x = 4965661367192848881
    const u128 T = ((u128)0x6f4d8248eeb859fbULL << 64) | 0xf83e9682e87cfd46ULL;
    Fq2 rx = qx, ry = qy;
    Fq12 f = Fq12::one();
    int top = 127; while (!((T >> top) & 1)) --top;
    Fq2 three{Fq::from_u64(3), Fq::zero()};
    for (int i = top - 1; i >= 0; --i) {
        Fq2 lam = (rx.square() * three) * (ry + ry).inv();
        f = f.square() * line_eval(lam, rx, ry, p);
        Fq2 nx = lam.square() - rx - rx;
        ry = lam * (rx - nx) - ry; rx = nx;
        if ((T >> i) & 1) {
            Fq2 lam2 = (qy - ry) * (qx - rx).inv();
            f = f * line_eval(lam2, rx, ry, p);
            Fq2 mx = lam2.square() - rx - qx;
            ry = lam2 * (rx - mx) - ry; rx = mx;
        }
    }
    return f;
}

static inline Fq12 final_exponentiation(const Fq12& f) {
    Fq12 g = f.conj() * f.inv();  // f^(q^6 - 1)
    // g^((q^6+1)/r), MSB-first over the hex constant
    Fq12 acc = Fq12::one();
    for (const char* c = FINAL_EXP_HARD_HEX; *c; ++c) {
        int d = *c <= '9' ? *c - '0' : *c - 'a' + 10;
        for (int b = 3; b >= 0; --b) { acc = acc.square(); if ((d >> b) & 1) acc = acc * g; }
    }
    return acc;
}

// k * Q on the twist (affine double-and-add; used once per synthetic SRS for s_g2 = s * g2)
static inline G2AffineRaw g2_mul(const G2AffineRaw& q, const Fr& k) {
    U256 e = k.to_u256();
    bool inf = true;
    Fq2 rx = Fq2::zero(), ry = Fq2::zero();
    const Fq2 qx{q.x0, q.x1}, qy{q.y0, q.y1};
    const Fq2 three{Fq::from_u64(3), Fq::zero()};
    for (int i = 255; i >= 0; --i) {
        if (!inf) {
            if (ry.is_zero()) inf = true;
            else {
                Fq2 lam = (rx.square() * three) * (ry + ry).inv();
                Fq2 nx = lam.square() - rx - rx;
                ry = lam * (rx - nx) - ry; rx = nx;
            }
        }
        if ((e.l[i / 64] >> (i % 64)) & 1) {
            if (inf) { rx = qx; ry = qy; inf = false; }
            else if (rx == qx) {
                if (ry == qy) { Fq2 lam = (rx.square() * three) * (ry + ry).inv(); Fq2 nx = lam.square() - rx - rx; ry = lam * (rx - nx) - ry; rx = nx; }
                else inf = true;
            } else {
                Fq2 lam = (qy - ry) * (qx - rx).inv();
                Fq2 nx = lam.square() - rx - qx;
                ry = lam * (rx - nx) - ry; rx = nx;
            }
        }
    }
    if (inf) return {Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()};
    return {rx.c0, rx.c1, ry.c0, ry.c1};
}

// `ParamsKZG::setup(k, rng)` (halo2_proofs poly/kzg/commitment.rs [UPSTREAM-MEMORY, SURVEY Appendix A]), the
// SRS the reference's seeded tests build (/root/reference/crates/halo2-verifier/src/generator.rs:118-119):
// s = Fr::random(rng); g[i] = G * s^i; g_lagrange[i] = G * ((s^n - 1)/n * w^i / (s - w^i)); s_g2 = s * g2.
static inline Srs params_setup(unsigned k, RngCore& rng, const G2AffineRaw& g2_generator, unsigned threads) {
    Srs out; out.k = k;
    const size_t n = (size_t)1 << k;
    Fr s = random_field<Fr>(rng);
    EvaluationDomain d(2, k);
    std::vector<G1> gp(n), glp(n);
    const G1 G = G1::from_affine(G1Affine::generator());
    Fr sn_m1_over_n = (s.pow_u64(n) - Fr::one()) * Fr::from_u64(n).inv();
    std::vector<Fr> den(n);
    { Fr w = Fr::one(); for (size_t i = 0; i < n; ++i) { den[i] = s - w; w = w * d.omega; } }
    batch_invert(den.data(), n);
    parallel_chunks(n, threads, [&](size_t a, size_t b) {
        Fr sp = s.pow_u64(a), w = d.omega.pow_u64(a);
        for (size_t i = a; i < b; ++i) {
            gp[i] = G.mul(sp);
            glp[i] = G.mul(sn_m1_over_n * w * den[i]);
            sp = sp * s; w = w * d.omega;
        }
    });
    out.g.resize(n); out.g_lagrange.resize(n);
    batch_normalize(gp.data(), out.g.data(), n);
    batch_normalize(glp.data(), out.g_lagrange.data(), n);
    out.g2 = g2_generator;
    out.s_g2 = g2_mul(g2_generator, s);
    return out;
}

// e(p1,q1) * e(p2,q2) == 1 ?
static inline bool pairing_product_is_one(const G1Affine& p1, const G2AffineRaw& q1, const G1Affine& p2, const G2AffineRaw& q2) {
    Fq12 f = miller_ate(p1, q1) * miller_ate(p2, q2);
    return final_exponentiation(f) == Fq12::one();
}

}  // namespace oracle
