// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under zkos-monorepo_b200/ may include, link or
// call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / CPU baseline.
//
// CPU restatement of the BN254 arithmetic that Shielder's prover gets from the un-vendored
// dependency halo2curves 0.6.1 (/root/reference/Cargo.lock:2367-2368; workspace pin
// /root/reference/Cargo.toml:37).  The source of that crate is not on this machine, so this file
// restates its *published algorithms* (4x64-bit Montgomery CIOS fields Fr/Fq, short-Weierstrass
// G1 y^2 = x^3 + 3, Jacobian arithmetic) and is pinned numerically by the reference's own
// fixtures (resources/ppot_0080_11.ptau, resources/ppot_0080_11_raw — see tests/test_oracle_kat.py)
// and by the constants the in-repo verifier template carries
// (/root/reference/crates/halo2-verifier/templates/Halo2Verifier.sol:222-223 moduli, :475 delta,
//  :94 generator (1,2) / curve equation).
//
// Memory layout of a field element == Rust `Fr`/`Fq` == halo2 `SerdeFormat::RawBytes`:
// four u64 limbs, little-endian, Montgomery form (value * 2^256 mod p).  Verified against
// ppot_0080_11_raw.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include <array>

namespace oracle {

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ---------------------------------------------------------------------------------------------
// 256-bit helpers
// ---------------------------------------------------------------------------------------------
struct U256 {
    u64 l[4];
    bool operator==(const U256& o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const U256& o) const { return !(*this == o); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool bit(unsigned i) const { return (l[i >> 6] >> (i & 63)) & 1; }
};

static inline int u256_cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; --i) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}
static inline u64 u256_add(U256& r, const U256& a, const U256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 u256_sub(U256& r, const U256& a, const U256& b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}

// ---------------------------------------------------------------------------------------------
// Prime field, 4x64 Montgomery (R = 2^256).  P supplies the modulus limbs.
// Follows the structure of halo2curves' derive/field.rs `field_arithmetic!` (CIOS mul, lazy
// subtract on add) [UPSTREAM-MEMORY]; results are canonical (< p) after every operation, which is
// all that parity needs.
// ---------------------------------------------------------------------------------------------
template <class P>
struct Fp {
    U256 v;  // Montgomery form, always < p

    struct Consts { U256 mod, r, r2, r3; u64 inv; };
    static const Consts& C() {
        static const Consts c = make_consts();
        return c;
    }
    static Consts make_consts() {
        Consts c;
        for (int i = 0; i < 4; ++i) c.mod.l[i] = P::MOD[i];
        // inv = -p^{-1} mod 2^64 (Newton)
        u64 x = 1;
        for (int i = 0; i < 7; ++i) x *= 2 - c.mod.l[0] * x;
        c.inv = (u64)0 - x;
        // R = 2^256 mod p by 256 modular doublings of 1; R2/R3 by further doublings
        U256 t{{1, 0, 0, 0}};
        auto dbl = [&](U256& a) {
            u64 top = a.l[3] >> 63;
            for (int i = 3; i > 0; --i) a.l[i] = (a.l[i] << 1) | (a.l[i - 1] >> 63);
            a.l[0] <<= 1;
            if (top || u256_cmp(a, c.mod) >= 0) { U256 s; u256_sub(s, a, c.mod); a = s; }
        };
        for (int i = 0; i < 256; ++i) dbl(t);
        c.r = t;
        for (int i = 0; i < 256; ++i) dbl(t);
        c.r2 = t;
        for (int i = 0; i < 256; ++i) dbl(t);
        c.r3 = t;
        return c;
    }

    static Fp zero() { Fp r; memset(&r, 0, sizeof r); return r; }
    static Fp one() { Fp r; r.v = C().r; return r; }
    static Fp from_raw_mont(const u64* limbs) { Fp r; memcpy(r.v.l, limbs, 32); return r; }
    // canonical integer (must be < p) -> field
    static Fp from_u256(const U256& x) { Fp a; a.v = x; Fp r2; r2.v = C().r2; return a * r2; }
    static Fp from_u64(u64 x) { U256 t{{x, 0, 0, 0}}; return from_u256(t); }
    // 512-bit little-endian integer (8 limbs) reduced mod p: lo*R2 + hi*R3 (Montgomery products),
    // as halo2curves `from_u512` [UPSTREAM-MEMORY, SURVEY Appendix A].
    static Fp from_u512(const u64* w) {
        Fp lo, hi, r2, r3;
        // the raw 256-bit halves may be >= p; mont_mul tolerates any a < 2^256 with b < p.
        memcpy(lo.v.l, w, 32); memcpy(hi.v.l, w + 4, 32);
        r2.v = C().r2; r3.v = C().r3;
        return mont_mul(lo.v, r2.v) + mont_mul(hi.v, r3.v);
    }
    U256 to_u256() const {  // canonical integer
        U256 one{{1, 0, 0, 0}};
        return mont_mul(v, one).v;
    }
    bool is_zero() const { return v.is_zero(); }
    bool operator==(const Fp& o) const { return v == o.v; }
    bool operator!=(const Fp& o) const { return v != o.v; }

    // CIOS Montgomery product, result fully reduced.  Accepts a < 2^256, b < p.
    static Fp mont_mul(const U256& a, const U256& b) {
        const Consts& c = C();
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 carry = 0;
            for (int j = 0; j < 4; ++j) {
                u128 s = (u128)a.l[j] * b.l[i] + t[j] + carry;
                t[j] = (u64)s; carry = s >> 64;
            }
            u128 s = (u128)t[4] + carry;
            t[4] = (u64)s; t[5] = (u64)(s >> 64);
            u64 m = t[0] * c.inv;
            carry = ((u128)m * c.mod.l[0] + t[0]) >> 64;
            for (int j = 1; j < 4; ++j) {
                u128 s2 = (u128)m * c.mod.l[j] + t[j] + carry;
                t[j - 1] = (u64)s2; carry = s2 >> 64;
            }
            s = (u128)t[4] + carry;
            t[3] = (u64)s;
            t[4] = t[5] + (u64)(s >> 64);
            t[5] = 0;
        }
        Fp r; memcpy(r.v.l, t, 32);
        // t < 2p when a < 2^256? Bound: t < (a*b + m*p)/R < (2^256 p + 2^256 p)/2^256 = 2p. one subtract.
        if (t[4] || u256_cmp(r.v, c.mod) >= 0) { U256 s; u256_sub(s, r.v, c.mod); r.v = s; }
        return r;
    }
    Fp operator*(const Fp& o) const { return mont_mul(v, o.v); }
    Fp square() const { return mont_mul(v, v); }
    Fp operator+(const Fp& o) const {
        Fp r; u64 carry = u256_add(r.v, v, o.v);
        if (carry || u256_cmp(r.v, C().mod) >= 0) { U256 s; u256_sub(s, r.v, C().mod); r.v = s; }
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u64 borrow = u256_sub(r.v, v, o.v);
        if (borrow) { U256 s; u256_add(s, r.v, C().mod); r.v = s; }
        return r;
    }
    Fp operator-() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }
    Fp& operator+=(const Fp& o) { *this = *this + o; return *this; }
    Fp& operator-=(const Fp& o) { *this = *this - o; return *this; }
    Fp& operator*=(const Fp& o) { *this = *this * o; return *this; }

    Fp pow(const U256& e) const {
        Fp acc = one();
        for (int i = 255; i >= 0; --i) { acc = acc.square(); if (e.bit(i)) acc = acc * *this; }
        return acc;
    }
    Fp pow_u64(u64 e) const { U256 t{{e, 0, 0, 0}}; return pow(t); }
    // a^(p-2); zero maps to zero (halo2 `invert().unwrap_or(zero)` behaviour in batch paths)
    Fp inv() const {
        U256 e = C().mod, two{{2, 0, 0, 0}};
        u256_sub(e, e, two);
        return pow(e);
    }
    // canonical little-endian 32 bytes (`to_repr`, crates/type-conversions/lib.rs:40-44)
    void to_bytes_le(uint8_t out[32]) const { U256 c = to_u256(); memcpy(out, c.l, 32); }
    // canonical big-endian 32 bytes (EVM word, crates/halo2-verifier/src/lib/verifier_contract.rs:14-20)
    void to_bytes_be(uint8_t out[32]) const {
        uint8_t le[32]; to_bytes_le(le);
        for (int i = 0; i < 32; ++i) out[i] = le[31 - i];
    }
    static bool from_bytes_le(const uint8_t in[32], Fp& out) {
        U256 x; memcpy(x.l, in, 32);
        if (u256_cmp(x, C().mod) >= 0) return false;
        out = from_u256(x); return true;
    }
    std::string hex() const {
        U256 c = to_u256(); char buf[80];
        snprintf(buf, sizeof buf, "%016lx%016lx%016lx%016lx", c.l[3], c.l[2], c.l[1], c.l[0]);
        return buf;
    }
};

struct FrParams { static constexpr u64 MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}; };
struct FqParams { static constexpr u64 MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}; };
typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// Montgomery batch inversion (halo2 `BatchInvert` / `Curve::batch_normalize` helper); zeros stay zero.
template <class F>
static inline void batch_invert(F* a, size_t n) {
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!a[i].is_zero()) acc = acc * a[i]; }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        F t = acc * pre[i]; acc = acc * a[i]; a[i] = t;
    }
}

// Fr constants of halo2curves bn256::Fr (verified numerically there).
struct FrConst {
    static const unsigned S = 28;
    static Fr generator() { return Fr::from_u64(7); }
    static Fr root_of_unity() {  // 7^((r-1)/2^28)
        static const Fr w = [] {
            U256 e = Fr::C().mod; e.l[0] -= 1;  // r-1
            // shift right by 28
            for (int i = 0; i < 4; ++i) e.l[i] = (e.l[i] >> 28) | (i < 3 ? (e.l[i + 1] << 36) : 0);
            return Fr::from_u64(7).pow(e);
        }();
        return w;
    }
    static Fr delta() {  // 7^(2^28), Halo2Verifier.sol:475
        static const Fr d = [] { Fr t = Fr::from_u64(7); for (int i = 0; i < 28; ++i) t = t.square(); return t; }();
        return d;
    }
    static Fr zeta() {  // cube root of unity used as the extended-domain coset generator
        static const Fr z = [] {
            static const u64 Z[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
            U256 t; memcpy(t.l, Z, 32); return Fr::from_u256(t);
        }();
        return z;
    }
};

// ---------------------------------------------------------------------------------------------
// G1: y^2 = x^3 + 3 over Fq.  Affine identity is (0,0) as in halo2curves (`G1Affine::identity`).
// Jacobian (X,Y,Z): x = X/Z^2, y = Y/Z^3; identity has Z = 0.
// ---------------------------------------------------------------------------------------------
struct G1Affine {
    Fq x, y;
    static G1Affine identity() { return {Fq::zero(), Fq::zero()}; }
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
    bool on_curve() const {
        if (is_identity()) return true;
        return y.square() == x.square() * x + Fq::from_u64(3);
    }
    G1Affine neg() const { return is_identity() ? *this : G1Affine{x, -y}; }
    static G1Affine generator() { return {Fq::from_u64(1), Fq::from_u64(2)}; }
};

struct G1 {
    Fq x, y, z;
    static G1 identity() { return {Fq::zero(), Fq::one(), Fq::zero()}; }
    static G1 from_affine(const G1Affine& a) {
        if (a.is_identity()) return identity();
        return {a.x, a.y, Fq::one()};
    }
    bool is_identity() const { return z.is_zero(); }
    G1 neg() const { return {x, -y, z}; }

    G1 dbl() const {  // dbl-2009-l (a = 0)
        if (is_identity()) return *this;
        Fq A = x.square(), B = y.square(), Cc = B.square();
        Fq D = ((x + B).square() - A - Cc).dbl();
        Fq E = A.dbl() + A, F = E.square();
        Fq X3 = F - D.dbl();
        Fq Y3 = E * (D - X3) - Cc.dbl().dbl().dbl();
        Fq Z3 = (y * z).dbl();
        return {X3, Y3, Z3};
    }
    G1 add(const G1& o) const {  // add-2007-bl
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq Z1Z1 = z.square(), Z2Z2 = o.z.square();
        Fq U1 = x * Z2Z2, U2 = o.x * Z1Z1;
        Fq S1 = y * o.z * Z2Z2, S2 = o.y * z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return identity();
        }
        Fq H = U2 - U1, I = H.dbl().square(), J = H * I, rr = (S2 - S1).dbl(), V = U1 * I;
        Fq X3 = rr.square() - J - V.dbl();
        Fq Y3 = rr * (V - X3) - (S1 * J).dbl();
        Fq Z3 = ((z + o.z).square() - Z1Z1 - Z2Z2) * H;
        return {X3, Y3, Z3};
    }
    G1 add_mixed(const G1Affine& o) const {  // madd-2007-bl
        if (o.is_identity()) return *this;
        if (is_identity()) return from_affine(o);
        Fq Z1Z1 = z.square();
        Fq U2 = o.x * Z1Z1, S2 = o.y * z * Z1Z1;
        if (x == U2) {
            if (y == S2) return dbl();
            return identity();
        }
        Fq H = U2 - x, HH = H.square(), I = HH.dbl().dbl(), J = H * I, rr = (S2 - y).dbl(), V = x * I;
        Fq X3 = rr.square() - J - V.dbl();
        Fq Y3 = rr * (V - X3) - (y * J).dbl();
        Fq Z3 = (z + H).square() - Z1Z1 - HH;
        return {X3, Y3, Z3};
    }
    G1Affine to_affine() const {
        if (is_identity()) return G1Affine::identity();
        Fq zi = z.inv(), zi2 = zi.square();
        return {x * zi2, y * zi2 * zi};
    }
    // scalar given as canonical 256-bit integer; plain double-and-add, MSB first
    G1 mul(const U256& k) const {
        G1 acc = identity();
        for (int i = 255; i >= 0; --i) { acc = acc.dbl(); if (k.bit(i)) acc = acc.add(*this); }
        return acc;
    }
    G1 mul(const Fr& k) const { return mul(k.to_u256()); }
};

// `Curve::batch_normalize`: one shared inversion.
static inline void batch_normalize(const G1* in, G1Affine* out, size_t n) {
    std::vector<Fq> zs(n);
    for (size_t i = 0; i < n; ++i) zs[i] = in[i].z;
    batch_invert(zs.data(), n);
    for (size_t i = 0; i < n; ++i) {
        if (in[i].is_identity()) { out[i] = G1Affine::identity(); continue; }
        Fq zi2 = zs[i].square();
        out[i] = {in[i].x * zi2, in[i].y * zi2 * zs[i]};
    }
}

}  // namespace oracle
