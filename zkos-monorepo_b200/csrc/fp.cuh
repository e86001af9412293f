// BN254 field elements on the device: 8 x u32 limbs, little-endian, Montgomery form (R = 2^256).
// Byte-identical to the host/Rust layout (4 x u64 LE Montgomery — halo2curves bn256::Fr/Fq memory,
// SURVEY.md §8a row a1), so buffers cross the C ABI without any repacking.
// The limb arithmetic itself is generated (gen_fp.py -> fp_gen.inc): PTX mad.lo.cc/madc.hi.cc carry
// chains that ptxas fuses into IMAD.WIDE.U32.X — 128 wide multiply-adds per Montgomery product.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "fp_gen.inc"

namespace zk {

struct FrTag {};
struct FqTag {};

template <class Tag>
struct __align__(16) Fe {
    uint32_t l[8];

    __host__ __device__ __forceinline__ static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = 0;
        return r;
    }
    __host__ __device__ __forceinline__ bool is_zero() const {
        return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0;
    }
    __host__ __device__ __forceinline__ bool operator==(const Fe& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) d |= l[i] ^ o.l[i];
        return d == 0;
    }
    __host__ __device__ __forceinline__ bool operator!=(const Fe& o) const { return !(*this == o); }
};
typedef Fe<FrTag> fr_t;
typedef Fe<FqTag> fq_t;

// ---- per-field dispatch ---------------------------------------------------------------------
__host__ __device__ __forceinline__ fr_t operator*(const fr_t& a, const fr_t& b) { fr_t r; fr_mul(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fr_t operator+(const fr_t& a, const fr_t& b) { fr_t r; fr_add(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fr_t operator-(const fr_t& a, const fr_t& b) { fr_t r; fr_sub(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fr_t sqr(const fr_t& a) { fr_t r; fr_sqr(r.l, a.l); return r; }
__host__ __device__ __forceinline__ fq_t operator*(const fq_t& a, const fq_t& b) { fq_t r; fq_mul(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fq_t operator+(const fq_t& a, const fq_t& b) { fq_t r; fq_add(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fq_t operator-(const fq_t& a, const fq_t& b) { fq_t r; fq_sub(r.l, a.l, b.l); return r; }
__host__ __device__ __forceinline__ fq_t sqr(const fq_t& a) { fq_t r; fq_sqr(r.l, a.l); return r; }

// a*b + c*d with ONE Montgomery reduction (gen_fp.py `dual`): 192 wide multiplies instead of 2 x 128, and no modular addition
__host__ __device__ __forceinline__ fr_t fe_mul_add2(const fr_t& a, const fr_t& b, const fr_t& c, const fr_t& d) { fr_t r; fr_mul_add2(r.l, a.l, b.l, c.l, d.l); return r; }
__host__ __device__ __forceinline__ fq_t fe_mul_add2(const fq_t& a, const fq_t& b, const fq_t& c, const fq_t& d) { fq_t r; fq_mul_add2(r.l, a.l, b.l, c.l, d.l); return r; }

__host__ __device__ __forceinline__ void fe_set_one(fr_t& r) { fr_set_one(r.l); }
__host__ __device__ __forceinline__ void fe_set_one(fq_t& r) { fq_set_one(r.l); }
__host__ __device__ __forceinline__ void fe_set_r2(fr_t& r) { fr_set_r2(r.l); }
__host__ __device__ __forceinline__ void fe_set_r2(fq_t& r) { fq_set_r2(r.l); }
__host__ __device__ __forceinline__ void fe_set_modm2(fr_t& r) { fr_set_modm2(r.l); }
__host__ __device__ __forceinline__ void fe_set_modm2(fq_t& r) { fq_set_modm2(r.l); }

template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> fe_one() { Fe<Tag> r; fe_set_one(r); return r; }
template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> fe_r2() { Fe<Tag> r; fe_set_r2(r); return r; }
template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> neg(const Fe<Tag>& a) {
    return a.is_zero() ? a : Fe<Tag>::zero() - a;
}
template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> dbl(const Fe<Tag>& a) { return a + a; }

// Montgomery -> canonical integer limbs (multiply by the integer 1)
template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> from_mont(const Fe<Tag>& a) {
    Fe<Tag> one = Fe<Tag>::zero();
    one.l[0] = 1;
    return a * one;
}
template <class Tag>
__host__ __device__ __forceinline__ Fe<Tag> to_mont(const Fe<Tag>& a) { return a * fe_r2<Tag>(); }

// Any 256-bit integer -> the same value mod r.  The Montgomery limb code needs operands < p (its dropped
// carries are provably zero only then — asserted by the host emulation in tests/test_fp_emulation.py),
// so raw RNG / hash words pass through here first.  2^256 < 6r: at most five subtractions.
__host__ __device__ __forceinline__ void fr_reduce_raw(fr_t& v) {
    const uint32_t M[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    for (int it = 0; it < 5; ++it) {
        uint32_t t[8];
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint64_t d = (uint64_t)v.l[i] - M[i] - borrow;
            t[i] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        if (borrow) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) v.l[i] = t[i];
    }
}

// a^e for a 256-bit exponent given as 8 u32 limbs (LE); used for inversion (e = p-2)
template <class Tag>
__host__ __device__ inline Fe<Tag> fe_pow(const Fe<Tag>& a, const uint32_t* e) {
    Fe<Tag> acc = fe_one<Tag>();
    for (int i = 255; i >= 0; --i) {
        acc = sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * a;
    }
    return acc;
}
template <class Tag>
__host__ __device__ inline Fe<Tag> fe_inv_pow(const Fe<Tag>& a) {  // a^(p-2); 0 -> 0
    Fe<Tag> e; fe_set_modm2(e);
    return fe_pow(a, e.l);
}

// ---- inversion by the binary extended Euclidean algorithm -------------------------------------------------
// ~2*254 iterations of shifts / conditional subtractions on plain 256-bit integers instead of ~380 dependent
// Montgomery products: about 6x shorter on the latency-critical single-thread paths (point normalisation, batch
// inversion seeds).  Input and output in Montgomery form; 0 -> 0.  Same value as a^(p-2) (checked in the tests).
namespace bininv {
__host__ __device__ __forceinline__ bool is_even(const uint32_t* x) { return (x[0] & 1u) == 0; }
__host__ __device__ __forceinline__ bool is_one(const uint32_t* x) { return x[0] == 1u && (x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0; }
__host__ __device__ __forceinline__ bool geq(const uint32_t* x, const uint32_t* y) {
    for (int i = 7; i >= 0; --i) if (x[i] != y[i]) return x[i] > y[i];
    return true;
}
__host__ __device__ __forceinline__ uint32_t add_n(uint32_t* r, const uint32_t* x, const uint32_t* y) {   // returns carry
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) { c += (uint64_t)x[i] + y[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
__host__ __device__ __forceinline__ void sub_n(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint64_t b = 0;
    for (int i = 0; i < 8; ++i) { uint64_t d = (uint64_t)x[i] - y[i] - b; r[i] = (uint32_t)d; b = (d >> 63) & 1; }
}
__host__ __device__ __forceinline__ void shr1(uint32_t* x, uint32_t top) {
    for (int i = 0; i < 7; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[7] = (x[7] >> 1) | (top << 31);
}
// x <- x / 2 mod p (p odd)
__host__ __device__ __forceinline__ void half_mod(uint32_t* x, const uint32_t* p) {
    uint32_t top = 0;
    if (!is_even(x)) top = add_n(x, x, p);
    shr1(x, top);
}
}  // namespace bininv

template <class Tag>
__host__ __device__ inline Fe<Tag> fe_inv_euclid(const Fe<Tag>& a) {
    if (a.is_zero()) return a;
    Fe<Tag> pm2; fe_set_modm2(pm2);
    uint32_t p[8];
    for (int i = 0; i < 8; ++i) p[i] = pm2.l[i];
    p[0] += 2;   // the moduli end in ...01 / ...47: no carry out of the low limb
    uint32_t u[8], v[8], x1[8], x2[8];
    for (int i = 0; i < 8; ++i) { u[i] = a.l[i]; v[i] = p[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    // invariant: x1 * a == u, x2 * a == v (mod p) with a = the input integer (aR)
    while (!bininv::is_one(u) && !bininv::is_one(v)) {
        while (bininv::is_even(u)) { bininv::shr1(u, 0); bininv::half_mod(x1, p); }
        while (bininv::is_even(v)) { bininv::shr1(v, 0); bininv::half_mod(x2, p); }
        if (bininv::geq(u, v)) {
            bininv::sub_n(u, u, v);
            if (bininv::geq(x1, x2)) bininv::sub_n(x1, x1, x2); else { uint32_t t[8]; bininv::sub_n(t, x2, x1); bininv::sub_n(x1, p, t); }
        } else {
            bininv::sub_n(v, v, u);
            if (bininv::geq(x2, x1)) bininv::sub_n(x2, x2, x1); else { uint32_t t[8]; bininv::sub_n(t, x1, x2); bininv::sub_n(x2, p, t); }
        }
    }
    Fe<Tag> r;   // (aR)^-1 as a plain integer
    for (int i = 0; i < 8; ++i) r.l[i] = bininv::is_one(u) ? x1[i] : x2[i];
    // (aR)^-1 * R^3 / R = a^-1 * R: back in Montgomery form
    Fe<Tag> r2 = fe_r2<Tag>();
    return r * (r2 * r2);
}

// Inversion with uniform control flow (Pornin, "Optimized binary GCD for modular inversion", 2020, with 32-bit words): 34 rounds;
// a round runs 15 binary-GCD steps on 32-bit approximations of (a, b) — their top 17 and low 15 bits — collecting the steps
// into a 2x2 matrix of factors |f|, |g| <= 2^15, then applies the matrix to the full-width (a, b) (exact division by 2^15) and to
// the Bezout pair (u, v) modulo p (Montgomery-style division by 2^15).  After 34 * 15 = 510 >= 2 * 254 - 1 steps a = 0, b = 1 and
// v = input^-1.  No data-dependent branch or loop count, so the 32 lanes of a warp stay converged (the binary Euclid above
// diverges on every step), and it is several times shorter: k_normalize / k_batch_inverse / the single-proof latency path.
// Input and output in Montgomery form; 0 -> 0.  Same value as fe_inv_euclid and a^(p-2) (tests/test_host_logic.py).
template <class Tag>
__host__ __device__ inline Fe<Tag> fe_inv(const Fe<Tag>& x) {
    Fe<Tag> pm2; fe_set_modm2(pm2);
    uint32_t p[8], a[8], b[8], u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pm2.l[i]; a[i] = x.l[i]; u[i] = 0; v[i] = 0; }
    p[0] += 2;   // the moduli end in ...01 / ...47: no carry out of the low limb
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = p[i];
    u[0] = 1;
    // -p^-1 mod 2^15 (Newton on the low limb)
    uint32_t pinv = p[0];
    pinv *= 2u - p[0] * pinv; pinv *= 2u - p[0] * pinv; pinv *= 2u - p[0] * pinv; pinv *= 2u - p[0] * pinv;
    const uint32_t npinv15 = (0u - pinv) & 0x7fffu;
    for (int round = 0; round < 34; ++round) {
        // window = the two words below the top set bit of (a | b), at least words (1, 0)
        uint32_t ah = a[7], al = a[6], bh = b[7], bl = b[6];
#pragma unroll
        for (int i = 6; i >= 1; --i) {
            const bool down = (ah | bh) == 0;
            ah = down ? al : ah; bh = down ? bl : bh;
            al = down ? a[i - 1] : al; bl = down ? b[i - 1] : bl;
        }
#ifdef __CUDA_ARCH__
        const unsigned lz = __clz((int)(ah | bh));
#else
        const unsigned lz = (ah | bh) ? (unsigned)__builtin_clz(ah | bh) : 32u;
#endif
        const uint32_t ta = (uint32_t)(((((uint64_t)ah << 32) | al) << lz) >> 32);
        const uint32_t tb = (uint32_t)(((((uint64_t)bh << 32) | bl) << lz) >> 32);
        uint32_t xa = (ta & 0xffff8000u) | (a[0] & 0x7fffu), xb = (tb & 0xffff8000u) | (b[0] & 0x7fffu);
        int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll
        for (int j = 0; j < 15; ++j) {
            const uint32_t odd = 0u - (xa & 1u);
            const uint32_t sw = odd & (0u - (uint32_t)(xa < xb));
            uint32_t d = (xa ^ xb) & sw; xa ^= d; xb ^= d;
            d = (uint32_t)(f0 ^ f1) & sw; f0 ^= (int32_t)d; f1 ^= (int32_t)d;
            d = (uint32_t)(g0 ^ g1) & sw; g0 ^= (int32_t)d; g1 ^= (int32_t)d;
            xa -= xb & odd; f0 -= f1 & (int32_t)odd; g0 -= g1 & (int32_t)odd;
            xa >>= 1; f1 <<= 1; g1 <<= 1;
        }
        // (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^15, made non-negative (the sign goes into the factors)
        {
            int64_t ca = 0, cb = 0;
            uint32_t na[9], nb[9];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ca += (int64_t)a[i] * f0 + (int64_t)b[i] * g0; na[i] = (uint32_t)ca; ca >>= 32;
                cb += (int64_t)a[i] * f1 + (int64_t)b[i] * g1; nb[i] = (uint32_t)cb; cb >>= 32;
            }
            na[8] = (uint32_t)ca; nb[8] = (uint32_t)cb;
            const uint32_t sa = 0u - (uint32_t)(ca < 0), sb = 0u - (uint32_t)(cb < 0);
            uint32_t ka = sa & 1u, kb = sb & 1u;   // two's complement negation: (x ^ mask) + carry
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t wa = ((na[i] >> 15) | (na[i + 1] << 17)) ^ sa, wb = ((nb[i] >> 15) | (nb[i + 1] << 17)) ^ sb;
                a[i] = wa + ka; ka = (a[i] < wa) ? 1u : 0u;
                b[i] = wb + kb; kb = (b[i] < wb) ? 1u : 0u;
            }
            f0 = (f0 ^ (int32_t)sa) - (int32_t)sa; g0 = (g0 ^ (int32_t)sa) - (int32_t)sa;
            f1 = (f1 ^ (int32_t)sb) - (int32_t)sb; g1 = (g1 ^ (int32_t)sb) - (int32_t)sb;
        }
        // (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^15 mod p: add the multiple of p that clears the low 15 bits, shift, fold into [0, p)
        {
            const uint32_t mu = (((uint32_t)((int32_t)u[0] * f0 + (int32_t)v[0] * g0)) * npinv15) & 0x7fffu;
            const uint32_t mv = (((uint32_t)((int32_t)u[0] * f1 + (int32_t)v[0] * g1)) * npinv15) & 0x7fffu;
            int64_t cu = 0, cv = 0;
            uint32_t nu[9], nv[9];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cu += (int64_t)u[i] * f0 + (int64_t)v[i] * g0 + (int64_t)((uint64_t)p[i] * mu); nu[i] = (uint32_t)cu; cu >>= 32;
                cv += (int64_t)u[i] * f1 + (int64_t)v[i] * g1 + (int64_t)((uint64_t)p[i] * mv); nv[i] = (uint32_t)cv; cv >>= 32;
            }
            nu[8] = (uint32_t)cu; nv[8] = (uint32_t)cv;
            const bool negu = cu < 0, negv = cv < 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { u[i] = (nu[i] >> 15) | (nu[i + 1] << 17); v[i] = (nv[i] >> 15) | (nv[i + 1] << 17); }
            // value in (-p, 2p): negative -> + p (the 256-bit wrap-around is the wanted result); >= p -> - p
            uint32_t t[8];
            if (negu) bininv::add_n(u, u, p); else if (bininv::geq(u, p)) { bininv::sub_n(t, u, p); for (int i = 0; i < 8; ++i) u[i] = t[i]; }
            if (negv) bininv::add_n(v, v, p); else if (bininv::geq(v, p)) { bininv::sub_n(t, v, p); for (int i = 0; i < 8; ++i) v[i] = t[i]; }
        }
    }
    Fe<Tag> r;   // (xR)^-1 as a plain integer; zero input leaves v = 0
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = v[i];
    Fe<Tag> r2 = fe_r2<Tag>();
    return r * (r2 * r2);   // (xR)^-1 * R^3 / R = x^-1 R
}

// 128-bit vector loads/stores (two per element)
template <class Tag>
__device__ __forceinline__ Fe<Tag> fe_load(const Fe<Tag>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fe<Tag> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
#ifdef __CUDACC__
template <class Tag>
__device__ __forceinline__ Fe<Tag> fe_ldg(const Fe<Tag>* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fe<Tag> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
#endif
template <class Tag>
__device__ __forceinline__ void fe_store(Fe<Tag>* p, const Fe<Tag>& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

}  // namespace zk
