// 254-bit Montgomery product on the FP64 pipe (device only): the same function as fr_mul / fq_mul —
//   r = a * b * 2^-256 mod p,   8 x u32 Montgomery limbs in and out, bit-identical results —
// computed with 52-bit limbs held in doubles and DFMA (after Emmart, Zheng, Weems, "Faster modular exponentiation
// using double precision floating point arithmetic on the GPU", ARITH 2018).
//
// Why: ncu shows every hot kernel of the prover saturating the IMAD ("fmaheavy") pipe while the FP64 pipe idles;
// B200 issues 17 T DFMA/s (tools/dfma_peak.cu).  Warps that run this variant next to warps that run the IMAD
// variant use both pipes at once.
//
// Method.  For integers x, y < 2^52 held exactly in doubles:
//     hi = fma_rz(x, y, 2^104)                 = 2^104 + floor(xy / 2^52) * 2^52      (ulp of [2^104, 2^105) is 2^52)
//     lo = fma_rz(x, y, (2^104 + 2^52) - hi)   = 2^52  + (xy mod 2^52)                (exact)
// so the raw bits of hi / lo are (0x467 << 52) | floor(xy / 2^52) and (0x433 << 52) | (xy mod 2^52): the two halves
// of the 104-bit product sit in the mantissas and are accumulated as 64-bit integers; the exponent fields add up
// to constants known at compile time, removed when a column is read.
// a = sum A_i 2^(52 i) (5 limbs), b' = 16 * b (so that five 52-bit reduction digits divide by 2^260 = 16 * 2^256),
// interleaved Montgomery reduction digit by digit with q_i = (t_i * (-p^-1)) mod 2^52, result < 1.19 p, one
// conditional subtraction.
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace zk {
namespace dfma {

constexpr uint64_t M52 = (1ull << 52) - 1;
constexpr uint64_t BL = 0x433ull << 52;   // exponent field of 2^52  (lo halves)
constexpr uint64_t BH = 0x467ull << 52;   // exponent field of 2^104 (hi halves)

struct Consts {
    double P[5];       // modulus, 52-bit limbs
    double pinv;       // -p^-1 mod 2^52
    uint32_t mod[8];
};

// limb k of (x << shift), x given as 8 x u32
__device__ __forceinline__ void to_limbs52(const uint32_t* a, int shift, double* out) {
    uint64_t w[5];
    uint64_t v0 = (uint64_t)a[0] | ((uint64_t)a[1] << 32), v1 = (uint64_t)a[2] | ((uint64_t)a[3] << 32);
    uint64_t v2 = (uint64_t)a[4] | ((uint64_t)a[5] << 32), v3 = (uint64_t)a[6] | ((uint64_t)a[7] << 32);
    if (shift) {
        w[0] = v0 << shift; w[1] = (v1 << shift) | (v0 >> (64 - shift)); w[2] = (v2 << shift) | (v1 >> (64 - shift));
        w[3] = (v3 << shift) | (v2 >> (64 - shift)); w[4] = v3 >> (64 - shift);
    } else { w[0] = v0; w[1] = v1; w[2] = v2; w[3] = v3; w[4] = 0; }
    uint64_t l[5];
    l[0] = w[0] & M52;
    l[1] = ((w[0] >> 52) | (w[1] << 12)) & M52;
    l[2] = ((w[1] >> 40) | (w[2] << 24)) & M52;
    l[3] = ((w[2] >> 28) | (w[3] << 36)) & M52;
    l[4] = ((w[3] >> 16) | (w[4] << 48)) & M52;
#pragma unroll
    for (int i = 0; i < 5; ++i) out[i] = __longlong_as_double((long long)(l[i] | BL)) - 4503599627370496.0;   // exact: 2^52 + l - 2^52
}

// number of lo / hi halves that land in column k from one 5x5 schoolbook product
__host__ __device__ constexpr int n_lo(int k) { return k <= 4 ? k + 1 : (k <= 8 ? 9 - k : 0); }
__host__ __device__ constexpr int n_hi(int k) { return k >= 1 ? n_lo(k - 1) : 0; }

template <class C>
__device__ __forceinline__ void mont_mul(uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
    const double C1 = 20282409603651670423947251286016.0;                      // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0; // 2^104 + 2^52 (exact)
    double A[5], B[5];
    to_limbs52(a, 0, A);
    to_limbs52(b, 4, B);
    uint64_t col[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) col[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double hi = __fma_rz(A[i], B[j], C1);
            double lo = __fma_rz(A[i], B[j], C2 - hi);
            col[i + j] += (uint64_t)__double_as_longlong(lo);
            col[i + j + 1] += (uint64_t)__double_as_longlong(hi);
        }
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        // halves already in column i: the product's, plus i lo and i hi halves from the earlier digits' q * P rows
        const uint64_t bias = (uint64_t)(n_lo(i) + i) * BL + (uint64_t)(n_hi(i) + i) * BH;
        uint64_t v = col[i] - bias + carry;
        double t = __longlong_as_double((long long)((v & M52) | BL)) - 4503599627370496.0;
        double qh = __fma_rz(t, C::pinv(), C1);
        double ql = __fma_rz(t, C::pinv(), C2 - qh);
        double q = ql - 4503599627370496.0;                    // (t * pinv) mod 2^52
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double hi = __fma_rz(q, C::P(j), C1);
            double lo = __fma_rz(q, C::P(j), C2 - hi);
            if (j == 0) v += (uint64_t)__double_as_longlong(lo) - BL;   // makes column i vanish mod 2^52
            else col[i + j] += (uint64_t)__double_as_longlong(lo);
            col[i + j + 1] += (uint64_t)__double_as_longlong(hi);
        }
        carry = v >> 52;
    }
    uint64_t l[5];
#pragma unroll
    for (int k = 5; k < 10; ++k) {
        // product halves + the q * P rows: lo halves from digits i' with 1 <= k - i' <= 4, hi halves with 0 <= k - 1 - i' <= 4
        const uint64_t bias = (uint64_t)(n_lo(k) + (9 - k)) * BL + (uint64_t)(n_hi(k) + (10 - k)) * BH;
        uint64_t v = col[k] - bias + carry;
        l[k - 5] = v & M52;
        carry = v >> 52;
    }
    uint64_t w0 = l[0] | (l[1] << 52), w1 = (l[1] >> 12) | (l[2] << 40), w2 = (l[2] >> 24) | (l[3] << 28), w3 = (l[3] >> 36) | (l[4] << 16);
    uint32_t x[8] = {(uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32), (uint32_t)w2, (uint32_t)(w2 >> 32), (uint32_t)w3, (uint32_t)(w3 >> 32)};
    uint32_t s[8];
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)x[i] - C::mod(i) - borrow;
        s[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = borrow ? x[i] : s[i];
}

// ---- field constants ------------------------------------------------------------------------------------------
// Fr: p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
struct FrC {
    __device__ __forceinline__ static double P(int j) {
        return j == 0 ? 0x1f593f0000001p0 : j == 1 ? 0x4879b9709143ep0 : j == 2 ? 0x181585d2833e8p0 : j == 3 ? 0xa029b85045b68p0 : 0x30644e72e131p0;
    }
    __device__ __forceinline__ static double pinv() { return 0x1f593efffffffp0; }
    __device__ __forceinline__ static uint32_t mod(int i) {
        return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
};
// Fq: p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
struct FqC {
    __device__ __forceinline__ static double P(int j) {
        return j == 0 ? 0x8c16d87cfd47p0 : j == 1 ? 0x916871ca8d3c2p0 : j == 2 ? 0x181585d97816ap0 : j == 3 ? 0xa029b85045b68p0 : 0x30644e72e131p0;
    }
    __device__ __forceinline__ static double pinv() { return 0x20782e4866389p0; }
    __device__ __forceinline__ static uint32_t mod(int i) {
        return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
};

}  // namespace dfma

__device__ __forceinline__ fr_t mul_dfma(const fr_t& a, const fr_t& b) { fr_t r; dfma::mont_mul<dfma::FrC>(r.l, a.l, b.l); return r; }
__device__ __forceinline__ fq_t mul_dfma(const fq_t& a, const fq_t& b) { fq_t r; dfma::mont_mul<dfma::FqC>(r.l, a.l, b.l); return r; }

}  // namespace zk
