// Scalar -> signed window digits for the bucket method (shared by msm.cu and the host-side tests: plain __host__ __device__ code).
#pragma once
#include "fp.cuh"

namespace zk {

__host__ __device__ __forceinline__ uint32_t get_bits(const uint32_t* l, unsigned pos, unsigned c) {
    unsigned word = pos >> 5, shift = pos & 31;
    if (word >= 8) return 0;
    uint32_t v = l[word] >> shift;
    if (shift + c > 32 && word + 1 < 8) v |= l[word + 1] << (32 - shift);
    return v & ((1u << c) - 1);
}

// Calls f(w, magnitude in [1, 2^(c-1)], negative) for each non-zero signed digit of canonical scalar s.
template <class F>
__host__ __device__ __forceinline__ void for_each_digit(const uint32_t* s, unsigned c, unsigned W, F f) {
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (unsigned w = 0; w < W; ++w) {
        uint32_t d = get_bits(s, w * c, c) + carry;
        if (d > half) { carry = 1; uint32_t mag = (1u << c) - d; if (mag) f(w, mag, true); }
        else { carry = 0; if (d) f(w, d, false); }
    }
}

// The integer a digit kernel decomposes for point i: canonical s_i, or in difference mode (the prefix-sum half of a fixed-base
// table, MsmPlan::diff_offset) canonical s_i - s_{i+1} with s_n = 0.  A value within 2^224 below r is replaced by r minus itself
// (flip: every digit changes sign), so that -1, -x for short x and small downward steps decompose as briefly as their negatives.
// Only then: flipping every value above r/2 would halve the range of the top window's digits and pile twice as many entries on
// the lowest buckets (measured: +20 % bucket-accumulation time on uniformly random scalars).
__host__ __device__ __forceinline__ fr_t digit_scalar(const fr_t& s_mont, const fr_t& next_mont, bool diff, bool& flip) {
    fr_t d = from_mont(diff ? s_mont - next_mont : s_mont);
    const uint32_t M[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t t[8], borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t v = (uint64_t)M[i] - d.l[i] - borrow;
        t[i] = (uint32_t)v;
        borrow = (uint32_t)(v >> 63);
    }
    flip = t[7] == 0;   // r - d < 2^224 (d = 0 gives t = r: no flip)
    if (flip) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d.l[i] = t[i];
    }
    return d;
}

}  // namespace zk
