// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
//
// Keccak-256, the seeded RNGs and the SRS file readers that sit either side of the hot path.
//   * Keccak-256: pinned by the known answers in
//     /root/reference/crates/shielder-account/src/secrets.rs:75-106 (tests/test_oracle_kat.py).
//   * SmallRng::seed_from_u64 (xoshiro256++ / SplitMix64, rand 0.8.5) — the "seeded RNG" of
//     /root/reference/crates/shielder-setup/lib.rs:29-40 [UPSTREAM-MEMORY, SURVEY Appendix A].
//   * ChaCha20Rng (rand_chacha 0.3.1, /root/reference/Cargo.lock:4061-4062) used by halo2's
//     vanishing-argument random polynomial [UPSTREAM-MEMORY].
//   * .ptau / RawBytes SRS readers: /root/reference/crates/powers-of-tau/lib.rs:61-231.
#pragma once
#include "bn254.hpp"
#include <stdexcept>
#include <fstream>

namespace oracle {

// ------------------------------------------------------------------------------------------
// Keccak-256 (original Keccak padding 0x01, as the EVM's KECCAK256)
// ------------------------------------------------------------------------------------------
static inline u64 rotl64(u64 x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static inline void keccak_f1600(u64 st[25]) {
    static const u64 RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const unsigned ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; ++round) {
        u64 C[5], D[5], B[25];
        for (int x = 0; x < 5; ++x) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; ++x) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) st[i] ^= D[i % 5];
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[round];
    }
}
static inline void keccak256(const uint8_t* in, size_t len, uint8_t out[32]) {
    u64 st[25] = {0};
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; ++i) { u64 w; memcpy(&w, in + 8 * i, 8); st[i] ^= w; }
        keccak_f1600(st); in += rate; len -= rate;
    }
    uint8_t blk[136] = {0};
    memcpy(blk, in, len);
    blk[len] ^= 0x01; blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; ++i) { u64 w; memcpy(&w, blk + 8 * i, 8); st[i] ^= w; }
    keccak_f1600(st);
    memcpy(out, st, 32);
}

// ------------------------------------------------------------------------------------------
// RNGs
// ------------------------------------------------------------------------------------------
struct RngCore {
    virtual u64 next_u64() = 0;
    virtual void fill_bytes(uint8_t* out, size_t n) = 0;
    virtual ~RngCore() {}
};

// rand 0.8.5 SmallRng on 64-bit targets = xoshiro256++, seed_from_u64 via SplitMix64.
struct SmallRng : RngCore {
    u64 s[4];
    explicit SmallRng(u64 seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9e3779b97f4a7c15ULL;
            u64 z = seed;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    // a running generator: the four state words as rand's Xoshiro256PlusPlus holds them
    explicit SmallRng(const u64 state[4]) { for (int i = 0; i < 4; ++i) s[i] = state[i]; }
    u64 next_u64() override {
        u64 result = rotl64(s[0] + s[3], 23) + s[0];
        u64 t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl64(s[3], 45);
        return result;
    }
    void fill_bytes(uint8_t* out, size_t n) override {  // successive next_u64, little-endian
        while (n) {
            u64 v = next_u64(); size_t m = n < 8 ? n : 8;
            memcpy(out, &v, m); out += m; n -= m;
        }
    }
};

// rand_chacha 0.3.1 ChaCha20Rng: 32-byte key seed, 64-bit block counter in words 12-13, stream 0.
struct ChaCha20Rng : RngCore {
    uint32_t key[8]; u64 counter = 0; uint32_t buf[16]; unsigned idx = 16;
    explicit ChaCha20Rng(const uint8_t seed[32]) { memcpy(key, seed, 32); }
    static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
    void refill() {
        uint32_t st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        for (int i = 0; i < 8; ++i) st[4 + i] = key[i];
        st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
        uint32_t x[16]; memcpy(x, st, 64);
#define ZK_QR(a, b, c, d) x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
                          x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        for (int r = 0; r < 10; ++r) {
            ZK_QR(0, 4, 8, 12) ZK_QR(1, 5, 9, 13) ZK_QR(2, 6, 10, 14) ZK_QR(3, 7, 11, 15)
            ZK_QR(0, 5, 10, 15) ZK_QR(1, 6, 11, 12) ZK_QR(2, 7, 8, 13) ZK_QR(3, 4, 9, 14)
        }
#undef ZK_QR
        for (int i = 0; i < 16; ++i) buf[i] = x[i] + st[i];
        counter++; idx = 0;
    }
    uint32_t next_u32() { if (idx >= 16) refill(); return buf[idx++]; }
    u64 next_u64() override { u64 lo = next_u32(); u64 hi = next_u32(); return lo | (hi << 32); }
    void fill_bytes(uint8_t* out, size_t n) override {  // whole words consumed, remainder discarded
        while (n) {
            uint32_t v = next_u32(); size_t m = n < 4 ? n : 4;
            memcpy(out, &v, m); out += m; n -= m;
        }
    }
};

// halo2curves `Field::random` = from_u512 of eight next_u64
template <class F>
static inline F random_field(RngCore& rng) {
    u64 w[8];
    for (int i = 0; i < 8; ++i) w[i] = rng.next_u64();
    return F::from_u512(w);
}

// ------------------------------------------------------------------------------------------
// SRS (ParamsKZG) — the two on-disk formats of crates/powers-of-tau/lib.rs
// ------------------------------------------------------------------------------------------
struct G2AffineRaw { Fq x0, x1, y0, y1; };  // x = x0 + x1*u, y = y0 + y1*u
struct Srs {
    unsigned k = 0;
    std::vector<G1Affine> g, g_lagrange;
    G2AffineRaw g2, s_g2;
};

static inline std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static inline Fq fq_from_mont_bytes(const uint8_t* p) { u64 l[4]; memcpy(l, p, 32); return Fq::from_raw_mont(l); }
static inline G2AffineRaw g2_from_mont_bytes(const uint8_t* p) {
    return {fq_from_mont_bytes(p), fq_from_mont_bytes(p + 32), fq_from_mont_bytes(p + 64), fq_from_mont_bytes(p + 96)};
}

// Format::Raw — `ParamsKZG::read_custom(RawBytes)` (lib.rs:64): k:u32 LE, g[n], g_lagrange[n], g2, s_g2.
static inline Srs srs_read_raw(const std::vector<uint8_t>& buf) {
    Srs s;
    if (buf.size() < 4) throw std::runtime_error("raw srs: short file");
    uint32_t k; memcpy(&k, buf.data(), 4); s.k = k;
    size_t n = (size_t)1 << k;
    if (buf.size() != 4 + 2 * n * 64 + 256) throw std::runtime_error("raw srs: size mismatch");
    s.g.resize(n); s.g_lagrange.resize(n);
    const uint8_t* p = buf.data() + 4;
    for (size_t i = 0; i < n; ++i, p += 64) s.g[i] = {fq_from_mont_bytes(p), fq_from_mont_bytes(p + 32)};
    for (size_t i = 0; i < n; ++i, p += 64) s.g_lagrange[i] = {fq_from_mont_bytes(p), fq_from_mont_bytes(p + 32)};
    s.g2 = g2_from_mont_bytes(p); s.s_g2 = g2_from_mont_bytes(p + 128);
    return s;
}

// Format::PerpetualPowersOfTau (lib.rs:66-72, offsets :76-123): header size u64 at byte 16; k u32 at
// 24+hs-8; G1 powers start at 24+hs+12; G2 powers at g1_off + 64*(2n-1) + 12.  Coordinates are
// Montgomery-form LE (the reader multiplies the plain `from_repr` value by R^-1, lib.rs:206-224).
// g_lagrange is NOT in the file: from_parts(.., None, ..) derives it with g_to_lagrange (caller).
static inline Srs srs_read_ptau(const std::vector<uint8_t>& buf) {
    Srs s;
    u64 hs; memcpy(&hs, buf.data() + 16, 8);
    uint32_t k; memcpy(&k, buf.data() + 24 + hs - 8, 4); s.k = k;
    size_t n = (size_t)1 << k;
    size_t g1_off = 24 + hs + 12;
    size_t g2_off = g1_off + 2 * 32 * (2 * n - 1) + 12;
    if (buf.size() < g2_off + 256) throw std::runtime_error("ptau: short file");
    s.g.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const uint8_t* p = buf.data() + g1_off + 64 * i;
        s.g[i] = {fq_from_mont_bytes(p), fq_from_mont_bytes(p + 32)};
        if (!s.g[i].on_curve()) throw std::runtime_error("ptau: point not on curve");  // from_xy(..).unwrap()
    }
    s.g2 = g2_from_mont_bytes(buf.data() + g2_off);
    s.s_g2 = g2_from_mont_bytes(buf.data() + g2_off + 128);
    return s;
}

}  // namespace oracle
