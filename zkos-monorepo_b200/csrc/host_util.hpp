// Host-side pieces of the prover that sit on the latency-critical path between GPU phases:
// Keccak-256 + the EVM transcript (reference: zkOS-circuits `transcript` crate; in-repo spec
// /root/reference/crates/halo2-verifier/templates/Halo2Verifier.sol:101-124,247-307), the seeded RNG
// (`SmallRng::seed_from_u64`, /root/reference/crates/shielder-setup/lib.rs:29-40 = xoshiro256++ seeded by
// SplitMix64, rand 0.8.5), and byte codecs (big-endian canonical words of the proof,
// /root/reference/crates/halo2-verifier/src/lib/verifier_contract.rs:14-20).
// Product code: independent of oracle/.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "fp.cuh"
#include "ec.cuh"

namespace zk {

// ---- Keccak-256 (Ethereum padding) --------------------------------------------------------------
static inline uint64_t rotl64_(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static inline void keccak_permute(uint64_t s[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int r = 0; r < 24; ++r) {
        uint64_t bc[5];
        for (int i = 0; i < 5; ++i) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; ++i) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl64_(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        uint64_t t = s[1];
        for (int i = 0; i < 24; ++i) { int j = PI[i]; uint64_t b = s[j]; s[j] = rotl64_(t, RHO[i]); t = b; }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; ++i) bc[i] = s[j + i];
            for (int i = 0; i < 5; ++i) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= RC[r];
    }
}
static inline void keccak256(const uint8_t* in, size_t len, uint8_t out[32]) {
    uint64_t s[25] = {0};
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; ++i) { uint64_t w; memcpy(&w, in + 8 * i, 8); s[i] ^= w; }
        keccak_permute(s); in += rate; len -= rate;
    }
    uint8_t last[136] = {0};
    memcpy(last, in, len);
    last[len] ^= 0x01; last[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; ++i) { uint64_t w; memcpy(&w, last + 8 * i, 8); s[i] ^= w; }
    keccak_permute(s);
    memcpy(out, s, 32);
}

// ---- field <-> bytes ---------------------------------------------------------------------------
template <class Tag>
static inline void fe_to_be_bytes(const Fe<Tag>& mont, uint8_t out[32]) {
    Fe<Tag> c = from_mont(mont);
    for (int i = 0; i < 8; ++i) {
        uint32_t w = c.l[7 - i];
        out[4 * i] = (uint8_t)(w >> 24); out[4 * i + 1] = (uint8_t)(w >> 16); out[4 * i + 2] = (uint8_t)(w >> 8); out[4 * i + 3] = (uint8_t)w;
    }
}
// 32-byte big-endian integer (any value < 2^256) -> Fr Montgomery, reduced mod r
static inline fr_t fr_from_be_bytes_reduce(const uint8_t in[32]) {
    fr_t raw;
    for (int i = 0; i < 8; ++i)
        raw.l[7 - i] = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
    fr_reduce_raw(raw);  // the limb code needs operands < r
    return raw * fe_r2<FrTag>();
}

// ---- transcript --------------------------------------------------------------------------------
struct Transcript {
    std::vector<uint8_t> buf;  // previous hash ‖ absorbed bytes
    bool fresh = false;        // buf is exactly the previous hash
    uint8_t* out;              // proof write cursor
    explicit Transcript(uint8_t* proof_out) : out(proof_out) { buf.reserve(4096); }
    void absorb(const uint8_t* b, size_t n) { buf.insert(buf.end(), b, b + n); fresh = false; }
    void common_scalar(const fr_t& s) { uint8_t w[32]; fe_to_be_bytes(s, w); absorb(w, 32); }
    void write_scalar(const fr_t& s) { uint8_t w[32]; fe_to_be_bytes(s, w); memcpy(out, w, 32); out += 32; absorb(w, 32); }
    void write_point(const g1_affine_t& p) {
        uint8_t w[64]; fe_to_be_bytes(p.x, w); fe_to_be_bytes(p.y, w + 32);
        memcpy(out, w, 64); out += 64; absorb(w, 64);
    }
    fr_t squeeze() {
        if (fresh) buf.push_back(0x01);
        uint8_t h[32]; keccak256(buf.data(), buf.size(), h);
        buf.assign(h, h + 32); fresh = true;
        return fr_from_be_bytes_reduce(h);
    }
};

// ---- proof RNGs ---------------------------------------------------------------------------------
// `generate_proof(.., rng: &mut impl RngCore)` (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111) is called
// with a running `SmallRng` in the seeded tests (/root/reference/crates/halo2-verifier/src/generator.rs:117-130) and with
// `OsRng` / `thread_rng()` in production (/root/reference/crates/shielder-account/src/call_data.rs:499,
// /root/reference/crates/shielder_bindings/src/circuits/deposit.rs:108).  ProofRng is the `RngCore` the prover draws from:
//   xoshiro256++ (rand 0.8.5 SmallRng on 64-bit targets) from a u64 seed or from a caller-owned running state, or
//   ChaCha20 (rand_chacha 0.3.1 ChaCha20Rng::from_seed) from 32 bytes of caller entropy.
struct SmallRng {
    uint64_t s[4];
    explicit SmallRng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9e3779b97f4a7c15ULL;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    explicit SmallRng(const uint64_t state[4]) { memcpy(s, state, 32); }
    uint64_t next_u64() {
        uint64_t r = rotl64_(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl64_(s[3], 45);
        return r;
    }
};

// rand_chacha 0.3.1 ChaCha20Rng: key = seed, 64-bit block counter in words 12-13, stream id 0; output words in block order.
struct ChaCha20Host {
    uint32_t key[8]; uint64_t counter = 0; uint32_t buf[16]; unsigned idx = 16;
    explicit ChaCha20Host(const uint8_t seed[32]) { memcpy(key, seed, 32); }
    static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
    void refill() {
        uint32_t st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        for (int i = 0; i < 8; ++i) st[4 + i] = key[i];
        st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
        uint32_t x[16]; memcpy(x, st, 64);
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) buf[i] = x[i] + st[i];
        ++counter; idx = 0;
    }
    uint32_t next_u32() { if (idx >= 16) refill(); return buf[idx++]; }
    uint64_t next_u64() { uint64_t lo = next_u32(), hi = next_u32(); return lo | (hi << 32); }
};

enum { RNG_SEED_U64 = 0, RNG_XOSHIRO_STATE = 1, RNG_CHACHA20_SEED = 2 };
static inline size_t rng_data_stride(int mode) { return mode == RNG_SEED_U64 ? 8 : 32; }

struct ProofRng {
    int mode;
    SmallRng xo;
    ChaCha20Host cc;
    static const uint8_t* zero32() { static const uint8_t z[32] = {0}; return z; }
    ProofRng(int mode_, const uint8_t* data)
        : mode(mode_), xo((uint64_t)0), cc(mode_ == RNG_CHACHA20_SEED ? data : zero32()) {
        if (mode == RNG_SEED_U64) { uint64_t seed; memcpy(&seed, data, 8); xo = SmallRng(seed); }
        else if (mode == RNG_XOSHIRO_STATE) { uint64_t st[4]; memcpy(st, data, 32); xo = SmallRng(st); }
    }
    uint64_t next_u64() { return mode == RNG_CHACHA20_SEED ? cc.next_u64() : xo.next_u64(); }
    // `Fr::random` consumes eight u64 (from_u512); the reduction is done on the device
    void next_wide(uint64_t out[8]) { for (int i = 0; i < 8; ++i) out[i] = next_u64(); }
    void skip_wide() { for (int i = 0; i < 8; ++i) next_u64(); }
    // RngCore::fill_bytes of 32 bytes: four next_u64 (xoshiro) / eight output words (ChaCha block rng) — the same bytes
    void fill_bytes32(uint8_t out[32]) { for (int i = 0; i < 4; ++i) { uint64_t v = next_u64(); memcpy(out + 8 * i, &v, 8); } }
    // running state handed back to the caller (xoshiro state mode): the host's rng continues where the proof stopped
    void store_state(uint8_t* data) const { if (mode == RNG_XOSHIRO_STATE) memcpy(data, xo.s, 32); }
};

}  // namespace zk
