// Reader for the proving-key artefact the reference ships: `pk.bin` = `k: u32 LE` ‖ `ProvingKey::to_bytes(RawBytesUnchecked)`
// (`marshall_pk`, written by /root/reference/crates/shielder_bindings/build.rs:19-33, read by
// src/circuits/mod.rs:35-48,89-101 and cached by /root/reference/crates/shielder-cli/src/shielder_ops/pk.rs:68-126).
//
// Layout [UPSTREAM-MEMORY of halo2_proofs v0.3.0 plonk.rs / poly.rs / permutation.rs / helpers.rs; not verifiable here — the
// header variant is isolated in `parse_vk` so that a different halo2 revision needs one function changed]:
//   ProvingKey::write          vk | l0 | l_last | l_active_row | fixed_values | fixed_polys | fixed_cosets | permutation pk
//   VerifyingKey::write        k: u32 BE | #fixed_commitments: u32 BE | commitments | permutation commitments (count known from
//                              the constraint system) | selectors: num_selectors x n bits, packed 8 per byte
//   Polynomial::write          len: u32 BE | len field elements
//   write_polynomial_slice     count: u32 BE | polynomials
//   permutation::ProvingKey    permutations (sigma, Lagrange values) | polys (coefficients) | cosets (extended domain)
//   RawBytesUnchecked          Fr: 32 B little-endian Montgomery limbs; G1Affine: x ‖ y, 64 B
// The file does NOT hold the constraint system (gates, queries, lookup expressions, permutation columns): upstream rebuilds it from
// the circuit type (`ProvingKey::read::<_, C>`); here it comes from the constraint-system blob the Rust exporter of INTEGRATION.md
// writes, which also carries `vk.transcript_repr()` (a hash of Rust Debug output that cannot be recomputed outside Rust).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "ec.cuh"

namespace zk {

struct PkFile {
    uint32_t k = 0;
    std::vector<g1_affine_t> fixed_commitments, perm_commitments;
    // views into the caller's buffer (element counts are validated against the constraint system by the loader)
    struct Poly { const uint8_t* p = nullptr; size_t len = 0; };
    Poly l0, l_last, l_active_row;                              // extended domain
    std::vector<Poly> fixed_values, fixed_polys, fixed_cosets;  // n, n, 2^ek
    std::vector<Poly> perm_values, perm_polys, perm_cosets;

    struct Cursor {
        const uint8_t* p; const uint8_t* end;
        void need(size_t n) const { ZK_REQUIRE((size_t)(end - p) >= n, "pk.bin: truncated"); }
        uint32_t u32_be() { need(4); uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return v; }
        uint32_t u32_le() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
        g1_affine_t point() { need(64); g1_affine_t a; memcpy(&a, p, 64); p += 64; return a; }
        Poly poly(size_t max_len) {
            Poly r; r.len = u32_be();
            ZK_REQUIRE(r.len <= max_len, "pk.bin: polynomial longer than the extended domain");
            need(r.len * 32); r.p = p; p += r.len * 32;
            return r;
        }
        std::vector<Poly> slice(size_t max_count, size_t max_len) {
            uint32_t cnt = u32_be();
            ZK_REQUIRE(cnt <= max_count, "pk.bin: more polynomials than the constraint system has columns");
            std::vector<Poly> v(cnt);
            for (auto& q : v) q = poly(max_len);
            return v;
        }
    };

    // VerifyingKey::write — the revision-dependent part
    static void parse_vk(Cursor& c, PkFile& f, size_t num_perm_columns, size_t num_selectors) {
        uint32_t vk_k = c.u32_be();
        ZK_REQUIRE(vk_k == f.k, "pk.bin: the verifying key's k differs from the file's k prefix");
        uint32_t nf = c.u32_be();
        ZK_REQUIRE(nf <= 4096, "pk.bin: implausible number of fixed commitments");
        f.fixed_commitments.resize(nf);
        for (auto& p : f.fixed_commitments) p = c.point();
        f.perm_commitments.resize(num_perm_columns);
        for (auto& p : f.perm_commitments) p = c.point();
        const size_t n = (size_t)1 << f.k, per = (n + 7) / 8;
        c.need(num_selectors * per);
        c.p += num_selectors * per;   // selector bit-vectors: already folded into the fixed columns by `compress_selectors`
    }

    static PkFile parse(const uint8_t* data, size_t len, size_t num_fixed, size_t num_perm_columns, size_t num_selectors, unsigned ext_k) {
        PkFile f;
        Cursor c{data, data + len};
        f.k = c.u32_le();
        ZK_REQUIRE(f.k >= 1 && f.k <= 24 && ext_k >= f.k && ext_k <= 28, "pk.bin: k out of range");
        const size_t n = (size_t)1 << f.k, en = (size_t)1 << ext_k;
        parse_vk(c, f, num_perm_columns, num_selectors);
        f.l0 = c.poly(en); f.l_last = c.poly(en); f.l_active_row = c.poly(en);
        f.fixed_values = c.slice(num_fixed, n); f.fixed_polys = c.slice(num_fixed, n); f.fixed_cosets = c.slice(num_fixed, en);
        f.perm_values = c.slice(num_perm_columns, n); f.perm_polys = c.slice(num_perm_columns, n); f.perm_cosets = c.slice(num_perm_columns, en);
        ZK_REQUIRE(c.p == c.end, "pk.bin: trailing bytes (a different halo2 revision? see pk_file.hpp)");
        auto all = [&](const std::vector<Poly>& v, size_t count, size_t plen, const char* what) {
            ZK_REQUIRE(v.size() == count, std::string("pk.bin: wrong number of ") + what);
            for (auto& q : v) ZK_REQUIRE(q.len == plen, std::string("pk.bin: wrong length of ") + what);
        };
        ZK_REQUIRE(f.fixed_commitments.size() == num_fixed, "pk.bin: number of fixed commitments differs from the constraint system");
        ZK_REQUIRE(f.l0.len == en && f.l_last.len == en && f.l_active_row.len == en, "pk.bin: l0 / l_last / l_active_row are not extended-domain polynomials");
        all(f.fixed_values, num_fixed, n, "fixed_values"); all(f.fixed_polys, num_fixed, n, "fixed_polys"); all(f.fixed_cosets, num_fixed, en, "fixed_cosets");
        all(f.perm_values, num_perm_columns, n, "permutations"); all(f.perm_polys, num_perm_columns, n, "permutation polys");
        all(f.perm_cosets, num_perm_columns, en, "permutation cosets");
        return f;
    }
};

}  // namespace zk
