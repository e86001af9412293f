// Host build (g++, no nvcc, no GPU) of the product's host-side prover logic — csrc/host_util.hpp (Keccak-256,
// EVM transcript, SmallRng, byte codecs), csrc/plonk_types.hpp (constraint-system parsing, derived sizes,
// permutation assembly) and the host paths of fp.cuh / ec.cuh.  Test-only shim exporting a small C ABI.
#include <cstring>
#include "host_util.hpp"
#include "plonk_types.hpp"
using namespace zk;
extern "C" {
void hl_keccak256(const uint8_t* in, size_t len, uint8_t* out) { keccak256(in, len, out); }
void hl_smallrng(uint64_t seed, uint64_t* out, size_t n) { SmallRng r(seed); for (size_t i = 0; i < n; ++i) out[i] = r.next_u64(); }
// transcript: absorb `n_scalars` Montgomery scalars after the digest, then squeeze `n_ch` challenges (Montgomery out)
void hl_transcript(const uint32_t* digest, const uint32_t* scalars, size_t n_scalars, const uint32_t* points, size_t n_points,
                   uint32_t* challenges, size_t n_ch, uint8_t* proof_out) {
    Transcript tr(proof_out);
    fr_t d; memcpy(d.l, digest, 32); tr.common_scalar(d);
    for (size_t i = 0; i < n_scalars; ++i) { fr_t s; memcpy(s.l, scalars + 8 * i, 32); tr.write_scalar(s); }
    for (size_t i = 0; i < n_points; ++i) { g1_affine_t p; memcpy(&p, points + 16 * i, 64); tr.write_point(p); }
    for (size_t i = 0; i < n_ch; ++i) { fr_t c = tr.squeeze(); memcpy(challenges + 8 * i, c.l, 32); }
}
void hl_fr_from_be_reduce(const uint8_t* in, uint32_t* out) { fr_t v = fr_from_be_bytes_reduce(in); memcpy(out, v.l, 32); }
// info = k, n, degree, blinding_factors, chunk_len, num_perm_sets, num_quotients, extended_k, num_evals, proof_len; returns 0 / -1
int hl_cs_info(const uint8_t* blob, size_t len, uint64_t* info, char* err, size_t errlen) {
    try {
        CsDesc c = CsDesc::parse(blob, len);
        uint64_t v[10] = {c.k, c.n(), c.degree(), c.blinding_factors(), c.chunk_len(), c.num_perm_sets(), c.num_quotients(), c.extended_k(), c.num_evals(), c.proof_len()};
        memcpy(info, v, sizeof v);
        return 0;
    } catch (const std::exception& e) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; return -1; }
}
// sigma mapping after applying the blob's copy constraints: out_col/out_row [S*n]
int hl_perm_mapping(const uint8_t* blob, size_t len, uint32_t* out_col, uint32_t* out_row) {
    try {
        CsDesc c = CsDesc::parse(blob, len);
        PermAssembly as(c.perm_columns.size(), c.n());
        for (auto& cp : c.copies) as.copy(cp.lcol, cp.lrow, cp.rcol, cp.rrow);
        memcpy(out_col, as.map_col.data(), as.map_col.size() * 4); memcpy(out_row, as.map_row.data(), as.map_row.size() * 4);
        return 0;
    } catch (...) { return -1; }
}
}
extern "C" {
// field inversion: which 0 = fe_inv (uniform-flow binary GCD), 1 = Fermat (fe_inv_pow), 2 = binary extended Euclid; field 0 = Fr, 1 = Fq
void hl_inv(int field, int which, const uint32_t* a, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) { fr_t x; memcpy(x.l, a + 8 * i, 32); fr_t r = which == 1 ? fe_inv_pow(x) : which == 2 ? fe_inv_euclid(x) : fe_inv(x); memcpy(out + 8 * i, r.l, 32); }
        else { fq_t x; memcpy(x.l, a + 8 * i, 32); fq_t r = which == 1 ? fe_inv_pow(x) : which == 2 ? fe_inv_euclid(x) : fe_inv(x); memcpy(out + 8 * i, r.l, 32); }
    }
}
}
#include "pk_file.hpp"
extern "C" {
// pk.bin reader (csrc/pk_file.hpp) against a constraint-system-only blob: info = k, #fixed commitments, #perm commitments,
// fixed_values count, fixed_cosets[0] length, perm_cosets count, l0 length; `first` receives the first element of
// fixed_values[0], fixed_polys[0], fixed_cosets[0], perm_values[0], perm_cosets[last], l_active_row (6 x 8 u32)
int hl_pk_file(const uint8_t* cs_blob, size_t cs_len, const uint8_t* pk_bin, size_t pk_len, uint64_t* info, uint32_t* first, char* err, size_t errlen) {
    try {
        CsDesc c = CsDesc::parse(cs_blob, cs_len);
        if (!c.cs_only) throw std::runtime_error("not a constraint-system-only blob");
        PkFile f = PkFile::parse(pk_bin, pk_len, c.num_fixed, c.perm_columns.size(), c.num_selectors, c.extended_k());
        uint64_t v[7] = {f.k, f.fixed_commitments.size(), f.perm_commitments.size(), f.fixed_values.size(), f.fixed_cosets.empty() ? 0 : f.fixed_cosets[0].len,
                         f.perm_cosets.size(), f.l0.len};
        memcpy(info, v, sizeof v);
        const uint8_t* src[6] = {f.fixed_values[0].p, f.fixed_polys[0].p, f.fixed_cosets[0].p, f.perm_values[0].p, f.perm_cosets.back().p, f.l_active_row.p};
        for (int i = 0; i < 6; ++i) memcpy(first + 8 * i, src[i], 32);
        memcpy(first + 48, c.transcript_repr.l, 32);
        return 0;
    } catch (const std::exception& e) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; return -1; }
}
// the prover's rng (csrc/host_util.hpp ProofRng): n u64 of mode 0 / 1 / 2; for mode 1 the advanced state is written back
void hl_proof_rng(int mode, uint8_t* data, uint64_t* out, size_t n) {
    ProofRng r(mode, data);
    for (size_t i = 0; i < n; ++i) out[i] = r.next_u64();
    r.store_state(data);
}
}
#include "msm_digits.cuh"
extern "C" {
// csrc/msm_digits.cuh: the integer the digit sort decomposes for point i and its signed c-bit digits.
// s, next: n Montgomery scalars each; out_mag [n * W] signed digits (int32: +-magnitude), out_flip [n]
void hl_msm_digits(const uint32_t* s, const uint32_t* next, size_t n, int diff, unsigned c, unsigned W, int32_t* out_digits, uint8_t* out_flip) {
    for (size_t i = 0; i < n; ++i) {
        fr_t a, b; memcpy(a.l, s + 8 * i, 32); memcpy(b.l, next + 8 * i, 32);
        bool flip = false;
        fr_t d = digit_scalar(a, b, diff != 0, flip);
        out_flip[i] = flip;
        for (unsigned w = 0; w < W; ++w) out_digits[i * W + w] = 0;
        for_each_digit(d.l, c, W, [&](unsigned w, uint32_t mag, bool negative) { out_digits[i * W + w] = (negative != flip) ? -(int32_t)mag : (int32_t)mag; });
    }
}
}
