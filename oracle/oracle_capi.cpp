// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
// C ABI over the oracle so tests/ and bench.py (cpu_baseline / --impl reference) can drive it
// through ctypes.  All field buffers are 4xu64 Montgomery little-endian limbs ("Mont-LE"), all
// points affine (x,y) Mont-LE with identity (0,0) — the same layouts libzkgpu's C ABI uses
// (include/zkgpu.h), so the same numpy arrays feed both sides of a parity test.
#include "bn254.hpp"
#include "arith.hpp"
#include "misc.hpp"
#include "pairing.hpp"
#include "plonk.hpp"
#include "poseidon2.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; } return 0;

template <class F> static F ld(const u64* p) { return F::from_raw_mont(p); }
template <class F> static void st(u64* p, const F& f) { memcpy(p, f.v.l, 32); }
static G1Affine ldp(const u64* p) { return {ld<Fq>(p), ld<Fq>(p + 4)}; }
static void stp(u64* p, const G1Affine& a) { st(p, a.x); st(p + 4, a.y); }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// Poseidon2 t=8: m hashes of `len` (1..7) Mont-LE elements each
int orc_poseidon2_hash(const u64* in, size_t len, size_t m, u64* out) {
    ORC_TRY
    const Poseidon2& P = Poseidon2::get();
    std::vector<Fr> v(len);
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < len; ++j) v[j] = ld<Fr>(in + 4 * (i * len + j));
        st(out + 4 * i, P.hash(v.data(), len));
    }
    ORC_CATCH
}
// m Merkle paths of height x 7 elements -> roots, consistency flags
int orc_merkle_root(const u64* paths, size_t height, size_t m, u64* roots, uint8_t* consistent) {
    ORC_TRY
    const Poseidon2& P = Poseidon2::get();
    std::vector<Fr> v(height * 7);
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < height * 7; ++j) v[j] = ld<Fr>(paths + 4 * (i * height * 7 + j));
        bool ok = true;
        st(roots + 4 * i, P.merkle_root(v.data(), height, &ok));
        if (consistent) consistent[i] = ok ? 1 : 0;
    }
    ORC_CATCH
}

// op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 square(a), 5 neg(a); field: 0 Fr, 1 Fq; elementwise over n
int orc_field_op(int field, int op, const u64* a, const u64* b, u64* out, size_t n) {
    ORC_TRY
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) {
            Fr x = ld<Fr>(a + 4 * i), y = b ? ld<Fr>(b + 4 * i) : Fr::zero(), r;
            switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break;
                          case 3: r = x.inv(); break; case 4: r = x.square(); break; default: r = -x; }
            st(out + 4 * i, r);
        } else {
            Fq x = ld<Fq>(a + 4 * i), y = b ? ld<Fq>(b + 4 * i) : Fq::zero(), r;
            switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break;
                          case 3: r = x.inv(); break; case 4: r = x.square(); break; default: r = -x; }
            st(out + 4 * i, r);
        }
    }
    ORC_CATCH
}
// canonical LE integers (< 2^256, need not be reduced... must be < p) -> Montgomery and back
int orc_to_mont(int field, const u64* canon, u64* out, size_t n) {
    ORC_TRY
    for (size_t i = 0; i < n; ++i) {
        U256 t; memcpy(t.l, canon + 4 * i, 32);
        if (field == 0) st(out + 4 * i, Fr::from_u256(t)); else st(out + 4 * i, Fq::from_u256(t));
    }
    ORC_CATCH
}
int orc_from_mont(int field, const u64* mont, u64* out, size_t n) {
    ORC_TRY
    for (size_t i = 0; i < n; ++i) {
        U256 t = field == 0 ? ld<Fr>(mont + 4 * i).to_u256() : ld<Fq>(mont + 4 * i).to_u256();
        memcpy(out + 4 * i, t.l, 32);
    }
    ORC_CATCH
}
int orc_fr_from_u512(const u64* wide, u64* out, size_t n) {
    ORC_TRY
    for (size_t i = 0; i < n; ++i) st(out + 4 * i, Fr::from_u512(wide + 8 * i));
    ORC_CATCH
}
// constants: 0 ROOT_OF_UNITY, 1 DELTA, 2 ZETA, 3 GENERATOR(7), 4 R (one)
int orc_fr_const(int which, u64* out) {
    ORC_TRY
    Fr v = which == 0 ? FrConst::root_of_unity() : which == 1 ? FrConst::delta() : which == 2 ? FrConst::zeta()
         : which == 3 ? FrConst::generator() : Fr::one();
    st(out, v);
    ORC_CATCH
}

// G1: op 0 add(a,b) 1 double(a) 2 scalar-mul(a, k=b as Fr mont); affine in/out
int orc_g1_op(int op, const u64* a, const u64* b, u64* out) {
    ORC_TRY
    G1 p = G1::from_affine(ldp(a)), r;
    if (op == 0) r = p.add(G1::from_affine(ldp(b)));
    else if (op == 1) r = p.dbl();
    else r = p.mul(ld<Fr>(b));
    stp(out, r.to_affine());
    ORC_CATCH
}
int orc_g1_on_curve(const u64* pts, size_t n) {
    for (size_t i = 0; i < n; ++i) if (!ldp(pts + 8 * i).on_curve()) return 0;
    return 1;
}

// best_multiexp(coeffs, bases) with the reference's thread chunking; affine-normalised result
int orc_msm(const u64* scalars, const u64* bases, size_t n, unsigned threads, u64* out_affine) {
    ORC_TRY
    std::vector<Fr> s(n); std::vector<G1Affine> b(n);
    for (size_t i = 0; i < n; ++i) { s[i] = ld<Fr>(scalars + 4 * i); b[i] = ldp(bases + 8 * i); }
    stp(out_affine, best_multiexp(s.data(), b.data(), n, threads).to_affine());
    ORC_CATCH
}
// best_fft in place
int orc_fft(u64* a, const u64* omega, unsigned log_n, unsigned threads) {
    ORC_TRY
    size_t n = (size_t)1 << log_n;
    best_fft(reinterpret_cast<Fr*>(a), n, ld<Fr>(omega), log_n, threads);
    ORC_CATCH
}
// EvaluationDomain::new(j,k): out[0]=extended_k; omegas: omega, omega_inv, ext_omega, ext_omega_inv (4x4 u64)
int orc_domain(unsigned j, unsigned k, unsigned* extended_k, u64* omegas) {
    ORC_TRY
    EvaluationDomain d(j, k);
    *extended_k = d.extended_k;
    st(omegas, d.omega); st(omegas + 4, d.omega_inv); st(omegas + 8, d.extended_omega); st(omegas + 12, d.extended_omega_inv);
    ORC_CATCH
}
// which: 0 lagrange_to_coeff (n->n), 1 coeff_to_lagrange (n->n), 2 coeff_to_extended (n->2^ek),
//        3 extended_to_coeff (2^ek -> n*(j-1)), 4 divide_by_vanishing_poly (2^ek in place)
int orc_domain_op(unsigned j, unsigned k, int which, const u64* in, u64* out, unsigned threads) {
    ORC_TRY
    EvaluationDomain d(j, k); d.threads = threads;
    size_t in_len = which <= 2 ? d.n : d.extended_len();
    std::vector<Fr> a(in_len);
    memcpy(a.data(), in, in_len * 32);
    std::vector<Fr> r;
    switch (which) {
        case 0: r = d.lagrange_to_coeff(a); break;
        case 1: r = d.coeff_to_lagrange(a); break;
        case 2: r = d.coeff_to_extended(a); break;
        case 3: r = d.extended_to_coeff(a); break;
        default: d.divide_by_vanishing_poly(a); r = a;
    }
    memcpy(out, r.data(), r.size() * 32);
    ORC_CATCH
}
int orc_g_to_lagrange(const u64* g, unsigned k, u64* out, unsigned threads) {
    ORC_TRY
    size_t n = (size_t)1 << k;
    std::vector<G1Affine> in(n);
    for (size_t i = 0; i < n; ++i) in[i] = ldp(g + 8 * i);
    auto r = g_to_lagrange(in, k, threads);
    for (size_t i = 0; i < n; ++i) stp(out + 8 * i, r[i]);
    ORC_CATCH
}
int orc_eval_polynomial(const u64* poly, size_t n, const u64* x, u64* out) {
    ORC_TRY
    st(out, eval_polynomial(reinterpret_cast<const Fr*>(poly), n, ld<Fr>(x)));
    ORC_CATCH
}

int orc_keccak256(const uint8_t* in, size_t len, uint8_t* out) { keccak256(in, len, out); return 0; }
int orc_smallrng(u64 seed, u64* out, size_t n) { SmallRng r(seed); for (size_t i = 0; i < n; ++i) out[i] = r.next_u64(); return 0; }
int orc_chacha20(const uint8_t* seed, u64* out, size_t n) { ChaCha20Rng r(seed); for (size_t i = 0; i < n; ++i) out[i] = r.next_u64(); return 0; }
// n uniform Fr from SmallRng(seed) by `Fr::random`
int orc_random_fr(u64 seed, u64* out, size_t n) {
    SmallRng r(seed);
    for (size_t i = 0; i < n; ++i) st(out + 4 * i, random_field<Fr>(r));
    return 0;
}

// the same from a running SmallRng (state advanced in place)
int orc_random_fr_rng(u64* state, u64* out, size_t n) {
    SmallRng r(state);
    for (size_t i = 0; i < n; ++i) st(out + 4 * i, random_field<Fr>(r));
    memcpy(state, r.s, 32);
    return 0;
}

// SRS: format 0 = Raw, 1 = PerpetualPowersOfTau.  First call with g==NULL to get k.
// g2s receives g2 ‖ s_g2 as 2 x (x0,x1,y0,y1) Mont-LE (32 u64).  g_lagrange is filled for Raw only.
int orc_srs_read(const char* path, int format, unsigned* k, u64* g, u64* g_lagrange, u64* g2s) {
    ORC_TRY
    auto buf = read_file(path);
    Srs s = format == 0 ? srs_read_raw(buf) : srs_read_ptau(buf);
    *k = s.k;
    if (g) for (size_t i = 0; i < s.g.size(); ++i) stp(g + 8 * i, s.g[i]);
    if (g_lagrange) for (size_t i = 0; i < s.g_lagrange.size(); ++i) stp(g_lagrange + 8 * i, s.g_lagrange[i]);
    if (g2s) {
        const G2AffineRaw* q[2] = {&s.g2, &s.s_g2};
        for (int i = 0; i < 2; ++i) { st(g2s + 16 * i, q[i]->x0); st(g2s + 16 * i + 4, q[i]->x1); st(g2s + 16 * i + 8, q[i]->y0); st(g2s + 16 * i + 12, q[i]->y1); }
    }
    ORC_CATCH
}

// pairing check  e(p1, q1) * e(p2, q2) == 1   (G1 affine 8 u64, G2 affine 16 u64 each, Mont-LE)
int orc_pairing_check(const u64* p1, const u64* q1, const u64* p2, const u64* q2, int* ok) {
    ORC_TRY
    G2AffineRaw a = {ld<Fq>(q1), ld<Fq>(q1 + 4), ld<Fq>(q1 + 8), ld<Fq>(q1 + 12)};
    G2AffineRaw b = {ld<Fq>(q2), ld<Fq>(q2 + 4), ld<Fq>(q2 + 8), ld<Fq>(q2 + 12)};
    *ok = pairing_product_is_one(ldp(p1), a, ldp(p2), b) ? 1 : 0;
    ORC_CATCH
}
int orc_g2_on_curve(const u64* q) {
    G2AffineRaw a = {ld<Fq>(q), ld<Fq>(q + 4), ld<Fq>(q + 8), ld<Fq>(q + 12)};
    return g2_on_curve(a) ? 1 : 0;
}

}  // extern "C"

#include "plonk_capi.inc"
