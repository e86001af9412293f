// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
//
// CPU restatement of the MSM / FFT / evaluation-domain routines on Shielder's prover hot path.
// They live in un-vendored dependencies — halo2curves 0.6.1 `msm.rs`/`fft.rs`
// (/root/reference/Cargo.lock:2367-2368) and halo2_proofs v0.3.0 `arithmetic.rs`, `poly/domain.rs`,
// `poly/kzg/commitment.rs` (/root/reference/Cargo.lock:2332-2334) — so the published algorithms are
// restated here [UPSTREAM-MEMORY, SURVEY.md §8a rows a3-a8, Appendix A] and pinned through the
// reference's own call sites and fixtures:
//   * /root/reference/crates/powers-of-tau/lib.rs:71  (from_parts -> g_to_lagrange, G1 iFFT)
//   * /root/reference/crates/powers-of-tau/lib.rs:248-264 (commit(iNTT(a)) == commit_lagrange(a))
//   * resources/ppot_0080_11_raw g_lagrange block (2048 full-width MSM known answers)
#pragma once
#include "bn254.hpp"
#include <cmath>
#include <thread>
#include <functional>

namespace oracle {

// ------------------------------------------------------------------------------------------
// parallel helper: `parallelize` / rayon scope with contiguous chunks
// ------------------------------------------------------------------------------------------
static inline void parallel_chunks(size_t n, unsigned threads, const std::function<void(size_t, size_t)>& f) {
    if (threads <= 1 || n < 2 * threads) { f(0, n); return; }
    std::vector<std::thread> ts;
    size_t chunk = (n + threads - 1) / threads;
    for (size_t s = 0; s < n; s += chunk) {
        size_t e = s + chunk < n ? s + chunk : n;
        ts.emplace_back([=, &f] { f(s, e); });
    }
    for (auto& t : ts) t.join();
}

// ------------------------------------------------------------------------------------------
// multiexp_serial — Pippenger as halo2curves 0.6.1 msm.rs: unsigned c-bit windows,
// c = 1 (n<4), 3 (n<32), else ceil(ln n); segments = 254/c + 1 walked MSB->LSB with c doublings;
// buckets in None/Affine/Projective states; running-sum reduction.
// ------------------------------------------------------------------------------------------
static inline unsigned msm_window(size_t n) {
    if (n < 4) return 1;
    if (n < 32) return 3;
    return (unsigned)std::ceil(std::log((double)(uint32_t)n));
}
static inline size_t msm_get_at(size_t segment, unsigned c, const uint8_t repr[32]) {
    size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = repr[skip_bytes + i];
    u64 tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= (u64)1 << c;
    return (size_t)tmp;
}

static inline void multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t n, G1& acc) {
    std::vector<std::array<uint8_t, 32>> reprs(n);
    for (size_t i = 0; i < n; ++i) coeffs[i].to_bytes_le(reprs[i].data());
    unsigned c = msm_window(n);
    size_t segments = 254 / c + 1;
    struct Bucket { int state; G1Affine a; G1 p; };  // 0 none, 1 affine, 2 projective
    std::vector<Bucket> buckets(((size_t)1 << c) - 1);
    for (size_t seg = segments; seg-- > 0;) {
        for (unsigned i = 0; i < c; ++i) acc = acc.dbl();
        for (auto& b : buckets) b.state = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t d = msm_get_at(seg, c, reprs[i].data());
            if (!d) continue;
            Bucket& b = buckets[d - 1];
            if (b.state == 0) { b.a = bases[i]; b.state = 1; }
            else if (b.state == 1) { b.p = G1::from_affine(b.a).add_mixed(bases[i]); b.state = 2; }
            else b.p = b.p.add_mixed(bases[i]);
        }
        G1 running = G1::identity();
        for (size_t j = buckets.size(); j-- > 0;) {
            Bucket& b = buckets[j];
            if (b.state == 1) running = running.add_mixed(b.a);
            else if (b.state == 2) running = running.add(b.p);
            acc = acc.add(running);
        }
    }
}

// best_multiexp — chunk over `threads` exactly as halo2curves does: chunk = n / threads,
// `coeffs.chunks(chunk)` (may yield threads+1 chunks), serial MSM per chunk, fold.
static inline G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t n, unsigned threads = 1) {
    if (threads < 1) threads = 1;
    if (n > threads) {
        size_t chunk = n / threads;
        size_t num_chunks = (n + chunk - 1) / chunk;
        std::vector<G1> results(num_chunks, G1::identity());
        std::vector<std::thread> ts;
        for (size_t ci = 0; ci < num_chunks; ++ci) {
            size_t s = ci * chunk, e = s + chunk < n ? s + chunk : n;
            if (threads == 1) multiexp_serial(coeffs + s, bases + s, e - s, results[ci]);
            else ts.emplace_back([=, &results] { multiexp_serial(coeffs + s, bases + s, e - s, results[ci]); });
        }
        for (auto& t : ts) t.join();
        G1 acc = G1::identity();
        for (auto& r : results) acc = acc.add(r);
        return acc;
    }
    G1 acc = G1::identity();
    multiexp_serial(coeffs, bases, n, acc);
    return acc;
}

// ------------------------------------------------------------------------------------------
// best_fft — radix-2 DIT: bit-reversal swap, twiddle table w^0..w^(n/2-1), log_n butterfly
// stages (the serial branch of halo2curves fft.rs; the recursive parallel branch computes the same
// values).  Generic over the "group" so the same routine drives Fr NTTs and the G1 FFT (K6).
// ------------------------------------------------------------------------------------------
static inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
struct FrGroupOps {
    static Fr add(const Fr& a, const Fr& b) { return a + b; }
    static Fr sub(const Fr& a, const Fr& b) { return a - b; }
    static Fr scale(const Fr& a, const Fr& s) { return a * s; }
};
struct G1GroupOps {
    static G1 add(const G1& a, const G1& b) { return a.add(b); }
    static G1 sub(const G1& a, const G1& b) { return a.add(b.neg()); }
    static G1 scale(const G1& a, const Fr& s) { return a.mul(s); }
};

template <class G, class Ops>
static inline void best_fft_generic(G* a, size_t n, const Fr& omega, unsigned log_n, unsigned threads = 1) {
    for (size_t k = 0; k < n; ++k) { size_t rk = bitreverse(k, log_n); if (k < rk) std::swap(a[k], a[rk]); }
    std::vector<Fr> tw(n / 2 ? n / 2 : 1);
    Fr w = Fr::one();
    for (size_t i = 0; i < n / 2; ++i) { tw[i] = w; w = w * omega; }
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (unsigned s = 0; s < log_n; ++s) {
        size_t half = chunk / 2, nblocks = n / chunk;
        auto body = [&](size_t bs, size_t be) {
            for (size_t blk = bs; blk < be; ++blk) {
                G* left = a + blk * chunk; G* right = left + half;
                { G t = right[0]; right[0] = Ops::sub(left[0], t); left[0] = Ops::add(left[0], t); }
                for (size_t i = 1; i < half; ++i) {
                    G t = Ops::scale(right[i], tw[i * twiddle_chunk]);
                    right[i] = Ops::sub(left[i], t);
                    left[i] = Ops::add(left[i], t);
                }
            }
        };
        if (threads > 1 && nblocks >= threads) parallel_chunks(nblocks, threads, body);
        else if (threads > 1 && nblocks == 1 && half >= 4 * threads) {
            // few big blocks: split the inner loop instead
            G* left = a; G* right = a + half;
            parallel_chunks(half, threads, [&](size_t is, size_t ie) {
                for (size_t i = is; i < ie; ++i) {
                    G t = i ? Ops::scale(right[i], tw[i * twiddle_chunk]) : right[i];
                    right[i] = Ops::sub(left[i], t);
                    left[i] = Ops::add(left[i], t);
                }
            });
        } else body(0, nblocks);
        chunk *= 2; twiddle_chunk /= 2;
    }
}
static inline void best_fft(Fr* a, size_t n, const Fr& omega, unsigned log_n, unsigned threads = 1) {
    best_fft_generic<Fr, FrGroupOps>(a, n, omega, log_n, threads);
}
static inline void best_fft_g1(G1* a, size_t n, const Fr& omega, unsigned log_n, unsigned threads = 1) {
    best_fft_generic<G1, G1GroupOps>(a, n, omega, log_n, threads);
}

// ------------------------------------------------------------------------------------------
// EvaluationDomain — halo2_proofs v0.3.0 poly/domain.rs [UPSTREAM-MEMORY; SURVEY Appendix A].
// Reference call sites: crates/powers-of-tau/lib.rs:255-260; halo2-verifier codegen.rs:161-171.
// ------------------------------------------------------------------------------------------
struct EvaluationDomain {
    unsigned k, extended_k, quotient_poly_degree;
    size_t n;
    Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv;
    Fr ifft_divisor, extended_ifft_divisor, barycentric_weight;
    std::vector<Fr> t_evaluations;  // stored inverted, length 2^(extended_k-k)
    unsigned threads = 1;

    EvaluationDomain(unsigned j, unsigned k_) : k(k_) {
        quotient_poly_degree = j - 1;
        n = (size_t)1 << k;
        extended_k = k;
        while (((size_t)1 << extended_k) < n * quotient_poly_degree) extended_k++;
        extended_omega = FrConst::root_of_unity();
        for (unsigned i = extended_k; i < FrConst::S; ++i) extended_omega = extended_omega.square();
        omega = extended_omega;
        for (unsigned i = k; i < extended_k; ++i) omega = omega.square();
        omega_inv = omega.inv();
        extended_omega_inv = extended_omega.inv();
        g_coset = FrConst::zeta();
        g_coset_inv = g_coset.square();
        {
            Fr orig = g_coset.pow_u64(n), step = extended_omega.pow_u64(n), cur = orig;
            do { t_evaluations.push_back(cur); cur = cur * step; } while (cur != orig);
            for (auto& t : t_evaluations) t = t - Fr::one();
            batch_invert(t_evaluations.data(), t_evaluations.size());
        }
        ifft_divisor = Fr::from_u64((u64)1 << k).inv();
        extended_ifft_divisor = Fr::from_u64((u64)1 << extended_k).inv();
        barycentric_weight = Fr::from_u64(n).inv();
    }
    size_t extended_len() const { return (size_t)1 << extended_k; }

    void ifft(Fr* a, size_t len, const Fr& w_inv, unsigned log_n, const Fr& divisor) const {
        best_fft(a, len, w_inv, log_n, threads);
        for (size_t i = 0; i < len; ++i) a[i] = a[i] * divisor;
    }
    std::vector<Fr> lagrange_to_coeff(std::vector<Fr> a) const {
        ifft(a.data(), n, omega_inv, k, ifft_divisor); return a;
    }
    std::vector<Fr> coeff_to_lagrange(std::vector<Fr> a) const {
        best_fft(a.data(), n, omega, k, threads); return a;
    }
    // a[i] *= zeta^(i mod 3) (into coset) or zeta^(-(i mod 3)) (out of coset); zeta^3 = 1.
    void distribute_powers_zeta(std::vector<Fr>& a, bool into_coset) const {
        Fr p1 = into_coset ? g_coset : g_coset_inv, p2 = into_coset ? g_coset_inv : g_coset;
        for (size_t i = 0; i < a.size(); ++i) {
            size_t j = i % 3;
            if (j == 1) a[i] = a[i] * p1; else if (j == 2) a[i] = a[i] * p2;
        }
    }
    std::vector<Fr> coeff_to_extended(std::vector<Fr> a) const {
        distribute_powers_zeta(a, true);
        a.resize(extended_len(), Fr::zero());
        best_fft(a.data(), a.size(), extended_omega, extended_k, threads);
        return a;
    }
    std::vector<Fr> extended_to_coeff(std::vector<Fr> a) const {
        ifft(a.data(), a.size(), extended_omega_inv, extended_k, extended_ifft_divisor);
        distribute_powers_zeta(a, false);
        a.resize(n * quotient_poly_degree);
        return a;
    }
    void divide_by_vanishing_poly(std::vector<Fr>& a) const {
        size_t m = t_evaluations.size();
        for (size_t i = 0; i < a.size(); ++i) a[i] = a[i] * t_evaluations[i % m];
    }
    Fr rotate_omega(const Fr& v, int rotation) const {
        Fr p = rotation >= 0 ? omega.pow_u64((u64)rotation) : omega_inv.pow_u64((u64)(-(long)rotation));
        return v * p;
    }
};

// g_to_lagrange — halo2_proofs poly/kzg/commitment.rs (reached from ParamsKZG::from_parts with
// g_lagrange=None, crates/powers-of-tau/lib.rs:71): n^{-1} * FFT_{omega^{-1}}(g), then normalise.
static inline std::vector<G1Affine> g_to_lagrange(const std::vector<G1Affine>& g, unsigned k, unsigned threads = 1) {
    size_t n = (size_t)1 << k;
    std::vector<G1> p(n);
    for (size_t i = 0; i < n; ++i) p[i] = G1::from_affine(g[i]);
    Fr omega_inv = FrConst::root_of_unity().inv();
    for (unsigned i = k; i < FrConst::S; ++i) omega_inv = omega_inv.square();
    Fr n_inv = Fr::from_u64((u64)1 << k).inv();
    best_fft_g1(p.data(), n, omega_inv, k, threads);
    parallel_chunks(n, threads, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) p[i] = p[i].mul(n_inv); });
    std::vector<G1Affine> out(n);
    batch_normalize(p.data(), out.data(), n);
    return out;
}

// eval_polynomial / kate_division — halo2_proofs arithmetic.rs
static inline Fr eval_polynomial(const Fr* poly, size_t n, const Fr& x) {
    Fr acc = Fr::zero();
    for (size_t i = n; i-- > 0;) acc = acc * x + poly[i];
    return acc;
}
// quotient of a(X) by (X - b), discarding the remainder; returns n-1 coefficients
static inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
    std::vector<Fr> q(a.size() ? a.size() - 1 : 0);
    Fr tmp = Fr::zero();
    for (size_t i = a.size(); i-- > 1;) {
        Fr lead = a[i] + tmp;
        q[i - 1] = lead;
        tmp = lead * b;
    }
    return q;
}

}  // namespace oracle
