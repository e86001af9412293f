// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
//
// CPU restatement of the PLONK prover / verifier that Shielder reaches through
//   shielder_circuits::generate_proof -> halo2_proofs::plonk::create_proof::<KZGCommitmentScheme<Bn256>,
//   ProverSHPLONK, ChallengeEvm, _, Keccak256Transcript, _>
// (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111).  halo2_proofs v0.3.0 and the
// zkOS-circuits `transcript` crate are un-vendored (Cargo.lock:2332-2334, 5956-5964), so:
//   * the PROVER follows the published algorithm of halo2 v0.3.0 plonk/prover.rs, permutation/,
//     vanishing/, evaluation.rs and poly/kzg/multiopen/shplonk/prover.rs [UPSTREAM-MEMORY; SURVEY.md
//     §3.2, Appendix A], single phase, one instance column, no lookups (asserted);
//   * the VERIFIER follows the in-repo spec line by line: proof layout and transcript
//     (crates/halo2-verifier/src/lib/codegen/util.rs:133-245, templates/Halo2Verifier.sol:89-124,
//     247-307), Lagrange/instance evaluation (Halo2Verifier.sol:392-470), quotient numerator
//     (codegen/evaluator.rs:45-131, codegen.rs:237-254), quotient commitment (Halo2Verifier.sol:494-512),
//     SHPLONK pairing inputs (codegen/pcs.rs:60-104, pcs/bdfg21.rs:21-494) and the final pairing.
// "Parity unpinned" for proof BYTES: the reference holds no golden proofs and the real circuits are
// unavailable; what is pinned is acceptance by this verifier restatement, plus
// byte-equality between this CPU prover and the GPU prover on the same seed.
// The vk digest (halo2: Blake2b of the pinned constraint system's Debug output) cannot be
// reproduced here; it is an opaque 32-byte value derived with Keccak from the circuit description.
#pragma once
#include "bn254.hpp"
#include "arith.hpp"
#include "misc.hpp"
#include "pairing.hpp"
#include <map>
#include <set>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Constraint system description ("shape")
// ---------------------------------------------------------------------------------------------
enum ExprOp : uint32_t { OP_CONST = 0, OP_FIXED = 1, OP_ADVICE = 2, OP_INSTANCE = 3, OP_NEG = 4, OP_ADD = 5, OP_MUL = 6, OP_SCALE = 7 };
struct ExprIns { uint32_t op, arg; };
typedef std::vector<ExprIns> Expr;  // postfix; query ops index {fixed,advice,instance}_queries, CONST/SCALE index constants
enum ColType : uint32_t { COL_ADVICE = 0, COL_FIXED = 1, COL_INSTANCE = 2 };
struct ColumnRef { uint32_t type, index; bool operator==(const ColumnRef& o) const { return type == o.type && index == o.index; } };
struct Query { uint32_t column; int32_t rotation; };
struct Copy { uint32_t lcol, lrow, rcol, rrow; };  // columns index into perm_columns

struct ConstraintSystem {
    uint32_t k = 0, num_fixed = 0, num_advice = 0, num_instance = 0;
    std::vector<Query> advice_queries, fixed_queries, instance_queries;
    std::vector<Expr> gates;
    std::vector<Fr> constants;
    std::vector<ColumnRef> perm_columns;
    struct Lookup { std::vector<Expr> inputs, tables; };
    std::vector<Lookup> lookups;
    size_t num_lookups() const { return lookups.size(); }

    size_t n() const { return (size_t)1 << k; }
    static unsigned expr_degree(const Expr& e) {
        std::vector<unsigned> st;
        for (auto& i : e) switch (i.op) {
            case OP_CONST: st.push_back(0); break;
            case OP_FIXED: case OP_ADVICE: case OP_INSTANCE: st.push_back(1); break;
            case OP_NEG: case OP_SCALE: break;
            case OP_ADD: { unsigned b = st.back(); st.pop_back(); st.back() = std::max(st.back(), b); break; }
            case OP_MUL: { unsigned b = st.back(); st.pop_back(); st.back() += b; break; }
        }
        return st.empty() ? 0 : st.back();
    }
    // ConstraintSystem::degree(): permutation needs 3, gates their own degree
    // ConstraintSystem::degree(): permutation needs 3, a lookup max(4, 2 + input_degree + table_degree), gates their own
    unsigned degree() const {
        unsigned d = 3;   // permutation.required_degree(), whether or not a column is copy-enabled (halo2 ConstraintSystem::degree)
        for (auto& l : lookups) {
            unsigned di = 1, dt = 1;
            for (auto& e : l.inputs) di = std::max(di, expr_degree(e));
            for (auto& e : l.tables) dt = std::max(dt, expr_degree(e));
            d = std::max(d, std::max(4u, 2 + di + dt));
        }
        for (auto& g : gates) d = std::max(d, expr_degree(g));
        return d;
    }
    // ConstraintSystem::blinding_factors(): max(3, max #queries of one advice column) + 2
    unsigned blinding_factors() const {
        std::vector<unsigned> cnt(num_advice, 0);
        for (auto& q : advice_queries) cnt[q.column]++;
        unsigned f = 1;
        for (unsigned c : cnt) f = std::max(f, c);
        return std::max(3u, f) + 2;
    }
    unsigned chunk_len() const { return degree() - 2; }
    unsigned num_perm_sets() const { return perm_columns.empty() ? 0 : (unsigned)((perm_columns.size() + chunk_len() - 1) / chunk_len()); }
    unsigned num_quotients() const { return degree() - 1; }
    int rotation_last() const { return -(int)(blinding_factors() + 1); }
    size_t usable_rows() const { return n() - (blinding_factors() + 1); }
    size_t num_evals() const {
        return advice_queries.size() + fixed_queries.size() + 1 + perm_columns.size() + (num_perm_sets() ? 3 * num_perm_sets() - 1 : 0) + 5 * num_lookups();
    }
    // crates/halo2-verifier/src/lib/codegen/util.rs:175-186
    size_t proof_len() const {
        return 64 * (num_advice + 3 * num_lookups() + num_perm_sets() + 1 + num_quotients()) + 32 * num_evals() + 128;
    }
};

struct Circuit {
    ConstraintSystem cs;
    std::vector<std::vector<Fr>> fixed;  // num_fixed x n (rows beyond the assignment are zero)
    std::vector<Copy> copies;
    std::vector<uint8_t> blob;           // serialised form (digest input)
};

// Blob layout (little-endian u32 unless noted): see tests/circuits.py `serialize`.
struct BlobReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t u32() { if (p + 4 > end) throw std::runtime_error("circuit blob: truncated"); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    Fr fr() { if (p + 32 > end) throw std::runtime_error("circuit blob: truncated"); u64 l[4]; memcpy(l, p, 32); p += 32; return Fr::from_raw_mont(l); }
};
static inline Circuit parse_circuit(const uint8_t* data, size_t len) {
    Circuit c; BlobReader r{data, data + len};
    if (r.u32() != 0x5a4b4353) throw std::runtime_error("circuit blob: bad magic");
    ConstraintSystem& cs = c.cs;
    cs.k = r.u32(); cs.num_fixed = r.u32(); cs.num_advice = r.u32(); cs.num_instance = r.u32();
    auto rq = [&](std::vector<Query>& v) { uint32_t m = r.u32(); v.resize(m); for (auto& q : v) { q.column = r.u32(); q.rotation = (int32_t)r.u32(); } };
    rq(cs.advice_queries); rq(cs.fixed_queries); rq(cs.instance_queries);
    uint32_t nc = r.u32(); cs.constants.resize(nc); for (auto& f : cs.constants) f = r.fr();
    uint32_t ng = r.u32(); cs.gates.resize(ng);
    for (auto& g : cs.gates) { uint32_t m = r.u32(); g.resize(m); for (auto& i : g) { i.op = r.u32(); i.arg = r.u32(); } }
    uint32_t np = r.u32(); cs.perm_columns.resize(np); for (auto& pc : cs.perm_columns) { pc.type = r.u32(); pc.index = r.u32(); }
    auto rexpr = [&](Expr& g) { uint32_t m = r.u32(); g.resize(m); for (auto& i : g) { i.op = r.u32(); i.arg = r.u32(); } };
    cs.lookups.resize(r.u32());
    for (auto& l : cs.lookups) {
        l.inputs.resize(r.u32()); for (auto& e : l.inputs) rexpr(e);
        l.tables.resize(r.u32()); for (auto& e : l.tables) rexpr(e);
        if (l.inputs.empty() || l.inputs.size() != l.tables.size()) throw std::runtime_error("lookup: input/table expression counts differ");
    }
    if (cs.num_instance != 1) throw std::runtime_error("exactly one instance column is supported (as Shielder's circuits)");
    size_t n = cs.n();
    c.fixed.assign(cs.num_fixed, std::vector<Fr>(n, Fr::zero()));
    for (auto& col : c.fixed) for (size_t i = 0; i < n; ++i) col[i] = r.fr();
    uint32_t ncp = r.u32(); c.copies.resize(ncp);
    for (auto& cp : c.copies) { cp.lcol = r.u32(); cp.lrow = r.u32(); cp.rcol = r.u32(); cp.rrow = r.u32(); }
    if (r.p != r.end) throw std::runtime_error("circuit blob: trailing bytes");
    c.blob.assign(data, data + len);
    return c;
}

// ---------------------------------------------------------------------------------------------
// Keccak256Transcript (zkOS-circuits `transcript` crate; spec Halo2Verifier.sol:101-124,247-307)
// ---------------------------------------------------------------------------------------------
struct Transcript {
    std::vector<uint8_t> buf;    // pending hash input: previous hash (32 B) ‖ absorbed data
    std::vector<uint8_t> proof;  // written points / scalars
    const uint8_t* rd = nullptr; const uint8_t* rd_end = nullptr;  // reader side
    bool fresh_hash = false;     // buf is exactly the previous hash

    void common_bytes(const uint8_t* b, size_t n) { buf.insert(buf.end(), b, b + n); fresh_hash = false; }
    void common_scalar(const Fr& s) { uint8_t w[32]; s.to_bytes_be(w); common_bytes(w, 32); }
    void common_point(const G1Affine& p) { uint8_t w[64]; p.x.to_bytes_be(w); p.y.to_bytes_be(w + 32); common_bytes(w, 64); }
    void write_scalar(const Fr& s) { uint8_t w[32]; s.to_bytes_be(w); proof.insert(proof.end(), w, w + 32); common_bytes(w, 32); }
    void write_point(const G1Affine& p) {
        uint8_t w[64]; p.x.to_bytes_be(w); p.y.to_bytes_be(w + 32);
        proof.insert(proof.end(), w, w + 64); common_bytes(w, 64);
    }
    Fr squeeze_challenge() {
        if (fresh_hash) buf.push_back(0x01);
        uint8_t h[32]; keccak256(buf.data(), buf.size(), h);
        buf.assign(h, h + 32); fresh_hash = true;
        // challenge = hash (big-endian integer) mod r
        u64 w[8] = {0};
        for (int i = 0; i < 32; ++i) w[i / 8] |= (u64)h[31 - i] << (8 * (i % 8));
        return Fr::from_u512(w);
    }
    // reader
    static bool word_to_fq(const uint8_t* w, Fq& out) { uint8_t le[32]; for (int i = 0; i < 32; ++i) le[i] = w[31 - i]; return Fq::from_bytes_le(le, out); }
    static bool word_to_fr(const uint8_t* w, Fr& out) { uint8_t le[32]; for (int i = 0; i < 32; ++i) le[i] = w[31 - i]; return Fr::from_bytes_le(le, out); }
    bool read_point(G1Affine& p) {  // read_ec_point: x,y < q and on curve (Halo2Verifier.sol:89-99)
        if (rd + 64 > rd_end) return false;
        bool ok = word_to_fq(rd, p.x) && word_to_fq(rd + 32, p.y);
        if (ok) ok = p.y.square() == p.x.square() * p.x + Fq::from_u64(3);
        common_bytes(rd, 64); rd += 64;
        return ok;
    }
    bool read_scalar(Fr& s) {
        if (rd + 32 > rd_end) return false;
        bool ok = word_to_fr(rd, s);
        common_bytes(rd, 32); rd += 32;
        return ok;
    }
};

// ---------------------------------------------------------------------------------------------
// Keys
// ---------------------------------------------------------------------------------------------
struct ParamsKZG {
    unsigned k; std::vector<G1Affine> g, g_lagrange; G2AffineRaw g2, s_g2; unsigned threads = 1;
    G1Affine commit(const std::vector<Fr>& coeffs) const { return best_multiexp(coeffs.data(), g.data(), coeffs.size(), threads).to_affine(); }
    G1Affine commit_lagrange(const std::vector<Fr>& vals) const { return best_multiexp(vals.data(), g_lagrange.data(), vals.size(), threads).to_affine(); }
};

struct VerifyingKey {
    ConstraintSystem cs;
    std::vector<G1Affine> fixed_commitments, perm_commitments;
    Fr transcript_repr;  // opaque digest, see header
};
struct ProvingKey {
    VerifyingKey vk;
    EvaluationDomain domain;
    std::vector<std::vector<Fr>> fixed_values, fixed_polys, fixed_cosets;
    std::vector<std::vector<Fr>> perm_values, perm_polys, perm_cosets;  // sigma columns
    std::vector<Fr> l0, l_last, l_active_row;                           // extended-domain cosets
    ProvingKey(const ConstraintSystem& cs) : domain(cs.degree(), cs.k) {}
};

static inline Fr digest_of(const Circuit& c, const std::vector<G1Affine>& fc, const std::vector<G1Affine>& pc) {
    std::vector<uint8_t> in(c.blob);
    auto add = [&](const G1Affine& p) { uint8_t w[64]; p.x.to_bytes_be(w); p.y.to_bytes_be(w + 32); in.insert(in.end(), w, w + 64); };
    for (auto& p : fc) add(p);
    for (auto& p : pc) add(p);
    uint8_t h[32]; keccak256(in.data(), in.size(), h);
    u64 w[8] = {0};
    for (int i = 0; i < 32; ++i) w[i / 8] |= (u64)h[31 - i] << (8 * (i % 8));
    return Fr::from_u512(w);
}

// permutation::keygen::Assembly — cycle-merging copy constraints [UPSTREAM-MEMORY]
struct PermAssembly {
    size_t ncols, n;
    std::vector<std::pair<uint32_t, uint32_t>> mapping, aux;
    std::vector<uint32_t> sizes;
    PermAssembly(size_t ncols_, size_t n_) : ncols(ncols_), n(n_), mapping(ncols_ * n_), aux(ncols_ * n_), sizes(ncols_ * n_, 1) {
        for (size_t c = 0; c < ncols; ++c) for (size_t r = 0; r < n; ++r) mapping[c * n + r] = aux[c * n + r] = {(uint32_t)c, (uint32_t)r};
    }
    size_t at(std::pair<uint32_t, uint32_t> p) const { return (size_t)p.first * n + p.second; }
    void copy(uint32_t lc, uint32_t lr, uint32_t rc, uint32_t rr) {
        if (lc >= ncols || rc >= ncols || lr >= n || rr >= n) throw std::runtime_error("copy constraint out of bounds");
        auto left = aux[(size_t)lc * n + lr], right = aux[(size_t)rc * n + rr];
        if (left == right) return;
        if (sizes[at(left)] < sizes[at(right)]) std::swap(left, right);
        sizes[at(left)] += sizes[at(right)];
        auto i = right;
        do { aux[at(i)] = left; i = mapping[at(i)]; } while (i != right);
        std::swap(mapping[(size_t)lc * n + lr], mapping[(size_t)rc * n + rr]);
    }
};

static inline ProvingKey keygen(const ParamsKZG& params, const Circuit& c) {
    const ConstraintSystem& cs = c.cs;
    if (params.k != cs.k) throw std::runtime_error("keygen: params.k != circuit k");
    ProvingKey pk(cs);
    pk.domain.threads = params.threads;
    const EvaluationDomain& d = pk.domain;
    size_t n = cs.n();
    pk.vk.cs = cs;
    pk.fixed_values = c.fixed;
    for (auto& col : pk.fixed_values) {
        pk.vk.fixed_commitments.push_back(params.commit_lagrange(col));
        pk.fixed_polys.push_back(d.lagrange_to_coeff(col));
        pk.fixed_cosets.push_back(d.coeff_to_extended(pk.fixed_polys.back()));
    }
    // sigma columns: value delta^col * omega^row of the mapped cell
    PermAssembly as(cs.perm_columns.size(), n);
    for (auto& cp : c.copies) as.copy(cp.lcol, cp.lrow, cp.rcol, cp.rrow);
    std::vector<Fr> omega_pows(n);
    { Fr w = Fr::one(); for (size_t i = 0; i < n; ++i) { omega_pows[i] = w; w = w * d.omega; } }
    std::vector<Fr> delta_pows(cs.perm_columns.size());
    { Fr dl = Fr::one(); for (auto& x : delta_pows) { x = dl; dl = dl * FrConst::delta(); } }
    for (size_t col = 0; col < cs.perm_columns.size(); ++col) {
        std::vector<Fr> v(n);
        for (size_t row = 0; row < n; ++row) { auto m = as.mapping[col * n + row]; v[row] = delta_pows[m.first] * omega_pows[m.second]; }
        pk.vk.perm_commitments.push_back(params.commit_lagrange(v));
        pk.perm_polys.push_back(d.lagrange_to_coeff(v));
        pk.perm_cosets.push_back(d.coeff_to_extended(pk.perm_polys.back()));
        pk.perm_values.push_back(std::move(v));
    }
    unsigned bf = cs.blinding_factors();
    std::vector<Fr> l0(n, Fr::zero()), lblind(n, Fr::zero()), llast(n, Fr::zero());
    l0[0] = Fr::one();
    for (size_t i = n - bf; i < n; ++i) lblind[i] = Fr::one();
    llast[n - bf - 1] = Fr::one();
    pk.l0 = d.coeff_to_extended(d.lagrange_to_coeff(l0));
    std::vector<Fr> lb = d.coeff_to_extended(d.lagrange_to_coeff(lblind));
    pk.l_last = d.coeff_to_extended(d.lagrange_to_coeff(llast));
    pk.l_active_row.resize(d.extended_len());
    for (size_t i = 0; i < d.extended_len(); ++i) pk.l_active_row[i] = Fr::one() - pk.l_last[i] - lb[i];
    pk.vk.transcript_repr = digest_of(c, pk.vk.fixed_commitments, pk.vk.perm_commitments);
    return pk;
}

// ---------------------------------------------------------------------------------------------
// Expression evaluation
// ---------------------------------------------------------------------------------------------
template <class FixedF, class AdviceF, class InstF>
static inline Fr eval_expr(const Expr& e, const std::vector<Fr>& consts, FixedF fx, AdviceF ad, InstF in) {
    Fr st[32]; int sp = 0;
    for (auto& i : e) switch (i.op) {
        case OP_CONST: st[sp++] = consts[i.arg]; break;
        case OP_FIXED: st[sp++] = fx(i.arg); break;
        case OP_ADVICE: st[sp++] = ad(i.arg); break;
        case OP_INSTANCE: st[sp++] = in(i.arg); break;
        case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case OP_ADD: st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; break;
        case OP_MUL: st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; break;
        case OP_SCALE: st[sp - 1] = st[sp - 1] * consts[i.arg]; break;
    }
    return st[0];
}

// MockProver-style check that a witness satisfies gates (usable rows) and copy constraints.
static inline std::string check_witness(const Circuit& c, const std::vector<std::vector<Fr>>& advice, const std::vector<Fr>& instance) {
    const ConstraintSystem& cs = c.cs; size_t n = cs.n(), usable = cs.usable_rows();
    std::vector<Fr> inst(n, Fr::zero());
    for (size_t i = 0; i < instance.size(); ++i) inst[i] = instance[i];
    auto rot = [&](size_t row, int r) { return (size_t)(((long)row + r) % (long)n + (long)n) % n; };
    for (size_t g = 0; g < cs.gates.size(); ++g)
        for (size_t row = 0; row < usable; ++row) {
            Fr v = eval_expr(cs.gates[g], cs.constants,
                [&](uint32_t q) { return c.fixed[cs.fixed_queries[q].column][rot(row, cs.fixed_queries[q].rotation)]; },
                [&](uint32_t q) { return advice[cs.advice_queries[q].column][rot(row, cs.advice_queries[q].rotation)]; },
                [&](uint32_t q) { return inst[rot(row, cs.instance_queries[q].rotation)]; });
            if (!v.is_zero()) return "gate " + std::to_string(g) + " not satisfied at row " + std::to_string(row);
        }
    auto cell = [&](uint32_t pc, uint32_t row) -> Fr {
        const ColumnRef& cr = cs.perm_columns[pc];
        return cr.type == COL_ADVICE ? advice[cr.index][row] : cr.type == COL_FIXED ? c.fixed[cr.index][row] : inst[row];
    };
    for (auto& cp : c.copies)
        if (cell(cp.lcol, cp.lrow) != cell(cp.rcol, cp.rrow)) return "copy constraint violated";
    // lookups: every input tuple of a usable row appears among the table tuples of the usable rows
    for (size_t l = 0; l < cs.lookups.size(); ++l) {
        auto tuple_at = [&](const std::vector<Expr>& exprs, size_t row) {
            std::vector<U256> t;
            for (auto& e : exprs)
                t.push_back(eval_expr(e, cs.constants,
                    [&](uint32_t q) { return c.fixed[cs.fixed_queries[q].column][rot(row, cs.fixed_queries[q].rotation)]; },
                    [&](uint32_t q) { return advice[cs.advice_queries[q].column][rot(row, cs.advice_queries[q].rotation)]; },
                    [&](uint32_t q) { return inst[rot(row, cs.instance_queries[q].rotation)]; }).to_u256());
            return t;
        };
        auto less = [](const std::vector<U256>& a, const std::vector<U256>& b) {
            for (size_t i = 0; i < a.size(); ++i) { int c_ = u256_cmp(a[i], b[i]); if (c_) return c_ < 0; }
            return false;
        };
        std::set<std::vector<U256>, decltype(less)> table(less);
        for (size_t row = 0; row < usable; ++row) table.insert(tuple_at(cs.lookups[l].tables, row));
        for (size_t row = 0; row < usable; ++row)
            if (!table.count(tuple_at(cs.lookups[l].inputs, row))) return "lookup " + std::to_string(l) + " not satisfied at row " + std::to_string(row);
    }
    return "";
}

// ---------------------------------------------------------------------------------------------
// SHPLONK helpers shared by prover and verifier: the query list and rotation sets in the order of
// codegen/pcs.rs:60-104 and pcs/bdfg21.rs:443-494.
// ---------------------------------------------------------------------------------------------
struct OpenQuery { int comm; int rot; int eval; };  // comm: commitment id; eval: index into the eval list
struct RotationSet { std::vector<int> rots, diffs; std::vector<int> comms; std::vector<std::vector<int>> evals; };

// commitment ids: [0,A) advice | P permutation z | L lookup z | L permuted inputs | L permuted tables | F fixed | S sigma | H | RANDOM
struct QueryPlan {
    std::vector<OpenQuery> queries;
    std::vector<int> superset;  // sorted rotations
    std::vector<RotationSet> sets;
    int id_perm_z0, id_lk_z0, id_lk_a0, id_lk_s0, id_fixed0, id_sigma0, id_h, id_random, num_comms;
    int e_lookup0;
    // eval indices (positions in the proof's evaluation list); h eval is "computed": index = num_evals
    explicit QueryPlan(const ConstraintSystem& cs) {
        int A = cs.num_advice, P = cs.num_perm_sets(), F = cs.num_fixed, S = (int)cs.perm_columns.size(), L = (int)cs.num_lookups();
        id_perm_z0 = A; id_lk_z0 = A + P; id_lk_a0 = id_lk_z0 + L; id_lk_s0 = id_lk_a0 + L; id_fixed0 = id_lk_s0 + L; id_sigma0 = id_fixed0 + F;
        id_h = id_sigma0 + S; id_random = id_h + 1; num_comms = id_random + 1;
        int e_adv = 0, e_fix = (int)cs.advice_queries.size(), e_rand = e_fix + (int)cs.fixed_queries.size(), e_sigma = e_rand + 1, e_z = e_sigma + S;
        e_lookup0 = e_z + (P ? 3 * P - 1 : 0);
        int e_h = (int)cs.num_evals();
        for (size_t i = 0; i < cs.advice_queries.size(); ++i) queries.push_back({(int)cs.advice_queries[i].column, cs.advice_queries[i].rotation, e_adv + (int)i});
        for (int s = 0; s < P; ++s) { queries.push_back({id_perm_z0 + s, 0, e_z + 3 * s}); queries.push_back({id_perm_z0 + s, 1, e_z + 3 * s + 1}); }
        for (int s = P - 2; s >= 0; --s) queries.push_back({id_perm_z0 + s, cs.rotation_last(), e_z + 3 * s + 2});
        for (int l = 0; l < L; ++l) {  // codegen/pcs.rs:80-92
            int e = e_lookup0 + 5 * l;
            queries.push_back({id_lk_z0 + l, 0, e});
            queries.push_back({id_lk_a0 + l, 0, e + 2});
            queries.push_back({id_lk_s0 + l, 0, e + 4});
            queries.push_back({id_lk_a0 + l, -1, e + 3});
            queries.push_back({id_lk_z0 + l, 1, e + 1});
        }
        for (size_t i = 0; i < cs.fixed_queries.size(); ++i) queries.push_back({id_fixed0 + (int)cs.fixed_queries[i].column, cs.fixed_queries[i].rotation, e_fix + (int)i});
        for (int s = 0; s < S; ++s) queries.push_back({id_sigma0 + s, 0, e_sigma + s});
        queries.push_back({id_h, 0, e_h});
        queries.push_back({id_random, 0, e_rand});
        // rotation_sets
        std::set<int> sup;
        std::vector<std::pair<int, std::map<int, int>>> cq;
        for (auto& q : queries) {
            sup.insert(q.rot);
            auto it = std::find_if(cq.begin(), cq.end(), [&](auto& x) { return x.first == q.comm; });
            if (it == cq.end()) cq.push_back({q.comm, {{q.rot, q.eval}}}); else it->second[q.rot] = q.eval;
        }
        superset.assign(sup.begin(), sup.end());
        for (auto& [comm, qs] : cq) {
            std::vector<int> rots, evs;
            for (auto& [r, e] : qs) { rots.push_back(r); evs.push_back(e); }
            auto it = std::find_if(sets.begin(), sets.end(), [&](RotationSet& s) { return s.rots == rots; });
            if (it == sets.end()) {
                RotationSet s; s.rots = rots;
                for (int r : superset) if (!qs.count(r)) s.diffs.push_back(r);
                s.comms.push_back(comm); s.evals.push_back(evs);
                sets.push_back(s);
            } else { it->comms.push_back(comm); it->evals.push_back(evs); }
        }
    }
};

// Lagrange interpolation through (points[i], evals[i]) — halo2 arithmetic.rs lagrange_interpolate
static inline std::vector<Fr> lagrange_interpolate(const std::vector<Fr>& pts, const std::vector<Fr>& evals) {
    size_t m = pts.size();
    std::vector<Fr> out(m, Fr::zero());
    for (size_t j = 0; j < m; ++j) {
        // numerator poly prod_{k != j} (X - x_k), denominator prod (x_j - x_k)
        std::vector<Fr> num{Fr::one()};
        Fr den = Fr::one();
        for (size_t kx = 0; kx < m; ++kx) {
            if (kx == j) continue;
            std::vector<Fr> nn(num.size() + 1, Fr::zero());
            for (size_t t = 0; t < num.size(); ++t) { nn[t + 1] += num[t]; nn[t] -= num[t] * pts[kx]; }
            num = nn;
            den = den * (pts[j] - pts[kx]);
        }
        Fr sc = evals[j] * den.inv();
        for (size_t t = 0; t < num.size(); ++t) out[t] += num[t] * sc;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// create_proof
// ---------------------------------------------------------------------------------------------
struct ProverStats {
    size_t msms = 0, ntts = 0, ext_ntts = 0;
    // optional stage trace (tests): called with a stage name and the stage's field elements
    std::function<void(const char*, const Fr*, size_t)> trace;
};

// rayon::current_num_threads() of the prover being restated (affects only the chunking of the vanishing argument's random polynomial)
static unsigned g_rayon_threads = 1;

static inline std::vector<uint8_t> create_proof(const ParamsKZG& params, const ProvingKey& pk,
                                                std::vector<std::vector<Fr>> advice,  // num_advice x n, assigned cells
                                                const std::vector<Fr>& instance, RngCore& rng, ProverStats* stats = nullptr) {
    const ConstraintSystem& cs = pk.vk.cs;
    const EvaluationDomain& d = pk.domain;
    const size_t n = cs.n(), en = d.extended_len();
    const unsigned bf = cs.blinding_factors();
    const size_t unusable_start = n - (bf + 1);
    ProverStats st;
    if (stats) st.trace = stats->trace;
    const bool timing = getenv("ORACLE_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[oracle] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    auto trace = [&](const char* name, const std::vector<Fr>& v) { if (st.trace) st.trace(name, v.data(), v.size()); };
    auto trace2 = [&](const char* name, const std::vector<std::vector<Fr>>& vv) { if (st.trace) for (auto& v : vv) st.trace(name, v.data(), v.size()); };
    if (advice.size() != cs.num_advice) throw std::runtime_error("create_proof: wrong number of advice columns");
    for (auto& col : advice) if (col.size() != n) throw std::runtime_error("create_proof: advice column length != n");
    if (instance.size() > unusable_start) throw std::runtime_error("create_proof: InstanceTooLarge");

    Transcript tr;
    tr.common_scalar(pk.vk.transcript_repr);
    for (auto& v : instance) tr.common_scalar(v);
    std::vector<Fr> instance_values(n, Fr::zero());
    for (size_t i = 0; i < instance.size(); ++i) instance_values[i] = instance[i];
    std::vector<Fr> instance_poly = d.lagrange_to_coeff(instance_values); st.ntts++;

    // advice: blind the unusable rows (column-major), one unused Blind per column, commit
    for (auto& col : advice) for (size_t i = unusable_start; i < n; ++i) col[i] = random_field<Fr>(rng);
    for (size_t i = 0; i < advice.size(); ++i) (void)random_field<Fr>(rng);
    trace2("advice_blinded", advice);
    for (auto& col : advice) { tr.write_point(params.commit_lagrange(col)); st.msms++; }

    lap("advice blind+commit");
    Fr theta = tr.squeeze_challenge();

    // lookup arguments, part 1 (halo2 lookup/prover.rs commit_permuted): compress the input / table expressions with
    // theta over the Lagrange rows, permute the pair, blind, commit
    const size_t L = cs.num_lookups();
    std::vector<std::vector<Fr>> lk_in(L), lk_tab(L), lk_a(L), lk_s(L);  // compressed input/table, permuted input/table (values)
    {
        auto rotl = [&](size_t row, int r) { return (size_t)(((long)row + r) % (long)n + (long)n) % n; };
        auto compress = [&](const std::vector<Expr>& exprs) {
            std::vector<Fr> out(n, Fr::zero());
            for (auto& e : exprs)
                for (size_t row = 0; row < n; ++row) {
                    Fr v = eval_expr(e, cs.constants,
                        [&](uint32_t q) { return pk.fixed_values[cs.fixed_queries[q].column][rotl(row, cs.fixed_queries[q].rotation)]; },
                        [&](uint32_t q) { return advice[cs.advice_queries[q].column][rotl(row, cs.advice_queries[q].rotation)]; },
                        [&](uint32_t q) { return instance_values[rotl(row, cs.instance_queries[q].rotation)]; });
                    out[row] = out[row] * theta + v;
                }
            return out;
        };
        for (size_t l = 0; l < L; ++l) {
            lk_in[l] = compress(cs.lookups[l].inputs);
            lk_tab[l] = compress(cs.lookups[l].tables);
            // permute_expression_pair
            std::vector<Fr> a(lk_in[l].begin(), lk_in[l].begin() + unusable_start);
            // `Fr: Ord` compares canonical representations (halo2curves derive/field.rs)
            std::sort(a.begin(), a.end(), [](const Fr& x, const Fr& y) { return u256_cmp(x.to_u256(), y.to_u256()) < 0; });
            struct U256Less { bool operator()(const U256& x, const U256& y) const { return u256_cmp(x, y) < 0; } };
            std::map<U256, uint32_t, U256Less> leftover;  // BTreeMap<Fr, u32>: iteration in ascending canonical order
            for (size_t i = 0; i < unusable_start; ++i) leftover[lk_tab[l][i].to_u256()]++;
            std::vector<Fr> pt(unusable_start, Fr::zero());
            std::vector<size_t> repeated;
            for (size_t row = 0; row < unusable_start; ++row) {
                if (row == 0 || a[row] != a[row - 1]) {
                    pt[row] = a[row];
                    auto it = leftover.find(a[row].to_u256());
                    if (it == leftover.end() || it->second == 0) throw std::runtime_error("create_proof: lookup input not in table (ConstraintSystemFailure)");
                    it->second--;
                } else repeated.push_back(row);
            }
            for (auto& kv : leftover)
                for (uint32_t c = 0; c < kv.second; ++c) { pt[repeated.back()] = Fr::from_u256(kv.first); repeated.pop_back(); }
            if (!repeated.empty()) throw std::runtime_error("create_proof: lookup permutation mismatch");
            for (size_t i = unusable_start; i < n; ++i) a.push_back(random_field<Fr>(rng));
            for (size_t i = unusable_start; i < n; ++i) pt.push_back(random_field<Fr>(rng));
            (void)random_field<Fr>(rng);  // Blind of the permuted input commitment
            (void)random_field<Fr>(rng);  // Blind of the permuted table commitment
            tr.write_point(params.commit_lagrange(a)); st.msms++;
            tr.write_point(params.commit_lagrange(pt)); st.msms++;
            lk_a[l] = std::move(a); lk_s[l] = std::move(pt);
        }
        trace2("lookup_permuted_input", lk_a);
        trace2("lookup_permuted_table", lk_s);
    }

    lap("lookup permuted");
    Fr beta = tr.squeeze_challenge(), gamma = tr.squeeze_challenge();

    // permutation grand products
    auto column_values = [&](const ColumnRef& c) -> const std::vector<Fr>& {
        return c.type == COL_ADVICE ? advice[c.index] : c.type == COL_FIXED ? pk.fixed_values[c.index] : instance_values;
    };
    const unsigned chunk = cs.chunk_len();
    std::vector<std::vector<Fr>> z_polys, z_cosets;
    {
        Fr deltaomega = Fr::one(), last_z = Fr::one();
        for (size_t c0 = 0; c0 < cs.perm_columns.size(); c0 += chunk) {
            size_t c1 = std::min(cs.perm_columns.size(), c0 + chunk);
            std::vector<Fr> mod(n, Fr::one());
            for (size_t c = c0; c < c1; ++c) {
                const auto& vals = column_values(cs.perm_columns[c]);
                for (size_t i = 0; i < n; ++i) mod[i] = mod[i] * (beta * pk.perm_values[c][i] + gamma + vals[i]);
            }
            batch_invert(mod.data(), n);
            for (size_t c = c0; c < c1; ++c) {
                const auto& vals = column_values(cs.perm_columns[c]);
                Fr dw = deltaomega;
                for (size_t i = 0; i < n; ++i) { mod[i] = mod[i] * (dw * beta + gamma + vals[i]); dw = dw * d.omega; }
                deltaomega = deltaomega * FrConst::delta();
            }
            std::vector<Fr> z(n);
            z[0] = last_z;
            for (size_t row = 1; row < n; ++row) z[row] = z[row - 1] * mod[row - 1];
            for (size_t i = n - bf; i < n; ++i) z[i] = random_field<Fr>(rng);
            last_z = z[n - (bf + 1)];
            (void)random_field<Fr>(rng);  // Blind
            trace("z", z);
            tr.write_point(params.commit_lagrange(z)); st.msms++;
            std::vector<Fr> zc = d.lagrange_to_coeff(z); st.ntts++;
            z_cosets.push_back(d.coeff_to_extended(zc)); st.ext_ntts++;
            z_polys.push_back(std::move(zc));
        }
    }

    lap("permutation z + commit + ntt");
    // lookup arguments, part 2 (commit_product): z[0] = 1, z[i+1] = z[i] (A_i + beta)(S_i + gamma) / ((A'_i + beta)(S'_i + gamma))
    std::vector<std::vector<Fr>> lk_z_polys(L), lk_a_polys(L), lk_s_polys(L), lk_z_cosets(L), lk_a_cosets(L), lk_s_cosets(L);
    for (size_t l = 0; l < L; ++l) {
        std::vector<Fr> den(n);
        for (size_t i = 0; i < n; ++i) den[i] = (beta + lk_a[l][i]) * (gamma + lk_s[l][i]);
        batch_invert(den.data(), n);
        for (size_t i = 0; i < n; ++i) den[i] = den[i] * (lk_in[l][i] + beta) * (lk_tab[l][i] + gamma);
        std::vector<Fr> z(n);
        z[0] = Fr::one();
        for (size_t row = 1; row < n - bf; ++row) z[row] = z[row - 1] * den[row - 1];
        for (size_t i = n - bf; i < n; ++i) z[i] = random_field<Fr>(rng);
        (void)random_field<Fr>(rng);  // Blind
        trace("lookup_z", z);
        tr.write_point(params.commit_lagrange(z)); st.msms++;
        lk_z_polys[l] = d.lagrange_to_coeff(z); st.ntts++;
        lk_a_polys[l] = d.lagrange_to_coeff(lk_a[l]); st.ntts++;
        lk_s_polys[l] = d.lagrange_to_coeff(lk_s[l]); st.ntts++;
        lk_z_cosets[l] = d.coeff_to_extended(lk_z_polys[l]); st.ext_ntts++;
        lk_a_cosets[l] = d.coeff_to_extended(lk_a_polys[l]); st.ext_ntts++;
        lk_s_cosets[l] = d.coeff_to_extended(lk_s_polys[l]); st.ext_ntts++;
    }

    lap("lookup z");
    // vanishing argument: random polynomial.  halo2 v0.3.0 vanishing/prover.rs fills it in chunks of n / num_threads
    // coefficients (num_threads = rayon::current_num_threads(); one more, shorter chunk when num_threads does not divide n),
    // each chunk from its own ChaCha20Rng whose 32-byte seed is drawn from the main rng, in chunk order [UPSTREAM-MEMORY,
    // SURVEY Appendix A / H3].  g_rayon_threads = 1 is the single-stream case (the `multicore` feature off).
    std::vector<Fr> random_poly(n);
    {
        const size_t T = g_rayon_threads ? g_rayon_threads : 1;
        const size_t chunk = std::max<size_t>(1, n / T);
        std::vector<uint8_t> chunk_seeds_tmp;
        for (size_t lo = 0; lo < n; lo += chunk) {
            uint8_t seed[32]; rng.fill_bytes(seed, 32);
            chunk_seeds_tmp.insert(chunk_seeds_tmp.end(), seed, seed + 32);
        }
        size_t ci = 0;
        for (size_t lo = 0; lo < n; lo += chunk, ++ci) {
            ChaCha20Rng crng(&chunk_seeds_tmp[32 * ci]);
            for (size_t i = lo; i < std::min(n, lo + chunk); ++i) random_poly[i] = random_field<Fr>(crng);
        }
        (void)random_field<Fr>(rng);  // Blind
        trace("random_poly", random_poly);
        tr.write_point(params.commit(random_poly)); st.msms++;
    }

    lap("random poly + commit");
    Fr y = tr.squeeze_challenge();

    std::vector<std::vector<Fr>> advice_polys;
    for (auto& col : advice) { advice_polys.push_back(d.lagrange_to_coeff(col)); st.ntts++; }

    trace2("advice_poly", advice_polys);
    trace2("z_poly", z_polys);
    trace2("z_coset", z_cosets);
    lap("advice intt");
    // evaluate_h on the extended coset
    std::vector<Fr> h(en, Fr::zero());
    {
        std::vector<std::vector<Fr>> advice_cosets, instance_cosets;
        for (auto& p : advice_polys) { advice_cosets.push_back(d.coeff_to_extended(p)); st.ext_ntts++; }
        instance_cosets.push_back(d.coeff_to_extended(instance_poly)); st.ext_ntts++;
        trace2("advice_coset", advice_cosets);
        trace2("instance_coset", instance_cosets);
        const long rot_scale = (long)1 << (d.extended_k - d.k);
        auto ridx = [&](size_t i, int r) { return (size_t)((((long)i + r * rot_scale) % (long)en + (long)en) % (long)en); };
        auto col_coset = [&](const ColumnRef& c) -> const std::vector<Fr>& {
            return c.type == COL_ADVICE ? advice_cosets[c.index] : c.type == COL_FIXED ? pk.fixed_cosets[c.index] : instance_cosets[0];
        };
        const size_t P = z_cosets.size();
        const int last_rot = cs.rotation_last();
        Fr one = Fr::one();
        parallel_chunks(en, params.threads, [&](size_t s, size_t e) {
            // current_delta = beta * zeta * ext_omega^i at the chunk start (zeta coset)
            for (size_t i = s; i < e; ++i) {
                Fr v = Fr::zero();
                for (auto& g : cs.gates) {
                    Fr t = eval_expr(g, cs.constants,
                        [&](uint32_t q) { return pk.fixed_cosets[cs.fixed_queries[q].column][ridx(i, cs.fixed_queries[q].rotation)]; },
                        [&](uint32_t q) { return advice_cosets[cs.advice_queries[q].column][ridx(i, cs.advice_queries[q].rotation)]; },
                        [&](uint32_t q) { return instance_cosets[0][ridx(i, cs.instance_queries[q].rotation)]; });
                    v = v * y + t;
                }
                if (P) {
                    size_t r_next = ridx(i, 1), r_last = ridx(i, last_rot);
                    v = v * y + (one - z_cosets[0][i]) * pk.l0[i];
                    v = v * y + (z_cosets[P - 1][i].square() - z_cosets[P - 1][i]) * pk.l_last[i];
                    for (size_t sidx = 1; sidx < P; ++sidx) v = v * y + (z_cosets[sidx][i] - z_cosets[sidx - 1][r_last]) * pk.l0[i];
                    Fr current_delta = beta * d.g_coset * d.extended_omega.pow_u64(i);
                    for (size_t sidx = 0; sidx < P; ++sidx) {
                        size_t c0 = sidx * chunk, c1 = std::min(cs.perm_columns.size(), c0 + chunk);
                        Fr left = z_cosets[sidx][r_next], right = z_cosets[sidx][i];
                        for (size_t c = c0; c < c1; ++c) {
                            const Fr& val = col_coset(cs.perm_columns[c])[i];
                            left = left * (val + beta * pk.perm_cosets[c][i] + gamma);
                            right = right * (val + current_delta + gamma);
                            current_delta = current_delta * FrConst::delta();
                        }
                        v = v * y + (left - right) * pk.l_active_row[i];
                    }
                }
                for (size_t l = 0; l < L; ++l) {
                    size_t r_next = ridx(i, 1), r_prev = ridx(i, -1);
                    auto compress = [&](const std::vector<Expr>& exprs) {
                        Fr acc = Fr::zero();
                        for (auto& e : exprs)
                            acc = acc * theta + eval_expr(e, cs.constants,
                                [&](uint32_t q) { return pk.fixed_cosets[cs.fixed_queries[q].column][ridx(i, cs.fixed_queries[q].rotation)]; },
                                [&](uint32_t q) { return advice_cosets[cs.advice_queries[q].column][ridx(i, cs.advice_queries[q].rotation)]; },
                                [&](uint32_t q) { return instance_cosets[0][ridx(i, cs.instance_queries[q].rotation)]; });
                        return acc;
                    };
                    const Fr& z = lk_z_cosets[l][i]; const Fr& a = lk_a_cosets[l][i]; const Fr& sp = lk_s_cosets[l][i];
                    Fr a_minus_s = a - sp;
                    v = v * y + (one - z) * pk.l0[i];
                    v = v * y + (z.square() - z) * pk.l_last[i];
                    v = v * y + (lk_z_cosets[l][r_next] * (a + beta) * (sp + gamma) - z * (compress(cs.lookups[l].inputs) + beta) * (compress(cs.lookups[l].tables) + gamma)) * pk.l_active_row[i];
                    v = v * y + a_minus_s * pk.l0[i];
                    v = v * y + a_minus_s * (a - lk_a_cosets[l][r_prev]) * pk.l_active_row[i];
                }
                h[i] = v;
            }
        });
    }

    lap("coset ntt + evaluate_h");
    // vanishing construct: divide by X^n - 1 on the coset, back to coefficients, split, commit pieces
    d.divide_by_vanishing_poly(h);
    trace("h_evals", h);
    std::vector<Fr> h_coeffs = d.extended_to_coeff(h); st.ext_ntts++;
    trace("h_coeffs", h_coeffs);
    const unsigned Q = cs.num_quotients();
    std::vector<std::vector<Fr>> h_pieces;
    for (unsigned i = 0; i < Q; ++i) h_pieces.emplace_back(h_coeffs.begin() + i * n, h_coeffs.begin() + (i + 1) * n);
    for (unsigned i = 0; i < Q; ++i) (void)random_field<Fr>(rng);  // h_blinds
    for (auto& p : h_pieces) { tr.write_point(params.commit(p)); st.msms++; }

    lap("h: intt + commit pieces");
    Fr x = tr.squeeze_challenge();
    Fr xn = x.pow_u64(n);

    std::vector<Fr> evals;
    for (auto& q : cs.advice_queries) evals.push_back(eval_polynomial(advice_polys[q.column].data(), n, d.rotate_omega(x, q.rotation)));
    for (auto& q : cs.fixed_queries) evals.push_back(eval_polynomial(pk.fixed_polys[q.column].data(), n, d.rotate_omega(x, q.rotation)));
    // vanishing evaluate: h(X) = sum_i xn^i * piece_i
    std::vector<Fr> h_poly(n, Fr::zero());
    for (unsigned i = Q; i-- > 0;) for (size_t t = 0; t < n; ++t) h_poly[t] = h_poly[t] * xn + h_pieces[i][t];
    evals.push_back(eval_polynomial(random_poly.data(), n, x));
    for (auto& p : pk.perm_polys) evals.push_back(eval_polynomial(p.data(), n, x));
    for (size_t s = 0; s < z_polys.size(); ++s) {
        evals.push_back(eval_polynomial(z_polys[s].data(), n, x));
        evals.push_back(eval_polynomial(z_polys[s].data(), n, d.rotate_omega(x, 1)));
        if (s + 1 < z_polys.size()) evals.push_back(eval_polynomial(z_polys[s].data(), n, d.rotate_omega(x, cs.rotation_last())));
    }
    for (size_t l = 0; l < L; ++l) {
        evals.push_back(eval_polynomial(lk_z_polys[l].data(), n, x));
        evals.push_back(eval_polynomial(lk_z_polys[l].data(), n, d.rotate_omega(x, 1)));
        evals.push_back(eval_polynomial(lk_a_polys[l].data(), n, x));
        evals.push_back(eval_polynomial(lk_a_polys[l].data(), n, d.rotate_omega(x, -1)));
        evals.push_back(eval_polynomial(lk_s_polys[l].data(), n, x));
    }
    trace("evals", evals);
    trace("h_poly", h_poly);
    for (auto& e : evals) tr.write_scalar(e);
    evals.push_back(eval_polynomial(h_poly.data(), n, x));  // "computed" quotient eval: not written

    lap("evaluations");
    // SHPLONK multiopen
    QueryPlan plan(cs);
    auto poly_of = [&](int id) -> const std::vector<Fr>& {
        if (id < plan.id_perm_z0) return advice_polys[id];
        if (id < plan.id_lk_z0) return z_polys[id - plan.id_perm_z0];
        if (id < plan.id_lk_a0) return lk_z_polys[id - plan.id_lk_z0];
        if (id < plan.id_lk_s0) return lk_a_polys[id - plan.id_lk_a0];
        if (id < plan.id_fixed0) return lk_s_polys[id - plan.id_lk_s0];
        if (id < plan.id_sigma0) return pk.fixed_polys[id - plan.id_fixed0];
        if (id < plan.id_h) return pk.perm_polys[id - plan.id_sigma0];
        return id == plan.id_h ? h_poly : random_poly;
    };
    Fr zeta = tr.squeeze_challenge(), nu = tr.squeeze_challenge();
    std::map<int, Fr> point;
    for (int r : plan.superset) point[r] = d.rotate_omega(x, r);
    std::vector<Fr> hx(n, Fr::zero());
    std::vector<std::vector<Fr>> set_combined;            // sum_j zeta^j p_ij(X)
    std::vector<std::vector<Fr>> set_r;                   // sum_j zeta^j r_ij(X)
    {
        Fr nu_pow = Fr::one();
        for (auto& s : plan.sets) {
            std::vector<Fr> pts; for (int r : s.rots) pts.push_back(point[r]);
            std::vector<Fr> comb(n, Fr::zero()), rcomb(pts.size(), Fr::zero());
            Fr zp = Fr::one();
            for (size_t j = 0; j < s.comms.size(); ++j) {
                const auto& p = poly_of(s.comms[j]);
                for (size_t t = 0; t < n; ++t) comb[t] += p[t] * zp;
                std::vector<Fr> ev; for (int e : s.evals[j]) ev.push_back(evals[e]);
                std::vector<Fr> rj = lagrange_interpolate(pts, ev);
                for (size_t t = 0; t < rj.size(); ++t) rcomb[t] += rj[t] * zp;
                zp = zp * zeta;
            }
            std::vector<Fr> num = comb;
            for (size_t t = 0; t < rcomb.size(); ++t) num[t] -= rcomb[t];
            for (auto& pt : pts) { num = kate_division(num, pt); }
            for (size_t t = 0; t < num.size(); ++t) hx[t] += num[t] * nu_pow;
            nu_pow = nu_pow * nu;
            set_combined.push_back(std::move(comb)); set_r.push_back(std::move(rcomb));
        }
    }
    trace2("set_combined", set_combined);
    trace("hx", hx);
    lap("shplonk h(X) build");
    tr.write_point(params.commit(hx)); st.msms++;
    Fr mu = tr.squeeze_challenge();
    {
        auto zeval = [&](const std::vector<int>& rots) { Fr a = Fr::one(); for (int r : rots) a = a * (mu - point[r]); return a; };
        Fr z_t = zeval(plan.superset);
        std::vector<Fr> lx(n, Fr::zero());
        Fr nu_pow = Fr::one(), zdiff0_inv = Fr::zero();
        for (size_t i = 0; i < plan.sets.size(); ++i) {
            Fr zd = zeval(plan.sets[i].diffs);
            if (i == 0) zdiff0_inv = zd.inv();
            Fr r_at_mu = eval_polynomial(set_r[i].data(), set_r[i].size(), mu);
            Fr sc = nu_pow * zd;
            for (size_t t = 0; t < n; ++t) lx[t] += set_combined[i][t] * sc;
            lx[0] -= r_at_mu * sc;
            nu_pow = nu_pow * nu;
        }
        for (size_t t = 0; t < n; ++t) lx[t] = (lx[t] - hx[t] * z_t) * zdiff0_inv;
        std::vector<Fr> wq = kate_division(lx, mu);
        wq.resize(n, Fr::zero());
        trace("lx", lx);
        trace("wq", wq);
        tr.write_point(params.commit(wq)); st.msms++;
    }
    lap("shplonk rest");
    if (stats) { stats->msms = st.msms; stats->ntts = st.ntts; stats->ext_ntts = st.ext_ntts; }
    if (tr.proof.size() != cs.proof_len()) throw std::runtime_error("create_proof: proof length mismatch");
    return tr.proof;
}

// ---------------------------------------------------------------------------------------------
// verify_proof — restatement of the generated Solidity verifier
// ---------------------------------------------------------------------------------------------
struct PairingInputs { G1Affine lhs, rhs; };

static inline bool verify_proof_to_pairing(const ParamsKZG& params, const VerifyingKey& vk, const uint8_t* proof, size_t proof_len,
                                           const std::vector<Fr>& instance, PairingInputs& out, std::string* why = nullptr) {
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    const ConstraintSystem& cs = vk.cs;
    const size_t n = cs.n();
    if (proof_len != cs.proof_len()) return fail("proof length");
    EvaluationDomain d(cs.degree(), cs.k);
    Transcript tr; tr.rd = proof; tr.rd_end = proof + proof_len;
    tr.common_scalar(vk.transcript_repr);
    for (auto& v : instance) tr.common_scalar(v);
    bool ok = true;
    const unsigned A = cs.num_advice, P = cs.num_perm_sets(), Q = cs.num_quotients();
    const size_t L = cs.num_lookups();
    std::vector<G1Affine> advice_c(A), z_c(P), h_c(Q), lk_a_c(L), lk_s_c(L), lk_z_c(L); G1Affine random_c, W, Wp;
    for (auto& p : advice_c) ok &= tr.read_point(p);
    Fr theta = tr.squeeze_challenge();
    for (size_t l = 0; l < L; ++l) { ok &= tr.read_point(lk_a_c[l]); ok &= tr.read_point(lk_s_c[l]); }
    Fr beta = tr.squeeze_challenge(), gamma = tr.squeeze_challenge();
    for (auto& p : z_c) ok &= tr.read_point(p);
    for (auto& p : lk_z_c) ok &= tr.read_point(p);
    ok &= tr.read_point(random_c);
    Fr y = tr.squeeze_challenge();
    for (auto& p : h_c) ok &= tr.read_point(p);
    Fr x = tr.squeeze_challenge();
    std::vector<Fr> evals(cs.num_evals());
    for (auto& e : evals) ok &= tr.read_scalar(e);
    Fr zeta = tr.squeeze_challenge(), nu = tr.squeeze_challenge();
    ok &= tr.read_point(W);
    Fr mu = tr.squeeze_challenge();
    ok &= tr.read_point(Wp);
    if (!ok) return fail("malformed proof element");

    // Lagrange evaluations (Halo2Verifier.sol:392-470)
    Fr xn = x.pow_u64(n);
    Fr xn_m1 = xn - Fr::one();
    const int rl = cs.rotation_last();
    long num_l = (long)std::max<size_t>(instance.size(), 1) - rl;  // j = rl .. num_instances-1 (at least l_0)
    std::vector<Fr> dens(num_l + 1), wpow(num_l);
    {
        Fr w = d.rotate_omega(Fr::one(), rl);
        for (long j = 0; j < num_l; ++j) { wpow[j] = w; dens[j] = x - w; w = w * d.omega; }
        dens[num_l] = xn_m1;
        for (auto& v : dens) if (v.is_zero()) return fail("batch inversion of zero");
        batch_invert(dens.data(), dens.size());
    }
    Fr common = xn_m1 * d.barycentric_weight;  // (x^n - 1)/n
    std::vector<Fr> lag(num_l);
    for (long j = 0; j < num_l; ++j) lag[j] = common * dens[j] * wpow[j];
    Fr l_last = lag[0], l_blind = Fr::zero();
    for (long j = 1; j < -rl; ++j) l_blind += lag[j];
    Fr l_0 = lag[-rl];
    Fr instance_eval = Fr::zero();
    for (size_t i = 0; i < instance.size(); ++i) instance_eval += lag[-rl + i] * instance[i];
    Fr xn_m1_inv = dens[num_l];

    // quotient numerator (codegen.rs:237-254, evaluator.rs:45-131)
    const size_t e_fix = cs.advice_queries.size(), e_rand = e_fix + cs.fixed_queries.size(), e_sigma = e_rand + 1, e_z = e_sigma + cs.perm_columns.size();
    Fr numer = Fr::zero();
    for (auto& g : cs.gates) {
        Fr t = eval_expr(g, cs.constants, [&](uint32_t q) { return evals[e_fix + q]; }, [&](uint32_t q) { return evals[q]; },
                         [&](uint32_t) { return instance_eval; });
        numer = numer * y + t;
    }
    auto col_eval = [&](const ColumnRef& c) -> Fr {
        if (c.type == COL_INSTANCE) return instance_eval;
        const auto& qs = c.type == COL_ADVICE ? cs.advice_queries : cs.fixed_queries;
        for (size_t i = 0; i < qs.size(); ++i) if (qs[i].column == c.index && qs[i].rotation == 0) return evals[(c.type == COL_ADVICE ? 0 : e_fix) + i];
        throw std::runtime_error("permutation column is not queried at rotation 0");
    };
    if (P) {
        auto zx = [&](size_t s) { return evals[e_z + 3 * s]; };
        auto zwx = [&](size_t s) { return evals[e_z + 3 * s + 1]; };
        auto zlast = [&](size_t s) { return evals[e_z + 3 * s + 2]; };
        numer = numer * y + (l_0 - l_0 * zx(0));
        numer = numer * y + l_last * (zx(P - 1).square() - zx(P - 1));
        for (size_t s = 0; s + 1 < P; ++s) numer = numer * y + l_0 * (zx(s + 1) - zlast(s));
        Fr cur = beta * x;
        const unsigned chunk = cs.chunk_len();
        for (size_t s = 0; s < P; ++s) {
            size_t c0 = s * chunk, c1 = std::min(cs.perm_columns.size(), c0 + chunk);
            Fr lhs = zwx(s), rhs = zx(s);
            for (size_t c = c0; c < c1; ++c) lhs = lhs * (col_eval(cs.perm_columns[c]) + beta * evals[e_sigma + c] + gamma);
            for (size_t c = c0; c < c1; ++c) { rhs = rhs * (col_eval(cs.perm_columns[c]) + cur + gamma); cur = cur * FrConst::delta(); }
            Fr lsr = lhs - rhs;
            numer = numer * y + (lsr - lsr * (l_last + l_blind));
        }
    }
    // lookup terms (codegen/evaluator.rs:126-223)
    {
        const size_t e_lk = e_z + (P ? 3 * P - 1 : 0);
        Fr l_active = Fr::one() - (l_blind + l_last);
        auto compress = [&](const std::vector<Expr>& exprs) {
            Fr acc = Fr::zero();
            for (auto& e : exprs)
                acc = acc * theta + eval_expr(e, cs.constants, [&](uint32_t q) { return evals[e_fix + q]; }, [&](uint32_t q) { return evals[q]; },
                                              [&](uint32_t) { return instance_eval; });
            return acc;
        };
        for (size_t l = 0; l < L; ++l) {
            const Fr &z = evals[e_lk + 5 * l], &z_next = evals[e_lk + 5 * l + 1], &p_in = evals[e_lk + 5 * l + 2], &p_in_prev = evals[e_lk + 5 * l + 3],
                     &p_tab = evals[e_lk + 5 * l + 4];
            numer = numer * y + (l_0 - l_0 * z);
            numer = numer * y + l_last * (z * z - z);
            Fr lhs = z_next * ((p_in + beta) * (p_tab + gamma));
            Fr rhs = z * ((compress(cs.lookups[l].inputs) + beta) * (compress(cs.lookups[l].tables) + gamma));
            numer = numer * y + l_active * (lhs - rhs);
            numer = numer * y + l_0 * (p_in - p_tab);
            numer = numer * y + l_active * ((p_in - p_tab) * (p_in - p_in_prev));
        }
    }
    Fr quotient_eval = numer * xn_m1_inv;

    // quotient commitment (Halo2Verifier.sol:494-512)
    G1 hq = G1::from_affine(h_c[Q - 1]);
    for (unsigned i = Q - 1; i-- > 0;) hq = hq.mul(xn).add_mixed(h_c[i]);
    G1Affine quotient_c = hq.to_affine();

    // SHPLONK (pcs/bdfg21.rs)
    QueryPlan plan(cs);
    std::vector<Fr> all_evals = evals; all_evals.push_back(quotient_eval);
    auto comm_of = [&](int id) -> G1Affine {
        if (id < plan.id_perm_z0) return advice_c[id];
        if (id < plan.id_lk_z0) return z_c[id - plan.id_perm_z0];
        if (id < plan.id_lk_a0) return lk_z_c[id - plan.id_lk_z0];
        if (id < plan.id_lk_s0) return lk_a_c[id - plan.id_lk_a0];
        if (id < plan.id_fixed0) return lk_s_c[id - plan.id_lk_s0];
        if (id < plan.id_sigma0) return vk.fixed_commitments[id - plan.id_fixed0];
        if (id < plan.id_h) return vk.perm_commitments[id - plan.id_sigma0];
        return id == plan.id_h ? quotient_c : random_c;
    };
    std::map<int, Fr> point, mu_minus;
    for (int r : plan.superset) { point[r] = d.rotate_omega(x, r); mu_minus[r] = mu - point[r]; }
    Fr vanishing_0 = Fr::one();
    for (int r : plan.sets[0].rots) vanishing_0 = vanishing_0 * mu_minus[r];
    size_t m = plan.sets.size();
    std::vector<Fr> diffs(m);
    for (size_t i = 0; i < m; ++i) { Fr dd = Fr::one(); for (int r : plan.sets[i].diffs) dd = dd * mu_minus[r]; diffs[i] = dd; }
    std::vector<std::vector<Fr>> coeffs(m);
    for (size_t i = 0; i < m; ++i)
        for (int ra : plan.sets[i].rots) {
            Fr c = Fr::one();
            for (int rb : plan.sets[i].rots) if (rb != ra) c = c * (point[ra] - point[rb]);
            coeffs[i].push_back(c * mu_minus[ra]);
        }
    // first batch inversion: diff_0 and all coeffs
    {
        std::vector<Fr> inv{diffs[0]};
        for (auto& cv : coeffs) for (auto& c : cv) inv.push_back(c);
        for (auto& v : inv) if (v.is_zero()) return fail("batch inversion of zero");
        batch_invert(inv.data(), inv.size());
        Fr diff0_inv = inv[0];
        size_t t = 1;
        for (auto& cv : coeffs) for (auto& c : cv) c = inv[t++];
        diffs[0] = diff0_inv;  // unused below for set 0
        for (size_t i = 1; i < m; ++i) diffs[i] = diffs[i] * diff0_inv;
    }
    std::vector<Fr> r_evals(m), sums(m);
    for (size_t i = 0; i < m; ++i) {
        const auto& s = plan.sets[i];
        Fr r = Fr::zero();
        for (size_t j = s.comms.size(); j-- > 0;) {
            Fr inner = Fr::zero();
            for (size_t a = 0; a < s.rots.size(); ++a) inner += coeffs[i][a] * all_evals[s.evals[j][a]];
            r = r * zeta + inner;
        }
        if (i) r = r * diffs[i];
        r_evals[i] = r;
        Fr sm = Fr::zero(); for (auto& c : coeffs[i]) sm += c; sums[i] = sm;
    }
    for (auto& v : sums) if (v.is_zero()) return fail("batch inversion of zero");
    batch_invert(sums.data(), m);
    Fr r_eval = Fr::zero();
    for (size_t i = m; i-- > 0;) r_eval = r_eval * nu + sums[i] * r_evals[i];
    // pairing lhs
    G1 acc = G1::identity();
    Fr nu_pow = Fr::one();
    for (size_t i = 0; i < m; ++i) {
        const auto& s = plan.sets[i];
        G1 t = G1::from_affine(comm_of(s.comms.back()));
        for (size_t j = s.comms.size() - 1; j-- > 0;) t = t.mul(zeta).add_mixed(comm_of(s.comms[j]));
        if (i == 0) acc = t;
        else { nu_pow = nu_pow * nu; acc = acc.add(t.mul(nu_pow * diffs[i])); }
    }
    acc = acc.add(G1::from_affine(params.g[0]).mul(-r_eval));
    acc = acc.add(G1::from_affine(W).mul(-vanishing_0));
    acc = acc.add(G1::from_affine(Wp).mul(mu));
    out.lhs = acc.to_affine();
    out.rhs = Wp;
    return true;
}

static inline bool verify_proof(const ParamsKZG& params, const VerifyingKey& vk, const uint8_t* proof, size_t len,
                                const std::vector<Fr>& instance, std::string* why = nullptr) {
    PairingInputs pi;
    if (!verify_proof_to_pairing(params, vk, proof, len, instance, pi, why)) return false;
    // e(lhs, g2) * e(-rhs, s_g2) == 1   (Halo2Verifier.sol:204-219, 552-559 pair (rhs, -s*G2))
    bool ok = pairing_product_is_one(pi.lhs, params.g2, pi.rhs.neg(), params.s_g2);
    if (!ok && why) *why = "pairing check failed";
    return ok;
}

// Batch verification with a random linear combination of the pairing inputs: every proof is replayed to
// its (lhs_i, rhs_i); accept iff e(sum r_i lhs_i, g2) * e(-sum r_i rhs_i, s_g2) == 1 for random r_i.  Sound up to 1/r.
// Returns the number of proofs that failed BEFORE the pairing (malformed, wrong length); `all_ok` is the batched check.
static inline size_t verify_proofs_batch(const ParamsKZG& params, const VerifyingKey& vk, const uint8_t* proofs, size_t proof_len, size_t m,
                                         const Fr* instances, size_t n_inst, unsigned threads, u64 seed, bool& all_ok) {
    std::vector<PairingInputs> pi(m);
    std::vector<uint8_t> good(m, 0);
    parallel_chunks(m, threads, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            std::vector<Fr> inst(instances + i * n_inst, instances + (i + 1) * n_inst);
            good[i] = verify_proof_to_pairing(params, vk, proofs + i * proof_len, proof_len, inst, pi[i]) ? 1 : 0;
        }
    });
    size_t bad = 0;
    for (auto g : good) bad += g ? 0 : 1;
    SmallRng rng(seed);
    std::vector<G1> lp(threads ? threads : 1, G1::identity()), rp(lp);
    std::vector<Fr> r(m);
    for (auto& x : r) x = random_field<Fr>(rng);
    size_t nchunks = lp.size(), chunk = (m + nchunks - 1) / nchunks;
    parallel_chunks(nchunks, threads, [&](size_t a, size_t b) {
        for (size_t c = a; c < b; ++c)
            for (size_t i = c * chunk; i < std::min(m, (c + 1) * chunk); ++i) {
                if (!good[i]) continue;
                lp[c] = lp[c].add(G1::from_affine(pi[i].lhs).mul(r[i]));
                rp[c] = rp[c].add(G1::from_affine(pi[i].rhs).mul(r[i]));
            }
    });
    G1 L = G1::identity(), R = G1::identity();
    for (auto& x : lp) L = L.add(x);
    for (auto& x : rp) R = R.add(x);
    all_ok = bad == 0 && pairing_product_is_one(L.to_affine(), params.g2, R.to_affine().neg(), params.s_g2);
    return bad;
}

}  // namespace oracle
