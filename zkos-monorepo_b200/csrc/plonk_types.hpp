// Constraint-system description consumed by the GPU prover: what halo2's `ConstraintSystem` +
// keygen assembly carry for one circuit (columns, queries, gate polynomials, permutation columns,
// fixed assignment, copy constraints).  Derived quantities follow halo2 v0.3.0 plonk/circuit.rs
// as mirrored by the in-repo verifier generator
// (/root/reference/crates/halo2-verifier/src/lib/codegen/util.rs:42-132: chunk_len = degree-2,
// num_quotients = degree-1, rotation_last = -(blinding_factors+1), num_evals).
// Blob layout (little-endian): see zkgpu/circuits.py `Circuit._serialize`.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "common.cuh"
#include "fp.cuh"

namespace zk {

enum : uint32_t { OP_CONST = 0, OP_FIXED = 1, OP_ADVICE = 2, OP_INSTANCE = 3, OP_NEG = 4, OP_ADD = 5, OP_MUL = 6, OP_SCALE = 7 };
enum : uint32_t { COL_ADVICE = 0, COL_FIXED = 1, COL_INSTANCE = 2 };
static const uint32_t CS_ONLY_MAGIC = 0x5a4b4354u;
struct ExprIns { uint32_t op, arg; };
struct ColRef { uint32_t type, index; };
struct QueryRef { uint32_t column; int32_t rotation; };
struct CopyRef { uint32_t lcol, lrow, rcol, rrow; };

struct CsDesc {
    uint32_t k = 0, num_fixed = 0, num_advice = 0, num_instance = 0;
    std::vector<QueryRef> advice_queries, fixed_queries, instance_queries;
    std::vector<fr_t> constants;
    std::vector<std::vector<ExprIns>> gates;
    std::vector<ColRef> perm_columns;
    struct Lookup { std::vector<std::vector<ExprIns>> inputs, tables; };   // cs.lookups(): input / table expressions
    std::vector<Lookup> lookups;
    std::vector<fr_t> fixed;  // num_fixed * n, column-major
    std::vector<CopyRef> copies;
    std::vector<uint8_t> blob;
    // constraint-system-only blobs (magic CS_ONLY_MAGIC, written by the Rust exporter of INTEGRATION.md next to a pk.bin): no fixed
    // assignment and no copy constraints — those come from the pk.bin — but the selector count of the verifying key section and
    // `vk.transcript_repr()`
    bool cs_only = false;
    uint32_t num_selectors = 0;
    fr_t transcript_repr;

    size_t n() const { return (size_t)1 << k; }
    static unsigned expr_degree(const std::vector<ExprIns>& e) {
        std::vector<unsigned> st;
        for (auto& i : e) switch (i.op) {
            case OP_CONST: st.push_back(0); break;
            case OP_FIXED: case OP_ADVICE: case OP_INSTANCE: st.push_back(1); break;
            case OP_ADD: { unsigned b = st.back(); st.pop_back(); st.back() = std::max(st.back(), b); break; }
            case OP_MUL: { unsigned b = st.back(); st.pop_back(); st.back() += b; break; }
            default: break;
        }
        return st.empty() ? 0 : st.back();
    }
    static unsigned expr_stack_depth(const std::vector<ExprIns>& e) {
        unsigned d = 0, mx = 0;
        for (auto& i : e) {
            if (i.op <= OP_INSTANCE) { ++d; mx = std::max(mx, d); }
            else if (i.op == OP_ADD || i.op == OP_MUL) --d;
        }
        return mx;
    }
    // ConstraintSystem::degree(): starts from the permutation argument's required degree (3, whether or not any column is
    // copy-enabled), lookups max(4, 2 + input_degree + table_degree), gates their own
    unsigned degree() const {
        unsigned d = 3;
        for (auto& l : lookups) {
            unsigned di = 1, dt = 1;
            for (auto& e : l.inputs) di = std::max(di, expr_degree(e));
            for (auto& e : l.tables) dt = std::max(dt, expr_degree(e));
            d = std::max(d, std::max(4u, 2 + di + dt));
        }
        for (auto& g : gates) d = std::max(d, expr_degree(g));
        return d;
    }
    size_t num_lookups() const { return lookups.size(); }
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
    size_t unusable_start() const { return n() - (blinding_factors() + 1); }
    unsigned extended_k() const {
        unsigned ek = k;
        while (((size_t)1 << ek) < n() * (degree() - 1)) ++ek;
        return ek;
    }
    size_t num_evals() const {
        return advice_queries.size() + fixed_queries.size() + 1 + perm_columns.size() + (num_perm_sets() ? 3 * num_perm_sets() - 1 : 0) + 5 * num_lookups();
    }
    // codegen/util.rs:175-186: advice | 2L permuted | P + L grand products | random | Q quotient pieces | evals | W, W'
    size_t proof_len() const { return 64 * ((size_t)num_advice + 3 * num_lookups() + num_perm_sets() + 1 + num_quotients()) + 32 * num_evals() + 128; }

    static CsDesc parse(const uint8_t* data, size_t len) {
        CsDesc c;
        const uint8_t* p = data; const uint8_t* end = data + len;
        auto u32 = [&]() { ZK_REQUIRE(p + 4 <= end, "circuit blob: truncated"); uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
        auto fr = [&]() { ZK_REQUIRE(p + 32 <= end, "circuit blob: truncated"); fr_t v; memcpy(v.l, p, 32); p += 32; return v; };
        const uint32_t magic = u32();
        ZK_REQUIRE(magic == 0x5a4b4353u || magic == CS_ONLY_MAGIC, "circuit blob: bad magic");
        c.cs_only = magic == CS_ONLY_MAGIC;
        // every count below is bounded by what the rest of the blob can hold (8 bytes per query / instruction / column)
        auto bounded = [&](uint32_t m, size_t each) { ZK_REQUIRE((size_t)(end - p) / each >= m, "circuit blob: count exceeds the blob"); return m; };
        c.k = u32(); c.num_fixed = u32(); c.num_advice = u32(); c.num_instance = u32();
        ZK_REQUIRE(c.k >= 3 && c.k <= 20, "circuit: k out of range");
        auto rq = [&](std::vector<QueryRef>& v) { uint32_t m = bounded(u32(), 8); v.resize(m); for (auto& q : v) { q.column = u32(); q.rotation = (int32_t)u32(); } };
        rq(c.advice_queries); rq(c.fixed_queries); rq(c.instance_queries);
        uint32_t nc = bounded(u32(), 32); c.constants.resize(nc); for (auto& f : c.constants) f = fr();
        uint32_t ng = bounded(u32(), 4); c.gates.resize(ng);
        for (auto& g : c.gates) { uint32_t m = bounded(u32(), 8); g.resize(m); for (auto& i : g) { i.op = u32(); i.arg = u32(); } }
        uint32_t np = bounded(u32(), 8); c.perm_columns.resize(np); for (auto& pc : c.perm_columns) { pc.type = u32(); pc.index = u32(); }
        auto rexpr = [&](std::vector<ExprIns>& g) { uint32_t m = u32(); ZK_REQUIRE(m <= 4096, "expression too long"); g.resize(m); for (auto& i : g) { i.op = u32(); i.arg = u32(); } };
        uint32_t nl = u32(); ZK_REQUIRE(nl <= 64, "too many lookups"); c.lookups.resize(nl);
        for (auto& l : c.lookups) {
            uint32_t ni = u32(); ZK_REQUIRE(ni >= 1 && ni <= 16, "lookup: bad input expression count"); l.inputs.resize(ni); for (auto& e : l.inputs) rexpr(e);
            uint32_t nt = u32(); ZK_REQUIRE(nt == ni, "lookup: input/table expression counts differ"); l.tables.resize(nt); for (auto& e : l.tables) rexpr(e);
        }
        ZK_REQUIRE(c.num_instance == 1, "exactly one instance column is supported (as in Shielder's circuits)");
        size_t n = c.n();
        if (c.cs_only) {
            c.num_selectors = u32();
            c.transcript_repr = fr();
        } else {
            ZK_REQUIRE((size_t)(end - p) >= (size_t)c.num_fixed * n * 32, "circuit blob: truncated fixed columns");
            c.fixed.resize((size_t)c.num_fixed * n);
            memcpy(c.fixed.data(), p, c.fixed.size() * 32); p += c.fixed.size() * 32;
            uint32_t ncp = bounded(u32(), 16); c.copies.resize(ncp);
            for (auto& cp : c.copies) { cp.lcol = u32(); cp.lrow = u32(); cp.rcol = u32(); cp.rrow = u32(); }
        }
        ZK_REQUIRE(p == end, "circuit blob: trailing bytes");
        // validation
        for (auto& q : c.advice_queries) ZK_REQUIRE(q.column < c.num_advice, "advice query out of range");
        for (auto& q : c.fixed_queries) ZK_REQUIRE(q.column < c.num_fixed, "fixed query out of range");
        for (auto& q : c.instance_queries) ZK_REQUIRE(q.column < c.num_instance, "instance query out of range");
        std::vector<const std::vector<ExprIns>*> all_exprs;
        for (auto& g : c.gates) all_exprs.push_back(&g);
        for (auto& l : c.lookups) { for (auto& e : l.inputs) all_exprs.push_back(&e); for (auto& e : l.tables) all_exprs.push_back(&e); }
        for (auto* gp : all_exprs) {
            const std::vector<ExprIns>& g = *gp;
            int depth = 0;
            for (auto& i : g) {
                switch (i.op) {
                    case OP_CONST: ZK_REQUIRE(i.arg < c.constants.size(), "gate: constant out of range"); ++depth; break;
                    case OP_FIXED: ZK_REQUIRE(i.arg < c.fixed_queries.size(), "gate: fixed query out of range"); ++depth; break;
                    case OP_ADVICE: ZK_REQUIRE(i.arg < c.advice_queries.size(), "gate: advice query out of range"); ++depth; break;
                    case OP_INSTANCE: ZK_REQUIRE(i.arg < c.instance_queries.size(), "gate: instance query out of range"); ++depth; break;
                    case OP_NEG: ZK_REQUIRE(depth >= 1, "gate: stack underflow"); break;
                    case OP_SCALE: ZK_REQUIRE(depth >= 1 && i.arg < c.constants.size(), "gate: bad scale"); break;
                    case OP_ADD: case OP_MUL: ZK_REQUIRE(depth >= 2, "gate: stack underflow"); --depth; break;
                    default: ZK_REQUIRE(false, "gate: unknown op");
                }
            }
            ZK_REQUIRE(depth == 1, "gate: expression must leave one value");
            ZK_REQUIRE(expr_stack_depth(g) <= 8, "gate: expression stack deeper than 8");
        }
        for (auto& pc : c.perm_columns) {
            ZK_REQUIRE(pc.type <= COL_INSTANCE, "permutation: bad column type");
            ZK_REQUIRE(pc.index < (pc.type == COL_ADVICE ? c.num_advice : pc.type == COL_FIXED ? c.num_fixed : c.num_instance), "permutation: column out of range");
        }
        for (auto& cp : c.copies)
            ZK_REQUIRE(cp.lcol < c.perm_columns.size() && cp.rcol < c.perm_columns.size() && cp.lrow < n && cp.rrow < n, "copy constraint out of range");
        c.blob.assign(data, data + len);
        return c;
    }
};

// permutation::keygen::Assembly — cycle merging (halo2 permutation/keygen.rs)
struct PermAssembly {
    size_t ncols, n;
    std::vector<uint32_t> map_col, map_row, aux_col, aux_row, sizes;
    PermAssembly(size_t ncols_, size_t n_) : ncols(ncols_), n(n_), map_col(ncols_ * n_), map_row(ncols_ * n_), aux_col(ncols_ * n_), aux_row(ncols_ * n_), sizes(ncols_ * n_, 1) {
        for (size_t c = 0; c < ncols; ++c)
            for (size_t r = 0; r < n; ++r) { map_col[c * n + r] = aux_col[c * n + r] = (uint32_t)c; map_row[c * n + r] = aux_row[c * n + r] = (uint32_t)r; }
    }
    void copy(uint32_t lc, uint32_t lr, uint32_t rc, uint32_t rr) {
        size_t li = (size_t)lc * n + lr, ri = (size_t)rc * n + rr;
        uint32_t Lc = aux_col[li], Lr = aux_row[li], Rc = aux_col[ri], Rr = aux_row[ri];
        if (Lc == Rc && Lr == Rr) return;
        if (sizes[(size_t)Lc * n + Lr] < sizes[(size_t)Rc * n + Rr]) { std::swap(Lc, Rc); std::swap(Lr, Rr); }
        sizes[(size_t)Lc * n + Lr] += sizes[(size_t)Rc * n + Rr];
        uint32_t ic = Rc, ir = Rr;
        do {
            size_t ii = (size_t)ic * n + ir;
            aux_col[ii] = Lc; aux_row[ii] = Lr;
            uint32_t nc = map_col[ii], nr = map_row[ii];
            ic = nc; ir = nr;
        } while (!(ic == Rc && ir == Rr));
        std::swap(map_col[li], map_col[ri]);
        std::swap(map_row[li], map_row[ri]);
    }
};

}  // namespace zk
