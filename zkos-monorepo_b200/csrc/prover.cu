// Batched halo2 `create_proof` on the device — the C++ host driver of the proving pipeline.
//
// Replaces, for a batch of independent proofs of ONE circuit, what Shielder reaches through
//   shielder_circuits::generate_proof(&params, &pk, circuit, &public_input, rng)
//     -> halo2_proofs::plonk::create_proof::<KZGCommitmentScheme<Bn256>, ProverSHPLONK, ChallengeEvm, _,
//        Keccak256Transcript, _>
// (/root/reference/crates/shielder_bindings/src/circuits/mod.rs:103-111,
//  /root/reference/tee/crates/shielder-prover-tee/src/circuits/mod.rs:70-78).  Step order, RNG draw
// order and transcript writes follow halo2 v0.3.0 plonk/prover.rs, permutation/prover.rs,
// vanishing/prover.rs and poly/kzg/multiopen/shplonk/prover.rs; the proof
// layout is the one the in-repo verifier generator reads
// (/root/reference/crates/halo2-verifier/src/lib/codegen/util.rs:226-245,
//  /root/reference/crates/halo2-verifier/templates/Halo2Verifier.sol:247-307).
//
// Everything that scales with n runs on the GPU (MSM, NTT, grand products, quotient evaluation,
// Horner evaluations, SHPLONK polynomial algebra); the host keeps only the Fiat-Shamir transcript
// (Keccak), the seeded RNG stream and O(#queries) scalar bookkeeping per proof.  A sub-batch of B
// proofs advances in lock step so every kernel launch covers B proofs; the six transcript round trips
// (theta/beta/gamma, y, x, zeta/nu, mu) are the only host synchronisation points.
// Witness synthesis (the circuit's own Rust code) is outside the path: the caller passes assigned
// advice columns, as `create_proof` has them after `synthesize`.
#include "api_util.hpp"
#include "prover_kernels.cuh"
#include "plonk_types.hpp"
#include "host_util.hpp"
#include "pk_file.hpp"
#include <map>
#include <atomic>
#include <set>
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <thread>
#include <exception>
#include <deque>
#include <condition_variable>

namespace zk {

typedef void (*trace_fn)(const char* name, const void* data, size_t bytes);
static trace_fn g_trace = nullptr;
// wall-clock seconds spent in each step of prove_sub_batch (every step ends in a stream sync)
static double g_step_s[8] = {0};
static std::mutex g_step_mu;   // two pipeline workers accumulate here
struct StepTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(int step) {
        auto t1 = std::chrono::steady_clock::now();
        { std::lock_guard<std::mutex> lk(g_step_mu); g_step_s[step] += std::chrono::duration<double>(t1 - t0).count(); }
        t0 = t1;
    }
};

// ---------------------------------------------------------------------------------------------
// SHPLONK query plan: queries in the order of codegen/pcs.rs:60-104, rotation sets as
// pcs/bdfg21.rs:443-494 builds them (first-seen order of distinct rotation sets).
// commitment ids: [0,A) advice | P perm z | L lookup z | L permuted inputs | L permuted tables | F fixed | S sigma | h | random
// ---------------------------------------------------------------------------------------------
struct RotSet {
    std::vector<int> rots, diffs, comms;
    std::vector<std::vector<int>> evals;  // [comm][rot] -> index into the eval list
};
struct QueryPlan {
    std::vector<int> superset;
    std::vector<RotSet> sets;
    int id_z0 = 0, id_lz0 = 0, id_la0 = 0, id_ls0 = 0, id_fixed0 = 0, id_sigma0 = 0, id_h = 0, id_random = 0;
    void build(const CsDesc& cs) {
        const int A = cs.num_advice, P = cs.num_perm_sets(), F = cs.num_fixed, S = (int)cs.perm_columns.size(), L = (int)cs.num_lookups();
        id_z0 = A; id_lz0 = A + P; id_la0 = id_lz0 + L; id_ls0 = id_la0 + L; id_fixed0 = id_ls0 + L; id_sigma0 = id_fixed0 + F;
        id_h = id_sigma0 + S; id_random = id_h + 1;
        const int e_fix = (int)cs.advice_queries.size(), e_rand = e_fix + (int)cs.fixed_queries.size();
        const int e_sigma = e_rand + 1, e_z = e_sigma + S, e_lk = e_z + (P ? 3 * P - 1 : 0), e_h = (int)cs.num_evals();
        struct Q { int comm, rot, eval; };
        std::vector<Q> qs;
        for (size_t i = 0; i < cs.advice_queries.size(); ++i) qs.push_back({(int)cs.advice_queries[i].column, cs.advice_queries[i].rotation, (int)i});
        for (int s = 0; s < P; ++s) { qs.push_back({id_z0 + s, 0, e_z + 3 * s}); qs.push_back({id_z0 + s, 1, e_z + 3 * s + 1}); }
        for (int s = P - 2; s >= 0; --s) qs.push_back({id_z0 + s, cs.rotation_last(), e_z + 3 * s + 2});
        for (int l = 0; l < L; ++l) {  // codegen/pcs.rs:80-92
            const int e = e_lk + 5 * l;
            qs.push_back({id_lz0 + l, 0, e}); qs.push_back({id_la0 + l, 0, e + 2}); qs.push_back({id_ls0 + l, 0, e + 4});
            qs.push_back({id_la0 + l, -1, e + 3}); qs.push_back({id_lz0 + l, 1, e + 1});
        }
        for (size_t i = 0; i < cs.fixed_queries.size(); ++i) qs.push_back({id_fixed0 + (int)cs.fixed_queries[i].column, cs.fixed_queries[i].rotation, e_fix + (int)i});
        for (int s = 0; s < S; ++s) qs.push_back({id_sigma0 + s, 0, e_sigma + s});
        qs.push_back({id_h, 0, e_h});
        qs.push_back({id_random, 0, e_rand});
        std::set<int> sup;
        std::vector<std::pair<int, std::map<int, int>>> per_comm;  // first-seen order of commitments
        for (auto& q : qs) {
            sup.insert(q.rot);
            size_t j = 0;
            while (j < per_comm.size() && per_comm[j].first != q.comm) ++j;
            if (j == per_comm.size()) per_comm.push_back({q.comm, {}});
            per_comm[j].second[q.rot] = q.eval;
        }
        superset.assign(sup.begin(), sup.end());
        for (auto& pc : per_comm) {
            std::vector<int> rots, evs;
            for (auto& re : pc.second) { rots.push_back(re.first); evs.push_back(re.second); }
            size_t j = 0;
            while (j < sets.size() && sets[j].rots != rots) ++j;
            if (j == sets.size()) {
                RotSet s; s.rots = rots;
                for (int r : superset) if (!pc.second.count(r)) s.diffs.push_back(r);
                sets.push_back(s);
            }
            sets[j].comms.push_back(pc.first);
            sets[j].evals.push_back(evs);
        }
        for (auto& s : sets) ZK_REQUIRE(s.rots.size() <= 4, "shplonk: more than 4 rotations in one set");
    }
};

// ---------------------------------------------------------------------------------------------
// proving key (device resident) + reusable per-batch workspace
// ---------------------------------------------------------------------------------------------
struct HostPinned {
    void* p = nullptr; size_t n = 0;
    ~HostPinned() { if (p) cudaFreeHost(p); }
    void ensure(size_t bytes) {
        if (bytes <= n) return;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        ZK_CUDA(cudaMallocHost(&p, bytes));
        n = bytes;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct ProverWs {
    size_t B = 0;
    cudaStream_t stream = nullptr;   // each worker owns a stream, an MSM workspace and the buffers below
    MsmWorkspace msm;
    DevBuf<fr_t> carries;
    // host advice of the worker's NEXT sub-batch is uploaded on copy_stream into adv_next while the current one computes
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev = nullptr;
    cudaEvent_t ev_pts = nullptr;    // "commitments of step 2 copied back" marker, reused by every sub-batch
    DevBuf<fr_t> adv_next;
    const fr_t* prefetched_src = nullptr;
    // single-proof regime: the advice / instance transforms run on a side stream WHILE the advice commitments are computed
    cudaStream_t side = nullptr;
    cudaEvent_t ev_blind = nullptr, ev_side = nullptr, ev_z = nullptr, ev_zside = nullptr;
    DevBuf<fr_t> adv_coef, inst_coef, z_coef, nd, scratch2;
    ~ProverWs() {
        if (side) cudaStreamDestroy(side);
        if (ev_blind) cudaEventDestroy(ev_blind);
        if (ev_side) cudaEventDestroy(ev_side);
        if (ev_z) cudaEventDestroy(ev_z);
        if (ev_zside) cudaEventDestroy(ev_zside);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (copy_ev) cudaEventDestroy(copy_ev);
        if (ev_pts) cudaEventDestroy(ev_pts);
    }
    DevBuf<fr_t> adv, inst, z, randp, adv_ext, z_ext, h, hpoly, comb, hx, lx, tmp1, tmp2, scratch, evals, low;
    DevBuf<fr_t> lk_in, lk_tab, lk_a, lk_s, lk_z, lk_ext, sort_a, sort_t;   // lookups: [B][L][n] (lk_ext: [B][L][3][Qc*n])
    DevBuf<uint64_t> raw_la, raw_ls, raw_lz;
    DevBuf<int> d_error;
    DevBuf<uint64_t> raw_adv, raw_z;
    DevBuf<uint8_t> seeds;
    DevBuf<Challenges> ch;
    DevBuf<g1_affine_t> aff;
    DevBuf<g1_xyzz_t> xyzz;
    DevBuf<EvalJob> eval_jobs;
    DevBuf<LinTerm> terms, terms2;
    DevBuf<uint32_t> job_off, job_off2;
    DevBuf<fr_t*> outs, outs2;
    DevBuf<DivJob> div_jobs;
    HostPinned h_aff, h_evals, h_stage;
};

static const unsigned BATCH_WORKERS = 4;   // pipeline workers of a zkgpu_prove_batch call (at most)
static const unsigned COALESCE_WORKERS = 3;   // dispatcher threads (per device) behind zkgpu_prove: one uploads while two compute
struct PkEntry {   // one replica per selected device
    Context* C = nullptr;
    std::mutex batch_mu;   // one zkgpu_prove_batch at a time per replica (its workers own ws[0..2])
    CsDesc cs;
    uint64_t srs_handle = 0;
    unsigned k = 0, ek = 0, A = 0, F = 0, S = 0, P = 0, Q = 0, bf = 0, chunk = 0, L = 0;
    // The quotient is evaluated on Qc = Q cosets g_c * H (g_c = zeta * ext_omega^c, H the size-n subgroup) instead of on
    // all 2^(ek-k) cosets of halo2's extended domain: h has Q*n coefficients, so Q cosets determine it.  cn = Qc * n rows.
    unsigned Qc = 0;
    size_t cn = 0;
    DevBuf<fr_t> coset_pows, coset_pows_inv, vinv;   // g_c^m [Qc][n], g_c^-m [Qc][n], inverse Vandermonde [Qc][Qc]
    DevBuf<uint32_t> lk_prog, lk_expr_off, lk_off;
    size_t n = 0, en = 0, ustart = 0, num_evals = 0, proof_len = 0;
    int rot_last = 0;
    fr_t omega, omega_inv, ext_omega, n_inv, zeta, digest;
    DevBuf<fr_t> fixed_vals, fixed_polys, fixed_ext, sigma_vals, sigma_polys, sigma_ext, l0, l_last, l_active, t_inv, delta_pows, constants;
    DevBuf<ColSrc> cols;
    DevBuf<uint32_t> prog, gate_off;
    DevBuf<int32_t> adv_q, fix_q, inst_q;
    const fr_t* omega_tw = nullptr;
    const fr_t* ext_tw = nullptr;
    std::vector<g1_affine_t> fixed_commitments, perm_commitments;
    QueryPlan plan;
    ProverWs ws[BATCH_WORKERS + COALESCE_WORKERS];   // ws[0] also serves keygen; ws[3..] belong to the zkgpu_prove dispatchers
    mutable size_t cached_batch = 0;
};

// One blocking single-proof request (zkgpu_prove): what one tokio task of the reference's prover server holds while it waits
// for `generate_proof` (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:157-195).
struct ProveReq {
    const fr_t* advice; const fr_t* instance; size_t num_pi;
    int rng_mode; uint8_t* rng_data;
    uint8_t* proof_out;
    int32_t status = 0;
    int rc = 0; std::string err;
    bool done = false;
};
struct PkShared;
// Request coalescer: concurrent zkgpu_prove callers enqueue and sleep; per device, COALESCE_WORKERS dispatcher threads turn
// whatever is waiting into one lock-step sub-batch each.  No caller ever queues on a mutex around the GPU.
struct Coalescer {
    std::mutex mu;
    std::condition_variable cv_work, cv_done, cv_slot;
    std::deque<ProveReq*> q;
    // Pinned staging slots (one request's advice columns each).  A caller copies its advice into a slot BEFORE it queues — the
    // hundred callers do that in parallel, on their own threads — so a dispatcher's upload of a sub-batch is one burst of DMA from
    // pinned memory instead of B staged copies from pageable buffers on the dispatcher's thread.
    std::vector<void*> free_slots;
    size_t slots_allocated = 0;
    static const size_t MAX_SLOTS = 256;
    unsigned idle = 0;
    bool stop = false;
    std::vector<std::thread> threads;
    uint64_t batches = 0, requests = 0, max_batch_seen = 0;   // statistics (zkgpu_prove_stats)
};
struct PkShared {
    std::vector<std::unique_ptr<PkEntry>> dev;   // replica per selected device, in Runtime::devs order
    Coalescer co;
    std::once_flag co_once;
    ~PkShared();
};

static std::map<uint64_t, std::shared_ptr<PkShared>> g_pks;   // guarded by rt().tab_mu
// rayon::current_num_threads() of the host being replaced: halo2's vanishing prover fills the random polynomial in chunks of
// n / num_threads coefficients, one ChaCha20 stream (seeded from the proof's main rng, in chunk order) per chunk, so the proof bytes
// depend on it.  1 = one stream (the `multicore` feature off); set with zkgpu_set_rayon_threads.
static std::atomic<unsigned> g_rayon_threads{1};
static size_t vanishing_chunk(size_t n) { unsigned t = g_rayon_threads.load(); return std::max<size_t>(1, n / (t ? t : 1)); }
static uint64_t g_next_pk = 1;
void prover_release_all() {
    std::map<uint64_t, std::shared_ptr<PkShared>> drop;
    { std::unique_lock<std::shared_mutex> tl(rt().tab_mu); drop.swap(g_pks); }
    drop.clear();   // joins the dispatcher threads, frees device memory
}
static std::shared_ptr<PkShared> find_pk(uint64_t handle) {
    std::shared_lock<std::shared_mutex> tl(rt().tab_mu);
    auto it = g_pks.find(handle);
    ZK_REQUIRE(it != g_pks.end(), "unknown proving key handle");
    return it->second;
}

static fr_t host_rotate(const PkEntry& pk, const fr_t& x, int rot) {
    if (rot >= 0) return x * fr_pow_u64(pk.omega, (uint64_t)rot);
    return x * fr_pow_u64(pk.omega_inv, (uint64_t)(-(long)rot));
}
static void host_batch_invert(std::vector<fr_t>& v) {
    std::vector<fr_t> pre(v.size());
    fr_t acc = fe_one<FrTag>();
    for (size_t i = 0; i < v.size(); ++i) { pre[i] = acc; acc = acc * v[i]; }
    acc = fe_inv(acc);
    for (size_t i = v.size(); i-- > 0;) { fr_t t = acc * pre[i]; acc = acc * v[i]; v[i] = t; }
}

// ---- NTT helpers ----------------------------------------------------------------------------
static const size_t SCRATCH_ELEMS = (size_t)1 << 25;  // 1 GiB of two-pass NTT scratch

// lagrange_to_coeff on `count` contiguous polynomials of 2^k values, in place
// (out: where the coefficients go, default in place; scratch: which staging buffer, default the worker's)
static void intt_n(PkEntry& pk, ProverWs& W, fr_t* p, size_t count, cudaStream_t st, fr_t* out = nullptr, DevBuf<fr_t>* scratch = nullptr) {
    const bool two = ntt_scratch_elems(pk.k, 1) != 0;   // two-pass transforms stage through a scratch buffer, chunk by chunk
    DevBuf<fr_t>& S = scratch ? *scratch : W.scratch;
    size_t per = two ? std::max<size_t>(1, SCRATCH_ELEMS >> pk.k) : count;
    if (two) S.ensure(std::min(count, per) << pk.k);
    for (size_t off = 0; off < count; off += per) {
        NttJob J;
        J.in = p + (off << pk.k); J.out = (out ? out : p) + (off << pk.k); J.scratch = S.p; J.batch = std::min(per, count - off); J.log_n = pk.k;
        J.omega = pk.omega_inv; J.has_scale = 1; J.scale = pk.n_inv;
        ntt_run(J, st);
    }
}
// coefficients -> values on the Qc quotient cosets: groups x cols polynomials at in[(g*cols + c)*n] -> out[g*out_group_stride + c*cn],
// each output column coset-major [Qc][n].  Every coset is one size-n NTT of a[m] * g_c^m (table pk.coset_pows).
static void coset_ext(PkEntry& pk, ProverWs& W, const fr_t* in, fr_t* out, size_t groups, size_t cols, size_t out_group_stride, cudaStream_t st,
                      DevBuf<fr_t>* scratch = nullptr) {
    const bool two = ntt_scratch_elems(pk.k, 1) != 0;
    DevBuf<fr_t>& S = scratch ? *scratch : W.scratch;
    const size_t per_group = cols * pk.Qc;
    size_t per = two ? std::max<size_t>(1, (SCRATCH_ELEMS >> pk.k) / per_group) : groups;
    if (two) S.ensure((std::min(groups, per) * per_group) << pk.k);
    for (size_t off = 0; off < groups; off += per) {
        size_t g = std::min(per, groups - off);
        NttJob J;
        J.in = in + off * cols * pk.n; J.out = out + off * out_group_stride; J.scratch = S.p;
        J.batch = g * per_group; J.log_n = pk.k; J.omega = pk.omega;
        J.in_broadcast = 1; J.in_inner = pk.Qc; J.in_outer_stride = pk.n;
        J.out_stride = pk.n; J.out_inner = per_group; J.out_outer_stride = out_group_stride;
        J.pre_table = pk.coset_pows.p; J.pre_count = pk.Qc;
        ntt_run(J, st);
    }
}
// A few commitments (the rounds of a single proof): the latency path of msm.cu over the SRS's narrow-window tables; every item names its
// basis (0 = g, 1 = g_lagrange), so the commitments of one Fiat-Shamir round share a launch group whatever their basis.
struct CommitItem { int basis; const fr_t* scalars; };
static bool commit_lat_ok(Context& C, const PkEntry& pk, size_t M) {
    SrsEntry& S = C.get_srs(pk.srs_handle);
    return S.lat_tables.p != nullptr && M >= 1 && M <= ZK_LAT_MAX_M;
}
static void commit_lat(Context& C, PkEntry& pk, ProverWs& W, const std::vector<CommitItem>& items, g1_affine_t* d_out, cudaStream_t st) {
    SrsEntry& S = C.get_srs(pk.srs_handle);
    MsmPlan plan = S.lat_plan;
    plan.n = pk.n; plan.tstride = S.n;
    const fr_t* sc[ZK_LAT_MAX_M];
    uint32_t mask = 0;
    for (size_t m = 0; m < items.size(); ++m) { sc[m] = items[m].scalars; if (items[m].basis) mask |= 1u << m; }
    if (S.direct_tables.p) msm_direct_run(sc, mask, S.direct_stride, S.direct_tables.p, pk.n, S.n, items.size(), d_out, W.msm, st);
    else msm_lat_run(plan, sc, mask, S.lat_stride(), S.lat_tables.p, items.size(), d_out, W.msm, st);
}

// M commitments; MSM m reads scalars at (m / inner) * outer_stride + (m % inner) * n.  Affine out.
static void commit(Context& C, PkEntry& pk, ProverWs& W, int basis, const fr_t* d_scalars, size_t M, size_t inner, size_t outer_stride,
                   g1_affine_t* d_out, cudaStream_t st) {
    SrsEntry& S = C.get_srs(pk.srs_handle);
    if (commit_lat_ok(C, pk, M)) {
        std::vector<CommitItem> items(M);
        for (size_t m = 0; m < M; ++m)
            items[m] = {basis, inner ? d_scalars + (m / inner) * outer_stride + (m % inner) * pk.n : d_scalars + m * pk.n};
        commit_lat(C, pk, W, items, d_out, st);
        return;
    }
    MsmPlan plan = S.plan;
    plan.n = pk.n; plan.tstride = S.n;
    if (inner == 0) { inner = 1; outer_stride = pk.n; }
    plan.inner = inner; plan.outer_stride = outer_stride;
    size_t chunk = std::max<size_t>(inner, (1024 / inner) * inner);
    W.xyzz.ensure(std::min(M, chunk));
    for (size_t off = 0; off < M; off += chunk) {
        size_t cnt = std::min(chunk, M - off);
        msm_run(plan, d_scalars + (off / inner) * outer_stride, S.table[basis].p, cnt, W.xyzz.p, W.msm, st);
        g1_normalize(W.xyzz.p, d_out + off, cnt, st);
    }
}

// Upload of a sub-batch's advice columns (0.8 GB for 128 withdraw proofs).  The host->device copy engine is a FIFO shared
// by every stream: a small pageable upload of either pipeline worker (challenges, RNG draws, job lists) issued while this
// transfer is in flight blocks its host thread until the transfer ends.  When the caller's buffer is pinned (and therefore
// device-mapped under UVA) a few CTAs pull it over PCIe instead, and the copy engine stays free for the small uploads.
static void upload_advice(void* dst, const void* src, size_t bytes, cudaStream_t st, bool background) {
    static const int mode = [] { const char* e = getenv("ZKGPU_H2D_PULL"); return e ? atoi(e) : 1; }();
    static const unsigned ctas = [] { const char* e = getenv("ZKGPU_H2D_PULL_CTAS"); int v = e ? atoi(e) : 0; return (unsigned)(v > 0 ? v : 32); }();
    // the copy engine is about twice as fast as SM reads over PCIe: a transfer nothing can hide (a worker's first sub-batch) uses it
    if (mode && background && bytes % 16 == 0) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            launch_pull_from_host(dst, at.devicePointer, bytes, ctas, st);
            return;
        }
        cudaGetLastError();   // unregistered host memory reports an error on some drivers: fall through to the copy engine
    }
    ZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
}
// Small host -> device uploads (rng draws, challenges, job lists: a few per Fiat-Shamir step).  From pageable memory every one of
// them synchronises the stream and is staged by the driver before the call returns (measured: ~50 us each, 0.3 ms of a single proof's
// first step).  A sub-batch therefore stages them through one pinned arena of its worker: a host memcpy, then a truly asynchronous
// copy.  The arena is rewound at the start of a sub-batch (the previous one ended with a stream synchronisation); a request that
// does not fit falls back to the pageable copy.
struct StageArena {
    uint8_t* base = nullptr; size_t cap = 0, off = 0;
    void* put(const void* src, size_t bytes) {
        const size_t at = (off + 63) & ~(size_t)63;
        if (!base || at + bytes > cap) return nullptr;
        memcpy(base + at, src, bytes);
        off = at + bytes;
        return base + at;
    }
};
static thread_local StageArena* tl_arena = nullptr;
static void h2d_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!bytes) return;
    void* staged = tl_arena ? tl_arena->put(src, bytes) : nullptr;
    ZK_CUDA(cudaMemcpyAsync(dst, staged ? staged : src, bytes, cudaMemcpyHostToDevice, st));
}
template <class T>
static void h2d(T* dst, const std::vector<T>& src, cudaStream_t st) {
    h2d_bytes(dst, src.data(), src.size() * sizeof(T), st);
}
template <class T>
static void upload(DevBuf<T>& buf, const std::vector<T>& src, cudaStream_t st) {
    buf.ensure(std::max<size_t>(1, src.size()));
    h2d(buf.p, src, st);
}

static void trace_dev(const char* name, const fr_t* d, size_t count, size_t reps, size_t stride, cudaStream_t st) {
    if (!g_trace) return;
    std::vector<fr_t> h(count);
    for (size_t r = 0; r < reps; ++r) {
        ZK_CUDA(cudaMemcpyAsync(h.data(), d + r * stride, count * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        g_trace(name, h.data(), count * sizeof(fr_t));
    }
}

// ---------------------------------------------------------------------------------------------
// keygen_vk + keygen_pk
// ---------------------------------------------------------------------------------------------
// `file` == nullptr: keygen_vk + keygen_pk from the circuit blob's fixed assignment and copy constraints.
// `file` != nullptr: ProvingKey::read — the blob carries the constraint system only and every assignment-derived part (fixed / sigma
// values, coefficient forms, extended cosets, l_0 / l_last / l_active_row, the verifying key's commitments) is taken from the pk.bin
// instead of being recomputed: no commitments, no transforms, only a gather of the quotient cosets out of the file's extended domain.
static std::unique_ptr<PkEntry> keygen(Context& C, uint64_t srs_handle, const uint8_t* blob, size_t len, const uint8_t* pk_bin = nullptr, size_t pk_bin_len = 0) {
    std::unique_ptr<PkEntry> pkp(new PkEntry);
    PkEntry& pk = *pkp;
    pk.C = &C;
    pk.cs = CsDesc::parse(blob, len);
    const CsDesc& cs = pk.cs;
    ZK_REQUIRE(cs.cs_only == (pk_bin != nullptr), pk_bin ? "pk_load: expected a constraint-system-only blob (the assignment comes from the pk.bin)"
                                                          : "pk_create: the blob has no fixed assignment; use zkgpu_pk_load with the pk.bin");
    SrsEntry& S = C.get_srs(srs_handle);
    ZK_REQUIRE(S.k == cs.k, "keygen: params.k != circuit k (downsize the params first)");
    pk.srs_handle = srs_handle;
    pk.k = cs.k; pk.n = cs.n(); pk.ek = cs.extended_k(); pk.en = (size_t)1 << pk.ek;
    pk.Qc = cs.num_quotients(); pk.cn = (size_t)pk.Qc * pk.n;
    pk.A = cs.num_advice; pk.F = cs.num_fixed; pk.S = (unsigned)cs.perm_columns.size(); pk.P = cs.num_perm_sets(); pk.Q = cs.num_quotients();
    pk.L = (unsigned)cs.num_lookups();
    pk.bf = cs.blinding_factors(); pk.chunk = cs.chunk_len(); pk.ustart = cs.unusable_start(); pk.rot_last = cs.rotation_last();
    pk.num_evals = cs.num_evals(); pk.proof_len = cs.proof_len();
    ZK_REQUIRE(pk.ek <= 24, "keygen: extended domain too large");
    ZK_REQUIRE(pk.bf + 1 < pk.n, "keygen: not enough rows");
    pk.omega = fr_omega(pk.k); pk.omega_inv = fr_omega_inv(pk.k);
    pk.ext_omega = fr_omega(pk.ek);
    pk.n_inv = fr_pow2_inv(pk.k);
    pk.zeta = fr_from_limbs(fr_consts::ZETA);
    pk.plan.build(cs);
    cudaStream_t st = C.stream;
    const size_t n = pk.n, en = pk.cn;   // rows per coset-major column
    pk.omega_tw = ntt_twiddles(pk.k, pk.omega, st);
    pk.ext_tw = ntt_twiddles(pk.ek, pk.ext_omega, st);
    // quotient cosets: power tables of g_c = zeta * ext_omega^c and the inverse Vandermonde matrix in G_c = g_c^n
    {
        const unsigned Qc = pk.Qc;
        ZK_REQUIRE(Qc >= 1 && Qc <= 16 && ((size_t)Qc << pk.k) <= pk.en, "keygen: unsupported number of quotient pieces");
        pk.coset_pows.alloc((size_t)Qc * n); pk.coset_pows_inv.alloc((size_t)Qc * n);
        std::vector<fr_t> G(Qc);
        fr_t g = pk.zeta;
        for (unsigned c = 0; c < Qc; ++c) {
            fr_power_table(pk.coset_pows.p + (size_t)c * n, pk.k, g, st);
            fr_power_table(pk.coset_pows_inv.p + (size_t)c * n, pk.k, fe_inv(g), st);
            G[c] = fr_pow_u64(g, n);
            g = g * pk.ext_omega;
        }
        // Gauss-Jordan on [V | I], V[c][j] = G_c^j
        const fr_t one = fe_one<FrTag>();
        std::vector<std::vector<fr_t>> M(Qc, std::vector<fr_t>(2 * Qc, fr_t::zero()));
        for (unsigned c = 0; c < Qc; ++c) {
            fr_t p = one;
            for (unsigned j = 0; j < Qc; ++j) { M[c][j] = p; p = p * G[c]; }
            M[c][Qc + c] = one;
        }
        for (unsigned col = 0; col < Qc; ++col) {
            unsigned piv = col;
            while (piv < Qc && M[piv][col].is_zero()) ++piv;
            ZK_REQUIRE(piv < Qc, "keygen: singular coset Vandermonde matrix");
            std::swap(M[piv], M[col]);
            fr_t inv = fe_inv(M[col][col]);
            for (auto& v : M[col]) v = v * inv;
            for (unsigned r = 0; r < Qc; ++r) {
                if (r == col || M[r][col].is_zero()) continue;
                fr_t f = M[r][col];
                for (unsigned j = 0; j < 2 * Qc; ++j) M[r][j] = M[r][j] - f * M[col][j];
            }
        }
        std::vector<fr_t> vinv((size_t)Qc * Qc);
        for (unsigned j = 0; j < Qc; ++j) for (unsigned c = 0; c < Qc; ++c) vinv[j * Qc + c] = M[j][Qc + c];
        upload(pk.vinv, vinv, st);
        ZK_CUDA(cudaStreamSynchronize(st));
    }

    auto commit_cols = [&](const fr_t* d_vals, size_t count, std::vector<g1_affine_t>& out) {
        out.resize(count);
        if (!count) return;
        pk.ws[0].aff.ensure(count);
        commit(C, pk, pk.ws[0], 1, d_vals, count, 0, 0, pk.ws[0].aff.p, st);
        ZK_CUDA(cudaMemcpyAsync(out.data(), pk.ws[0].aff.p, count * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
    };
    auto polys_and_cosets = [&](const DevBuf<fr_t>& vals, DevBuf<fr_t>& polys, DevBuf<fr_t>& ext, size_t count) {
        polys.alloc(std::max<size_t>(1, count * n)); ext.alloc(std::max<size_t>(1, count * en));
        if (!count) return;
        ZK_CUDA(cudaMemcpyAsync(polys.p, vals.p, count * n * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        intt_n(pk, pk.ws[0], polys.p, count, st);
        coset_ext(pk, pk.ws[0], polys.p, ext.p, 1, count, count * en, st);
    };
    if (pk_bin) {
        PkFile f = PkFile::parse(pk_bin, pk_bin_len, pk.F, pk.S, cs.num_selectors, pk.ek);
        ZK_REQUIRE(f.k == pk.k, "pk_load: pk.bin was generated for another k");
        pk.fixed_commitments = f.fixed_commitments; pk.perm_commitments = f.perm_commitments;
        const unsigned log_e = pk.ek - pk.k;   // the extended domain is the union of 2^log_e cosets g_c H; coset c = entries c + 2^log_e j
        DevBuf<fr_t> stage(pk.en);
        auto column = [&](const PkFile::Poly& q, fr_t* dst) { ZK_CUDA(cudaMemcpyAsync(dst, q.p, n * sizeof(fr_t), cudaMemcpyHostToDevice, st)); };
        auto cosets = [&](const PkFile::Poly& q, fr_t* dst) {
            ZK_CUDA(cudaMemcpyAsync(stage.p, q.p, pk.en * sizeof(fr_t), cudaMemcpyHostToDevice, st));
            launch_gather_cosets(stage.p, dst, pk.k, log_e, pk.Qc, st);
        };
        auto group = [&](const std::vector<PkFile::Poly>& vals, const std::vector<PkFile::Poly>& polys, const std::vector<PkFile::Poly>& ext,
                         DevBuf<fr_t>& d_vals, DevBuf<fr_t>& d_polys, DevBuf<fr_t>& d_ext) {
            const size_t count = vals.size();
            d_vals.alloc(std::max<size_t>(1, count * n)); d_polys.alloc(std::max<size_t>(1, count * n)); d_ext.alloc(std::max<size_t>(1, count * en));
            for (size_t c = 0; c < count; ++c) { column(vals[c], d_vals.p + c * n); column(polys[c], d_polys.p + c * n); cosets(ext[c], d_ext.p + c * en); }
        };
        group(f.fixed_values, f.fixed_polys, f.fixed_cosets, pk.fixed_vals, pk.fixed_polys, pk.fixed_ext);
        group(f.perm_values, f.perm_polys, f.perm_cosets, pk.sigma_vals, pk.sigma_polys, pk.sigma_ext);
        pk.l0.alloc(en); pk.l_last.alloc(en); pk.l_active.alloc(en);
        cosets(f.l0, pk.l0.p); cosets(f.l_last, pk.l_last.p); cosets(f.l_active_row, pk.l_active.p);
        std::vector<fr_t> dp(std::max<unsigned>(1, pk.S));
        fr_t d = fe_one<FrTag>(), delta = fr_from_limbs(fr_consts::DELTA);
        for (unsigned c = 0; c < pk.S; ++c) { dp[c] = d; d = d * delta; }
        upload(pk.delta_pows, dp, st);
        std::vector<ColSrc> cols(std::max<unsigned>(1, pk.S));
        for (unsigned c = 0; c < pk.S; ++c) { cols[c].type = cs.perm_columns[c].type; cols[c].index = cs.perm_columns[c].index; }
        upload(pk.cols, cols, st);
        ZK_CUDA(cudaStreamSynchronize(st));
    } else {
        // fixed columns
        pk.fixed_vals.alloc(std::max<size_t>(1, pk.F * n));
        h2d(pk.fixed_vals.p, cs.fixed, st);
        commit_cols(pk.fixed_vals.p, pk.F, pk.fixed_commitments);
        polys_and_cosets(pk.fixed_vals, pk.fixed_polys, pk.fixed_ext, pk.F);
        // permutation: sigma columns hold delta^col * omega^row of the mapped cell
        {
            PermAssembly as(pk.S, n);
            for (auto& cp : cs.copies) as.copy(cp.lcol, cp.lrow, cp.rcol, cp.rrow);
            std::vector<fr_t> dp(std::max<unsigned>(1, pk.S));
            fr_t d = fe_one<FrTag>(), delta = fr_from_limbs(fr_consts::DELTA);
            for (unsigned c = 0; c < pk.S; ++c) { dp[c] = d; d = d * delta; }
            upload(pk.delta_pows, dp, st);
            DevBuf<uint32_t> d_mc(std::max<size_t>(1, as.map_col.size())), d_mr(std::max<size_t>(1, as.map_row.size()));
            h2d(d_mc.p, as.map_col, st); h2d(d_mr.p, as.map_row, st);
            pk.sigma_vals.alloc(std::max<size_t>(1, pk.S * n));
            launch_sigma_values(d_mc.p, d_mr.p, pk.delta_pows.p, pk.omega_tw, pk.sigma_vals.p, pk.S, pk.k, st);
            ZK_CUDA(cudaStreamSynchronize(st));
            commit_cols(pk.sigma_vals.p, pk.S, pk.perm_commitments);
            polys_and_cosets(pk.sigma_vals, pk.sigma_polys, pk.sigma_ext, pk.S);
            std::vector<ColSrc> cols(std::max<unsigned>(1, pk.S));
            for (unsigned c = 0; c < pk.S; ++c) { cols[c].type = cs.perm_columns[c].type; cols[c].index = cs.perm_columns[c].index; }
            upload(pk.cols, cols, st);
        }
        // l_0, l_blind, l_last on the quotient cosets; l_active_row = 1 - l_last - l_blind
        {
            std::vector<fr_t> lag(3 * n, fr_t::zero());
            fr_t one = fe_one<FrTag>();
            lag[0] = one;
            for (size_t i = n - pk.bf; i < n; ++i) lag[n + i] = one;
            lag[2 * n + (n - pk.bf - 1)] = one;
            DevBuf<fr_t> vals(3 * n), polys, ext;
            h2d(vals.p, lag, st);
            polys_and_cosets(vals, polys, ext, 3);
            pk.l0.alloc(en); pk.l_last.alloc(en); pk.l_active.alloc(en);
            ZK_CUDA(cudaMemcpyAsync(pk.l0.p, ext.p, en * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
            ZK_CUDA(cudaMemcpyAsync(pk.l_last.p, ext.p + 2 * en, en * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
            launch_one_minus_sum(ext.p + 2 * en, ext.p + en, pk.l_active.p, en, st);
            ZK_CUDA(cudaStreamSynchronize(st));
        }
    }
    // t_evaluations (inverted): ((zeta * ext_omega^i)^n - 1)^-1, i < 2^(ek-k)
    {
        size_t tcount = (size_t)1 << (pk.ek - pk.k);
        std::vector<fr_t> t(tcount);
        fr_t cur = fr_pow_u64(pk.zeta, n), step = fr_pow_u64(pk.ext_omega, n), one = fe_one<FrTag>();
        for (size_t i = 0; i < tcount; ++i) { t[i] = cur - one; cur = cur * step; }
        host_batch_invert(t);
        upload(pk.t_inv, t, st);
    }
    // gate programs
    {
        std::vector<uint32_t> prog, off{0};
        for (auto& g : cs.gates) {
            for (auto& i : g) { prog.push_back(i.op); prog.push_back(i.arg); }
            off.push_back((uint32_t)(prog.size() / 2));
        }
        if (prog.empty()) prog.assign(2, 0);
        upload(pk.prog, prog, st); upload(pk.gate_off, off, st);
        std::vector<fr_t> consts = cs.constants;
        if (consts.empty()) consts.push_back(fr_t::zero());
        upload(pk.constants, consts, st);
        auto qv = [](const std::vector<QueryRef>& q) {
            std::vector<int32_t> v;
            for (auto& e : q) { v.push_back((int32_t)e.column); v.push_back(e.rotation); }
            if (v.empty()) v.assign(2, 0);
            return v;
        };
        upload(pk.adv_q, qv(cs.advice_queries), st); upload(pk.fix_q, qv(cs.fixed_queries), st); upload(pk.inst_q, qv(cs.instance_queries), st);
        // lookup expression programs: lookup l owns expressions [lk_off[l], lk_off[l+1]) (inputs, then tables)
        std::vector<uint32_t> lprog, eoff{0}, loff{0};
        for (auto& l : cs.lookups) {
            for (auto* group : {&l.inputs, &l.tables})
                for (auto& e : *group) {
                    for (auto& i : e) { lprog.push_back(i.op); lprog.push_back(i.arg); }
                    eoff.push_back((uint32_t)(lprog.size() / 2));
                }
            loff.push_back((uint32_t)(eoff.size() - 1));
        }
        if (lprog.empty()) lprog.assign(2, 0);
        upload(pk.lk_prog, lprog, st); upload(pk.lk_expr_off, eoff, st); upload(pk.lk_off, loff, st);
        ZK_CUDA(cudaStreamSynchronize(st));
    }
    // vk digest.  Upstream `transcript_repr` is a Blake2b hash of the verifying key's Rust Debug output: with a pk.bin it is the value
    // the exporter wrote into the blob; for keys generated here it is an opaque constant, keccak(blob ‖ fixed commitments ‖ sigma
    // commitments) mod r (see oracle/plonk.hpp header).
    if (pk_bin) pk.digest = cs.transcript_repr;
    else {
        std::vector<uint8_t> in(cs.blob);
        auto add = [&](const g1_affine_t& p) { uint8_t w[64]; fe_to_be_bytes(p.x, w); fe_to_be_bytes(p.y, w + 32); in.insert(in.end(), w, w + 64); };
        for (auto& p : pk.fixed_commitments) add(p);
        for (auto& p : pk.perm_commitments) add(p);
        uint8_t h[32]; keccak256(in.data(), in.size(), h);
        pk.digest = fr_from_be_bytes_reduce(h);
    }
    return pkp;
}

// ---------------------------------------------------------------------------------------------
// create_proof for a sub-batch of B proofs
// ---------------------------------------------------------------------------------------------
struct ProofState {
    Transcript tr;
    fr_t theta, beta, gamma, y, x, xn, zeta, nu, mu;
    std::vector<fr_t> evals;                 // num_evals + 1 (quotient eval last)
    std::vector<fr_t> points;                // x * omega^r for r in superset
    std::vector<std::vector<fr_t>> rcomb;    // per set: low-degree remainder coefficients
    explicit ProofState(uint8_t* out) : tr(out) {}
};

static LookupProgs lookup_progs(const PkEntry& pk) {
    LookupProgs lp;
    lp.prog = pk.lk_prog.p; lp.expr_off = pk.lk_expr_off.p; lp.lk_off = pk.lk_off.p; lp.constants = pk.constants.p;
    lp.adv_q = pk.adv_q.p; lp.fix_q = pk.fix_q.p; lp.inst_q = pk.inst_q.p; lp.L = pk.L;
    return lp;
}

// Sub-batch size of a proving key.  Computed once per key (cudaMemGetInfo and the other allocator entry points take driver-wide
// locks: on a box shared with other CUDA processes a call can stall for milliseconds, which is the single-proof latency budget).
static size_t default_batch(const PkEntry& pk) {
    if (const char* e = getenv("ZKGPU_PROVER_BATCH")) { long v = atol(e); if (v > 0) return (size_t)v; }
    if (pk.cached_batch) return pk.cached_batch;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t per = ((size_t)(pk.A + 1 + pk.P + 1 + 3 * pk.L) * pk.cn + (size_t)(pk.A + 2 * pk.P + 7 * pk.L + 8 + 3 * pk.plan.sets.size()) * pk.n) * sizeof(fr_t);
    size_t B = (size_t)(free_b * 0.2) / std::max<size_t>(per, 1);
    pk.cached_batch = std::max<size_t>(1, std::min<size_t>(B, 128));
    return pk.cached_batch;
}

static void ensure_ws(PkEntry& pk, ProverWs& W, size_t B) {
    const size_t n = pk.n, en = pk.cn, ns = pk.plan.sets.size();
    W.B = std::max(W.B, B);
    W.adv.ensure(B * pk.A * n); W.inst.ensure(B * n); W.z.ensure(std::max<size_t>(1, B * pk.P * n)); W.randp.ensure(B * n);
    // adv_ext doubles as scratch for the grand-product numerators / denominators (2 x [B][P + L][n]) before it is filled
    W.adv_ext.ensure(std::max(B * (pk.A + 1) * en, 2 * B * (pk.P + pk.L) * n)); W.z_ext.ensure(std::max<size_t>(1, B * pk.P * en)); W.h.ensure(B * en);
    W.hpoly.ensure(B * n); W.comb.ensure(B * ns * n); W.hx.ensure(B * n); W.lx.ensure(B * n);
    W.tmp1.ensure(B * ns * n); W.tmp2.ensure(B * ns * n);
    W.evals.ensure(B * (pk.num_evals + 1)); W.low.ensure(B * (ns + 1) * 4);
    if (pk.L) {
        const size_t BL = B * pk.L;
        W.lk_in.ensure(BL * n); W.lk_tab.ensure(BL * n); W.lk_a.ensure(BL * n); W.lk_s.ensure(BL * n); W.lk_z.ensure(BL * n);
        W.sort_a.ensure(BL * n); W.sort_t.ensure(BL * n); W.lk_ext.ensure(BL * 3 * en);
        W.raw_la.ensure(BL * (pk.bf + 1) * 8); W.raw_ls.ensure(BL * (pk.bf + 1) * 8); W.raw_lz.ensure(BL * pk.bf * 8);
        W.carries.ensure(BL);
    }
    W.d_error.ensure(B);
    W.raw_adv.ensure(std::max<size_t>(1, B * pk.A * (pk.bf + 1) * 8)); W.raw_z.ensure(std::max<size_t>(1, B * pk.P * pk.bf * 8));
    W.seeds.ensure(B * 32 * ((pk.n + vanishing_chunk(pk.n) - 1) / vanishing_chunk(pk.n))); W.ch.ensure(B);
    size_t max_pts = B * std::max<size_t>(std::max<size_t>(pk.A, pk.P + pk.L + 1), std::max<size_t>(pk.Q, 2 * pk.L + 1));
    W.aff.ensure(max_pts);
    W.h_aff.ensure(max_pts * sizeof(g1_affine_t)); W.h_evals.ensure(B * (pk.num_evals + 1) * sizeof(fr_t));
}

// Per-proof status codes (include/zkgpu.h)
enum { PROOF_OK = 0, PROOF_LOOKUP_FAILED = 1 };
struct RngRef { int mode; uint8_t* data; };   // data: seed (u64), xoshiro state (4 x u64, written back) or ChaCha20 seed (32 B)
// A sub-batch of B proofs of one circuit.  Advice is either one contiguous [B][A][n] block (host or device memory) or one
// host pointer per proof (coalesced single-proof requests); everything else is contiguous.
struct BatchView {
    size_t B = 0, num_pi = 0;
    const fr_t* advice = nullptr; bool advice_on_device = false;
    const fr_t* const* advice_ptrs = nullptr;
    const fr_t* instance = nullptr;     // [B][num_pi]
    const RngRef* rng = nullptr;        // [B]
    uint8_t* proofs = nullptr;          // [B][proof_len]
    int32_t* status = nullptr;          // [B], may be null
    // the worker's next sub-batch (contiguous host advice only): uploaded in the background
    const fr_t* next_advice = nullptr; size_t next_B = 0;
};

static void prove_sub_batch(PkEntry& pk, ProverWs& W, const BatchView& V) {
    Context& C = *pk.C;
    const CsDesc& cs = pk.cs;
    if (!W.stream) ZK_CUDA(cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking));
    cudaStream_t st = W.stream;
    const size_t B = V.B, num_pi = V.num_pi;
    const fr_t* advice = V.advice; const bool advice_on_device = V.advice_on_device;
    const fr_t* instance = V.instance; uint8_t* proofs = V.proofs;
    const fr_t* next_advice = V.next_advice; const size_t next_B = V.next_B;
    const size_t n = pk.n, en = pk.cn /* coset-major rows per column */, A = pk.A, P = pk.P, Q = pk.Q, bf = pk.bf;
    const size_t ns = pk.plan.sets.size();
    ensure_ws(pk, W, B);
    const fr_t one = fe_one<FrTag>();
    // pinned staging arena for this sub-batch's small uploads (see StageArena)
    W.h_stage.ensure(std::max<size_t>((size_t)4 << 20, B * ((size_t)(pk.A + 2 * pk.L + pk.P) * (pk.bf + 1) * 64 + (pk.num_evals + 64) * 128 + 4096)));
    StageArena arena;
    arena.base = W.h_stage.as<uint8_t>(); arena.cap = W.h_stage.n;
    struct ArenaScope { StageArena* prev; explicit ArenaScope(StageArena* a) : prev(tl_arena) { tl_arena = a; } ~ArenaScope() { tl_arena = prev; } } arena_scope(&arena);

    StepTimer timer;
    // ---- step 0: transcripts, RNG streams, uploads ------------------------------------------
    std::vector<ProofState> ps;
    ps.reserve(B);
    const size_t L = pk.L;
    std::vector<uint64_t> raw_adv(B * A * (bf + 1) * 8), raw_z(B * P * bf * 8);
    std::vector<uint64_t> raw_la(B * L * (bf + 1) * 8), raw_ls(B * L * (bf + 1) * 8), raw_lz(B * L * bf * 8);
    const size_t vchunk = vanishing_chunk(n), vnch = (n + vchunk - 1) / vchunk;
    std::vector<uint8_t> cseeds(B * vnch * 32);
    std::vector<int32_t> status(B, PROOF_OK);
    for (size_t b = 0; b < B; ++b) {
        ps.emplace_back(proofs + b * pk.proof_len);
        ProofState& p = ps.back();
        p.tr.common_scalar(pk.digest);
        for (size_t i = 0; i < num_pi; ++i) p.tr.common_scalar(instance[b * num_pi + i]);
        // The proof's whole rng stream is drawn up front, in the order create_proof draws it
        ProofRng rng(V.rng[b].mode, V.rng[b].data);
        // advice blinding rows column by column, then one unused Blind per column
        for (size_t t = 0; t < A * (bf + 1); ++t) rng.next_wide(&raw_adv[(b * A * (bf + 1) + t) * 8]);
        for (size_t c = 0; c < A; ++c) rng.skip_wide();
        // lookups (commit_permuted): blinding rows of the permuted input, of the permuted table, then the two Blinds
        for (size_t l = 0; l < L; ++l) {
            for (size_t t = 0; t < bf + 1; ++t) rng.next_wide(&raw_la[((b * L + l) * (bf + 1) + t) * 8]);
            for (size_t t = 0; t < bf + 1; ++t) rng.next_wide(&raw_ls[((b * L + l) * (bf + 1) + t) * 8]);
            rng.skip_wide(); rng.skip_wide();
        }
        // permutation: per set bf blinding rows, then its unused Blind
        for (size_t s = 0; s < P; ++s) {
            for (size_t t = 0; t < bf; ++t) rng.next_wide(&raw_z[((b * P + s) * bf + t) * 8]);
            rng.skip_wide();
        }
        // lookups (commit_product): bf blinding rows of z, then its Blind
        for (size_t l = 0; l < L; ++l) {
            for (size_t t = 0; t < bf; ++t) rng.next_wide(&raw_lz[((b * L + l) * bf + t) * 8]);
            rng.skip_wide();
        }
        // vanishing: ChaCha20 seeds of the random polynomial's chunks, its Blind, then the Q quotient-piece Blinds: the
        // last draws of create_proof (KZG ignores every Blind, SHPLONK draws nothing)
        for (size_t c = 0; c < vnch; ++c) rng.fill_bytes32(&cseeds[(b * vnch + c) * 32]);
        rng.skip_wide();
        for (size_t i = 0; i < Q; ++i) rng.skip_wide();
        rng.store_state(V.rng[b].data);   // a running SmallRng continues after the proof, as `&mut rng` does upstream
    }
    timer.lap(7);   // host-only part of step 0 (rng streams, transcripts)
    if (V.advice_ptrs) {
        for (size_t b = 0; b < B; ++b)
            ZK_CUDA(cudaMemcpyAsync(W.adv.p + b * A * n, V.advice_ptrs[b], A * n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    } else
    if (advice_on_device) ZK_CUDA(cudaMemcpyAsync(W.adv.p, advice, B * A * n * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
    else if (W.prefetched_src == advice) {
        // uploaded while the previous sub-batch was computing: take the staging buffer
        ZK_CUDA(cudaStreamWaitEvent(st, W.copy_ev, 0));
        std::swap(W.adv.p, W.adv_next.p); std::swap(W.adv.n, W.adv_next.n);
        W.prefetched_src = nullptr;
    } else upload_advice(W.adv.p, advice, B * A * n * sizeof(fr_t), st, false);
    // the next sub-batch's advice is pulled in the background once this one's own upload has landed (issued after the
    // first commitments are back, so the two transfers do not share the PCIe link)
    auto issue_prefetch = [&]() {
        if (!next_advice || advice_on_device) return;
        if (!W.copy_stream) {
            ZK_CUDA(cudaStreamCreateWithFlags(&W.copy_stream, cudaStreamNonBlocking));
            ZK_CUDA(cudaEventCreateWithFlags(&W.copy_ev, cudaEventDisableTiming));
        }
        W.adv_next.ensure(std::max(W.B, next_B) * A * n);
        upload_advice(W.adv_next.p, next_advice, next_B * A * n * sizeof(fr_t), W.copy_stream, true);
        ZK_CUDA(cudaEventRecord(W.copy_ev, W.copy_stream));
        W.prefetched_src = next_advice;
    };
    ZK_CUDA(cudaMemsetAsync(W.inst.p, 0, B * n * sizeof(fr_t), st));
    if (num_pi) {
        const void* staged = arena.put(instance, B * num_pi * sizeof(fr_t));
        ZK_CUDA(cudaMemcpy2DAsync(W.inst.p, n * sizeof(fr_t), staged ? staged : (const void*)instance, num_pi * sizeof(fr_t), num_pi * sizeof(fr_t), B,
                                  cudaMemcpyHostToDevice, st));
    }
    h2d(W.raw_adv.p, raw_adv, st); h2d(W.raw_z.p, raw_z, st); h2d(W.seeds.p, cseeds, st);
    if (L) { h2d(W.raw_la.p, raw_la, st); h2d(W.raw_ls.p, raw_ls, st); h2d(W.raw_lz.p, raw_lz, st); ZK_CUDA(cudaMemsetAsync(W.d_error.p, 0, B * sizeof(int), st)); }

    // Timing class KT_HOSTGAP (zkgpu_kernel_timing only): device time between the end of the work queued before a Fiat-Shamir round
    // trip and the first operation the host queues after it — the GPU idles there when a single pipeline worker runs (the headline
    // passes hide it behind the other workers' kernels).
    bool gap_open = false;
    auto gap_begin = [&]() { if (g_ktime_on && !gap_open) { ktime_begin(KT_HOSTGAP, st); gap_open = true; } };
    auto gap_end = [&]() { if (gap_open) { ktime_end(KT_HOSTGAP, st); gap_open = false; } };
    auto fetch_points = [&](size_t count) -> const g1_affine_t* {
        ZK_CUDA(cudaMemcpyAsync(W.h_aff.p, W.aff.p, count * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, st));
        gap_begin();
        ZK_CUDA(cudaStreamSynchronize(st));
        return W.h_aff.as<g1_affine_t>();
    };
    std::vector<Challenges> ch(B);
    auto push_challenges = [&]() {
        gap_end();
        for (size_t b = 0; b < B; ++b) { ch[b].theta = ps[b].theta; ch[b].beta = ps[b].beta; ch[b].gamma = ps[b].gamma; ch[b].y = ps[b].y; ch[b].x = ps[b].x; }
        h2d(W.ch.p, ch, st);
    };

    timer.lap(0);
    // ---- step 1: blind + commit advice --------------------------------------------------------
    launch_scatter_random(W.adv.p, A * n, n, pk.ustart, W.raw_adv.p, B, A, bf + 1, st);
    trace_dev("advice_blinded", W.adv.p, n, A, n, st);
    // Single-proof regime: the advice commitments leave most of the GPU idle (a few dozen MSMs), and the advice / instance
    // transforms need nothing but the blinded values: they run on a side stream meanwhile, into separate buffers (the Lagrange
    // values are still needed by the lookup and permutation arguments), and the buffers are swapped where the in-place transform
    // would have happened.
    const bool overlap = commit_lat_ok(C, pk, B * A) && !g_trace;
    if (overlap) {
        if (!W.side) {
            ZK_CUDA(cudaStreamCreateWithFlags(&W.side, cudaStreamNonBlocking));
            ZK_CUDA(cudaEventCreateWithFlags(&W.ev_blind, cudaEventDisableTiming));
            ZK_CUDA(cudaEventCreateWithFlags(&W.ev_side, cudaEventDisableTiming));
            ZK_CUDA(cudaEventCreateWithFlags(&W.ev_z, cudaEventDisableTiming));
            ZK_CUDA(cudaEventCreateWithFlags(&W.ev_zside, cudaEventDisableTiming));
        }
        W.adv_coef.ensure(W.adv.n); W.inst_coef.ensure(W.inst.n); W.nd.ensure(std::max<size_t>(1, 2 * B * (P + L) * n));
        ZK_CUDA(cudaEventRecord(W.ev_blind, st));
        ZK_CUDA(cudaStreamWaitEvent(W.side, W.ev_blind, 0));
        intt_n(pk, W, W.adv.p, B * A, W.side, W.adv_coef.p, &W.scratch2);
        intt_n(pk, W, W.inst.p, B, W.side, W.inst_coef.p, &W.scratch2);
        coset_ext(pk, W, W.adv_coef.p, W.adv_ext.p, B, A, (A + 1) * en, W.side, &W.scratch2);
        coset_ext(pk, W, W.inst_coef.p, W.adv_ext.p + A * en, B, 1, (A + 1) * en, W.side, &W.scratch2);
        ZK_CUDA(cudaEventRecord(W.ev_side, W.side));
    }
    commit(C, pk, W, 1, W.adv.p, B * A, 0, 0, W.aff.p, st);
    {
        const g1_affine_t* pts = fetch_points(B * A);
        issue_prefetch();
        for (size_t b = 0; b < B; ++b) {
            for (size_t c = 0; c < A; ++c) ps[b].tr.write_point(pts[b * A + c]);
            ps[b].theta = ps[b].tr.squeeze();
            ps[b].beta = fr_t::zero(); ps[b].gamma = fr_t::zero(); ps[b].y = fr_t::zero(); ps[b].x = fr_t::zero();
        }
    }
    if (L) {
        // lookup arguments, part 1: compress with theta, permute the (input, table) pairs, blind, commit
        push_challenges();
        LookupCompressArgs la;
        la.adv = W.adv.p; la.adv_proof_stride = A * n; la.inst = W.inst.p; la.inst_proof_stride = n; la.fixed_vals = pk.fixed_vals.p;
        la.ch = W.ch.p; la.lp = lookup_progs(pk); la.k = pk.k;
        launch_lookup_compress(la, W.lk_in.p, W.lk_tab.p, B, st);
        launch_lookup_permute(W.lk_in.p, W.lk_tab.p, W.lk_a.p, W.lk_s.p, W.sort_a.p, W.sort_t.p, pk.k, pk.ustart, B * L, (unsigned)L, W.d_error.p, st);
        launch_scatter_random(W.lk_a.p, L * n, n, pk.ustart, W.raw_la.p, B, L, bf + 1, st);
        launch_scatter_random(W.lk_s.p, L * n, n, pk.ustart, W.raw_ls.p, B, L, bf + 1, st);
        trace_dev("lookup_permuted_input", W.lk_a.p, n, L, n, st);
        trace_dev("lookup_permuted_table", W.lk_s.p, n, L, n, st);
        commit(C, pk, W, 1, W.lk_a.p, B * L, 0, 0, W.aff.p, st);
        commit(C, pk, W, 1, W.lk_s.p, B * L, 0, 0, W.aff.p + B * L, st);
        // A witness whose lookup input is missing from its table has no permuted pair: halo2 returns
        // Error::ConstraintSystemFailure for that proof.  The flag is per proof: the others of the sub-batch are unaffected
        // (the failed one keeps moving through the pipeline on whatever the kernel wrote; its output is zeroed at the end).
        std::vector<int> err(B, 0);
        ZK_CUDA(cudaMemcpyAsync(err.data(), W.d_error.p, B * sizeof(int), cudaMemcpyDeviceToHost, st));
        const g1_affine_t* pts = fetch_points(2 * B * L);
        for (size_t b = 0; b < B; ++b) if (err[b]) status[b] = PROOF_LOOKUP_FAILED;
        for (size_t b = 0; b < B; ++b)
            for (size_t l = 0; l < L; ++l) { ps[b].tr.write_point(pts[b * L + l]); ps[b].tr.write_point(pts[B * L + b * L + l]); }
    }
    for (size_t b = 0; b < B; ++b) { ps[b].beta = ps[b].tr.squeeze(); ps[b].gamma = ps[b].tr.squeeze(); }
    push_challenges();

    timer.lap(1);
    // ---- step 2: permutation grand products, random polynomial --------------------------------
    if (P) {
        fr_t* num = overlap ? W.nd.p : W.adv_ext.p; fr_t* den = num + B * P * n;  // adv_ext is free until step 4 (unless the side stream fills it)
        PermArgs pa;
        pa.adv = W.adv.p; pa.adv_proof_stride = A * n; pa.inst = W.inst.p; pa.inst_proof_stride = n;
        pa.fixed_vals = pk.fixed_vals.p; pa.sigma_vals = pk.sigma_vals.p; pa.cols = pk.cols.p; pa.delta_pows = pk.delta_pows.p;
        pa.omega_tw = pk.omega_tw; pa.ch = W.ch.p; pa.k = pk.k; pa.S = pk.S; pa.chunk = pk.chunk; pa.P = pk.P;
        KtScope kt(KT_PERM, st);
        launch_perm_num_den(pa, num, den, B, st);
        launch_batch_inverse(den, B * P * n, st);
        launch_perm_scan(num, den, W.z.p, pk.k, B * P, st);
        W.carries.ensure(B * P);
        launch_perm_finalize(W.z.p, W.carries.p, pk.k, pk.P, pk.bf, W.raw_z.p, B, st);
    }
    if (P) trace_dev("z", W.z.p, n, P, n, st);
    if (overlap && P) {
        // same idea for the permutation products: their transforms run beside the round's commitments
        W.z_coef.ensure(W.z.n);
        ZK_CUDA(cudaEventRecord(W.ev_z, st));
        ZK_CUDA(cudaStreamWaitEvent(W.side, W.ev_z, 0));
        intt_n(pk, W, W.z.p, B * P, W.side, W.z_coef.p, &W.scratch2);
        coset_ext(pk, W, W.z_coef.p, W.z_ext.p, B, P, P * en, W.side, &W.scratch2);
        ZK_CUDA(cudaEventRecord(W.ev_zside, W.side));
    }
    if (L) {
        // lookup arguments, part 2: grand products
        fr_t* num = (overlap ? W.nd.p : W.adv_ext.p) + 2 * B * P * n; fr_t* den = num + B * L * n;
        {
            KtScope kt(KT_PERM, st);
            launch_lookup_num_den(W.lk_in.p, W.lk_tab.p, W.lk_a.p, W.lk_s.p, W.ch.p, num, den, pk.k, pk.L, B, st);
            launch_batch_inverse(den, B * L * n, st);
            launch_perm_scan(num, den, W.lk_z.p, pk.k, B * L, st);
            launch_perm_finalize(W.lk_z.p, W.carries.p, pk.k, 1, pk.bf, W.raw_lz.p, B * L, st);
        }
        trace_dev("lookup_z", W.lk_z.p, n, L, n, st);
    }
    launch_chacha_poly(W.seeds.p, W.randp.p, n, B, vchunk, vnch, st);
    trace_dev("random_poly", W.randp.p, n, 1, n, st);
    // the round's commitments: permutation products, lookup products (both over g_lagrange) and the random polynomial (over g)
    if (commit_lat_ok(C, pk, B * (P + L + 1))) {
        std::vector<CommitItem> items;
        for (size_t i = 0; i < B * P; ++i) items.push_back({1, W.z.p + i * n});
        for (size_t i = 0; i < B * L; ++i) items.push_back({1, W.lk_z.p + i * n});
        for (size_t b = 0; b < B; ++b) items.push_back({0, W.randp.p + b * n});
        commit_lat(C, pk, W, items, W.aff.p, st);
    } else {
        if (P) commit(C, pk, W, 1, W.z.p, B * P, 0, 0, W.aff.p, st);
        if (L) commit(C, pk, W, 1, W.lk_z.p, B * L, 0, 0, W.aff.p + B * P, st);
        commit(C, pk, W, 0, W.randp.p, B, 0, 0, W.aff.p + B * (P + L), st);
    }
    ZK_CUDA(cudaMemcpyAsync(W.h_aff.p, W.aff.p, B * (P + L + 1) * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, st));
    if (!W.ev_pts) ZK_CUDA(cudaEventCreateWithFlags(&W.ev_pts, cudaEventDisableTiming));
    cudaEvent_t ev_pts = W.ev_pts;
    ZK_CUDA(cudaEventRecord(ev_pts, st));
    // queue the transforms that do not depend on y behind the commitments
    if (overlap && P) {
        ZK_CUDA(cudaStreamWaitEvent(st, W.ev_zside, 0));
        std::swap(W.z.p, W.z_coef.p); std::swap(W.z.n, W.z_coef.n);
    } else if (P) {
        intt_n(pk, W, W.z.p, B * P, st);
        coset_ext(pk, W, W.z.p, W.z_ext.p, B, P, P * en, st);
    }
    if (L) {
        intt_n(pk, W, W.lk_z.p, B * L, st); intt_n(pk, W, W.lk_a.p, B * L, st); intt_n(pk, W, W.lk_s.p, B * L, st);
        coset_ext(pk, W, W.lk_z.p, W.lk_ext.p, B * L, 1, 3 * en, st);
        coset_ext(pk, W, W.lk_a.p, W.lk_ext.p + en, B * L, 1, 3 * en, st);
        coset_ext(pk, W, W.lk_s.p, W.lk_ext.p + 2 * en, B * L, 1, 3 * en, st);
    }
    if (overlap) {
        ZK_CUDA(cudaStreamWaitEvent(st, W.ev_side, 0));
        std::swap(W.adv.p, W.adv_coef.p); std::swap(W.adv.n, W.adv_coef.n);
        std::swap(W.inst.p, W.inst_coef.p); std::swap(W.inst.n, W.inst_coef.n);
    } else {
        intt_n(pk, W, W.adv.p, B * A, st);
        intt_n(pk, W, W.inst.p, B, st);
        coset_ext(pk, W, W.adv.p, W.adv_ext.p, B, A, (A + 1) * en, st);
        coset_ext(pk, W, W.inst.p, W.adv_ext.p + A * en, B, 1, (A + 1) * en, st);
    }
    ZK_CUDA(cudaEventSynchronize(ev_pts));
    {
        const g1_affine_t* pts = W.h_aff.as<g1_affine_t>();
        for (size_t b = 0; b < B; ++b) {
            for (size_t s = 0; s < P; ++s) ps[b].tr.write_point(pts[b * P + s]);
            for (size_t l = 0; l < L; ++l) ps[b].tr.write_point(pts[B * P + b * L + l]);
            ps[b].tr.write_point(pts[B * (P + L) + b]);
            ps[b].y = ps[b].tr.squeeze();
        }
    }
    push_challenges();
    trace_dev("advice_poly", W.adv.p, n, A, n, st);
    trace_dev("z_poly", W.z.p, n, P, n, st);
    trace_dev("z_quotient_cosets", W.z_ext.p, en, P, en, st);
    trace_dev("advice_quotient_cosets", W.adv_ext.p, en, A, en, st);
    trace_dev("instance_quotient_cosets", W.adv_ext.p + A * en, en, 1, en, st);

    timer.lap(2);
    // ---- step 3: quotient ---------------------------------------------------------------------
    {
        EvalHArgs ea;
        ea.adv_ext = W.adv_ext.p; ea.adv_ext_proof_stride = (A + 1) * en; ea.z_ext = W.z_ext.p; ea.z_ext_proof_stride = P * en;
        ea.fixed_ext = pk.fixed_ext.p; ea.sigma_ext = pk.sigma_ext.p; ea.l0 = pk.l0.p; ea.l_last = pk.l_last.p; ea.l_active = pk.l_active.p;
        ea.t_inv = pk.t_inv.p; ea.ext_tw = pk.ext_tw; ea.delta_pows = pk.delta_pows.p; ea.cols = pk.cols.p; ea.ch = W.ch.p;
        ea.prog = pk.prog.p; ea.gate_off = pk.gate_off.p; ea.constants = pk.constants.p;
        ea.adv_q = pk.adv_q.p; ea.fix_q = pk.fix_q.p; ea.inst_q = pk.inst_q.p;
        ea.n_prog = (unsigned)(pk.prog.n / 2); ea.n_adv_q = (unsigned)cs.advice_queries.size(); ea.n_fix_q = (unsigned)cs.fixed_queries.size();
        ea.n_inst_q = (unsigned)cs.instance_queries.size();
        ea.num_gates = (unsigned)cs.gates.size(); ea.A = pk.A; ea.S = pk.S; ea.chunk = pk.chunk; ea.P = pk.P; ea.k = pk.k; ea.ek = pk.ek;
        ea.rotation_last = pk.rot_last; ea.zeta = pk.zeta; ea.Qc = pk.Qc;
        ea.lk_ext = W.lk_ext.p; ea.lk_ext_proof_stride = 3 * L * en; ea.lp = lookup_progs(pk);
        launch_eval_h(ea, W.h.p, B, st);
    }
    trace_dev("h_quotient_cosets", W.h.p, en, 1, en, st);
    intt_n(pk, W, W.h.p, B * pk.Qc, st);
    launch_coset_combine(W.h.p, pk.coset_pows_inv.p, pk.vinv.p, pk.Qc, pk.k, B, st);
    trace_dev("h_coeffs", W.h.p, Q * n, 1, en, st);
    commit(C, pk, W, 0, W.h.p, B * Q, Q, en, W.aff.p, st);
    {
        const g1_affine_t* pts = fetch_points(B * Q);
        for (size_t b = 0; b < B; ++b) {
            for (size_t i = 0; i < Q; ++i) ps[b].tr.write_point(pts[b * Q + i]);
            ps[b].x = ps[b].tr.squeeze();
            ps[b].xn = ps[b].x;
            for (unsigned i = 0; i < pk.k; ++i) ps[b].xn = sqr(ps[b].xn);
        }
    }

    timer.lap(3);
    // ---- step 4: evaluations ------------------------------------------------------------------
    const size_t NE = pk.num_evals;
    {
        // h(X) = sum_i xn^i * piece_i
        std::vector<LinTerm> terms(B * Q);
        std::vector<uint32_t> off(B + 1);
        std::vector<fr_t*> outs(B);
        for (size_t b = 0; b < B; ++b) {
            fr_t c = one;
            off[b] = (uint32_t)(b * Q);
            for (size_t i = 0; i < Q; ++i) { terms[b * Q + i].poly = W.h.p + b * en + i * n; terms[b * Q + i].coef = c; c = c * ps[b].xn; }
            outs[b] = W.hpoly.p + b * n;
        }
        off[B] = (uint32_t)(B * Q);
        gap_end();
        upload(W.terms, terms, st); upload(W.job_off, off, st); upload(W.outs, outs, st);
        launch_lincomb(W.terms.p, W.job_off.p, W.outs.p, B, n, st);

        std::vector<EvalJob> jobs(B * (NE + 1));
        for (size_t b = 0; b < B; ++b) {
            ProofState& p = ps[b];
            EvalJob* j = &jobs[b * (NE + 1)];
            const fr_t x = p.x;
            std::map<int, fr_t> rx;
            auto at = [&](int r) -> const fr_t& {
                auto it = rx.find(r);
                if (it == rx.end()) it = rx.emplace(r, host_rotate(pk, x, r)).first;
                return it->second;
            };
            for (auto& q : cs.advice_queries) { j->poly = W.adv.p + (b * A + q.column) * n; j->x = at(q.rotation); ++j; }
            for (auto& q : cs.fixed_queries) { j->poly = pk.fixed_polys.p + (size_t)q.column * n; j->x = at(q.rotation); ++j; }
            j->poly = W.randp.p + b * n; j->x = x; ++j;
            for (size_t c = 0; c < pk.S; ++c) { j->poly = pk.sigma_polys.p + c * n; j->x = x; ++j; }
            for (size_t s = 0; s < P; ++s) {
                const fr_t* zp = W.z.p + (b * P + s) * n;
                j->poly = zp; j->x = x; ++j;
                j->poly = zp; j->x = at(1); ++j;
                if (s + 1 < P) { j->poly = zp; j->x = at(pk.rot_last); ++j; }
            }
            for (size_t l = 0; l < L; ++l) {
                const fr_t *zp = W.lk_z.p + (b * L + l) * n, *ap = W.lk_a.p + (b * L + l) * n, *sp = W.lk_s.p + (b * L + l) * n;
                j->poly = zp; j->x = x; ++j;
                j->poly = zp; j->x = at(1); ++j;
                j->poly = ap; j->x = x; ++j;
                j->poly = ap; j->x = at(-1); ++j;
                j->poly = sp; j->x = x; ++j;
            }
            j->poly = W.hpoly.p + b * n; j->x = x; ++j;  // quotient evaluation: computed, not written
            ZK_REQUIRE((size_t)(j - &jobs[b * (NE + 1)]) == NE + 1, "internal: evaluation count mismatch");
            p.points.clear();
            for (int r : pk.plan.superset) p.points.push_back(at(r));
        }
        upload(W.eval_jobs, jobs, st);
        launch_poly_eval(W.eval_jobs.p, W.evals.p, jobs.size(), pk.k, st);
        ZK_CUDA(cudaMemcpyAsync(W.h_evals.p, W.evals.p, jobs.size() * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        gap_begin();
        ZK_CUDA(cudaStreamSynchronize(st));
        const fr_t* he = W.h_evals.as<fr_t>();
        for (size_t b = 0; b < B; ++b) {
            ProofState& p = ps[b];
            p.evals.assign(he + b * (NE + 1), he + (b + 1) * (NE + 1));
            for (size_t i = 0; i < NE; ++i) p.tr.write_scalar(p.evals[i]);
            p.zeta = p.tr.squeeze(); p.nu = p.tr.squeeze();
        }
        if (g_trace) { g_trace("evals", ps[0].evals.data(), NE * sizeof(fr_t)); }
        trace_dev("h_poly", W.hpoly.p, n, 1, n, st);
    }

    timer.lap(4);
    // ---- step 5: SHPLONK h(X) -----------------------------------------------------------------
    auto poly_of = [&](size_t b, int id) -> const fr_t* {
        if (id < pk.plan.id_z0) return W.adv.p + (b * A + id) * n;
        if (id < pk.plan.id_lz0) return W.z.p + (b * P + (id - pk.plan.id_z0)) * n;
        if (id < pk.plan.id_la0) return W.lk_z.p + (b * L + (id - pk.plan.id_lz0)) * n;
        if (id < pk.plan.id_ls0) return W.lk_a.p + (b * L + (id - pk.plan.id_la0)) * n;
        if (id < pk.plan.id_fixed0) return W.lk_s.p + (b * L + (id - pk.plan.id_ls0)) * n;
        if (id < pk.plan.id_sigma0) return pk.fixed_polys.p + (size_t)(id - pk.plan.id_fixed0) * n;
        if (id < pk.plan.id_h) return pk.sigma_polys.p + (size_t)(id - pk.plan.id_sigma0) * n;
        return id == pk.plan.id_h ? W.hpoly.p + b * n : W.randp.p + b * n;
    };
    auto point_of = [&](const ProofState& p, int rot) -> const fr_t& {
        for (size_t i = 0; i < pk.plan.superset.size(); ++i) if (pk.plan.superset[i] == rot) return p.points[i];
        throw Error(ZK_ERR_INTERNAL, "rotation not in the superset");
    };
    std::vector<const fr_t*> set_num(B * ns);  // quotient of each set after its divisions
    {
        std::vector<LinTerm> terms;
        std::vector<uint32_t> off;
        std::vector<fr_t*> outs;
        std::vector<fr_t> low(B * ns * 4, fr_t::zero());
        std::vector<std::vector<DivJob>> rounds(4);
        for (size_t b = 0; b < B; ++b) {
            ProofState& p = ps[b];
            p.rcomb.assign(ns, {});
            // denominators of the Lagrange basis of every set, inverted together
            std::vector<fr_t> dens;
            for (auto& s : pk.plan.sets)
                for (size_t j = 0; j < s.rots.size(); ++j) {
                    fr_t d = one;
                    for (size_t t = 0; t < s.rots.size(); ++t) if (t != j) d = d * (point_of(p, s.rots[j]) - point_of(p, s.rots[t]));
                    dens.push_back(d);
                }
            host_batch_invert(dens);
            size_t di = 0;
            for (size_t si = 0; si < ns; ++si) {
                const RotSet& s = pk.plan.sets[si];
                const size_t m = s.rots.size();
                off.push_back((uint32_t)terms.size());
                std::vector<fr_t> w(m, fr_t::zero());  // sum_c zeta^c * eval_{c, rot j}
                fr_t zp = one;
                for (size_t c = 0; c < s.comms.size(); ++c) {
                    LinTerm t; t.poly = poly_of(b, s.comms[c]); t.coef = zp;
                    terms.push_back(t);
                    for (size_t j = 0; j < m; ++j) w[j] = w[j] + zp * p.evals[s.evals[c][j]];
                    zp = zp * p.zeta;
                }
                outs.push_back(W.comb.p + (b * ns + si) * n);
                // r(X) = sum_j w_j / den_j * prod_{t != j} (X - x_t)
                std::vector<fr_t> r(m, fr_t::zero());
                for (size_t j = 0; j < m; ++j) {
                    std::vector<fr_t> np{one};
                    for (size_t t = 0; t < m; ++t) {
                        if (t == j) continue;
                        const fr_t& xt = point_of(p, s.rots[t]);
                        std::vector<fr_t> nn(np.size() + 1, fr_t::zero());
                        for (size_t u = 0; u < np.size(); ++u) { nn[u + 1] = nn[u + 1] + np[u]; nn[u] = nn[u] - np[u] * xt; }
                        np.swap(nn);
                    }
                    fr_t sc = w[j] * dens[di + j];
                    for (size_t u = 0; u < m; ++u) r[u] = r[u] + np[u] * sc;
                }
                di += m;
                for (size_t u = 0; u < m; ++u) low[(b * ns + si) * 4 + u] = r[u];
                p.rcomb[si] = r;
                // divisions by (X - x_t), ping-pong between tmp1 / tmp2
                const fr_t* cur = W.comb.p + (b * ns + si) * n;
                for (size_t t = 0; t < m; ++t) {
                    DivJob dj;
                    dj.in = cur; dj.out = ((t & 1) ? W.tmp2.p : W.tmp1.p) + (b * ns + si) * n; dj.pt = point_of(p, s.rots[t]);
                    dj.low = t == 0 ? W.low.p + (b * ns + si) * 4 : nullptr;
                    rounds[t].push_back(dj);
                    cur = dj.out;
                }
                set_num[b * ns + si] = cur;
            }
        }
        off.push_back((uint32_t)terms.size());
        gap_end();
        upload(W.terms, terms, st); upload(W.job_off, off, st); upload(W.outs, outs, st);
        h2d(W.low.p, low, st);
        launch_lincomb(W.terms.p, W.job_off.p, W.outs.p, B * ns, n, st);
        trace_dev("set_combined", W.comb.p, n, ns, n, st);
        size_t total_div = 0;
        for (auto& r : rounds) total_div += r.size();
        W.div_jobs.ensure(std::max<size_t>(1, total_div));
        size_t doff = 0;
        for (auto& r : rounds) {
            if (r.empty()) continue;
            h2d(W.div_jobs.p + doff, r, st);
            launch_kate_div(W.div_jobs.p + doff, r.size(), pk.k, st);
            doff += r.size();
        }
        // h(X) = sum_i nu^i * num_i(X)
        std::vector<LinTerm> t2(B * ns);
        std::vector<uint32_t> off2(B + 1);
        std::vector<fr_t*> outs2(B);
        for (size_t b = 0; b < B; ++b) {
            fr_t np = one;
            off2[b] = (uint32_t)(b * ns);
            for (size_t si = 0; si < ns; ++si) { t2[b * ns + si].poly = set_num[b * ns + si]; t2[b * ns + si].coef = np; np = np * ps[b].nu; }
            outs2[b] = W.hx.p + b * n;
        }
        off2[B] = (uint32_t)(B * ns);
        upload(W.terms2, t2, st); upload(W.job_off2, off2, st); upload(W.outs2, outs2, st);
        launch_lincomb(W.terms2.p, W.job_off2.p, W.outs2.p, B, n, st);
        trace_dev("hx", W.hx.p, n, 1, n, st);
        commit(C, pk, W, 0, W.hx.p, B, 0, 0, W.aff.p, st);
        const g1_affine_t* pts = fetch_points(B);
        for (size_t b = 0; b < B; ++b) { ps[b].tr.write_point(pts[b]); ps[b].mu = ps[b].tr.squeeze(); }
    }

    timer.lap(5);
    // ---- step 6: SHPLONK linearisation L(X) / (X - mu) ----------------------------------------
    {
        std::vector<LinTerm> terms(B * (ns + 1));
        std::vector<uint32_t> off(B + 1);
        std::vector<fr_t*> outs(B);
        std::vector<fr_t> low(B * 4, fr_t::zero());
        std::vector<DivJob> divs(B);
        for (size_t b = 0; b < B; ++b) {
            ProofState& p = ps[b];
            auto zeval = [&](const std::vector<int>& rots) { fr_t a = one; for (int r : rots) a = a * (p.mu - point_of(p, r)); return a; };
            fr_t z_t = zeval(pk.plan.superset);
            std::vector<fr_t> zd(ns);
            for (size_t si = 0; si < ns; ++si) zd[si] = zeval(pk.plan.sets[si].diffs);
            fr_t zd0_inv = fe_inv(zd[0]);
            fr_t np = one, K = fr_t::zero();
            off[b] = (uint32_t)(b * (ns + 1));
            for (size_t si = 0; si < ns; ++si) {
                fr_t sc = np * zd[si];
                // r_i(mu)
                fr_t rm = fr_t::zero();
                for (size_t u = p.rcomb[si].size(); u-- > 0;) rm = rm * p.mu + p.rcomb[si][u];
                K = K + rm * sc;
                terms[b * (ns + 1) + si].poly = W.comb.p + (b * ns + si) * n;
                terms[b * (ns + 1) + si].coef = sc * zd0_inv;
                np = np * p.nu;
            }
            terms[b * (ns + 1) + ns].poly = W.hx.p + b * n;
            terms[b * (ns + 1) + ns].coef = neg(z_t * zd0_inv);
            low[b * 4] = K * zd0_inv;
            outs[b] = W.lx.p + b * n;
            divs[b].in = W.lx.p + b * n; divs[b].out = W.tmp1.p + b * n; divs[b].pt = p.mu; divs[b].low = W.low.p + b * 4;
        }
        off[B] = (uint32_t)(B * (ns + 1));
        gap_end();
        upload(W.terms, terms, st); upload(W.job_off, off, st); upload(W.outs, outs, st);
        h2d(W.low.p, low, st);
        launch_lincomb(W.terms.p, W.job_off.p, W.outs.p, B, n, st);
        W.div_jobs.ensure(B);
        h2d(W.div_jobs.p, divs, st);
        launch_kate_div(W.div_jobs.p, B, pk.k, st);
        trace_dev("wq", W.tmp1.p, n, 1, n, st);
        commit(C, pk, W, 0, W.tmp1.p, B, 0, 0, W.aff.p, st);
        const g1_affine_t* pts = fetch_points(B);
        for (size_t b = 0; b < B; ++b) {
            ps[b].tr.write_point(pts[b]);
            ZK_REQUIRE((size_t)(ps[b].tr.out - (proofs + b * pk.proof_len)) == pk.proof_len, "internal: proof length mismatch");
        }
    }
    gap_end();
    // failed proofs: no bytes, only a status (the reference fails that request alone, tee/.../server.rs:189-190)
    for (size_t b = 0; b < B; ++b) {
        if (status[b] != PROOF_OK) memset(proofs + b * pk.proof_len, 0, pk.proof_len);
        if (V.status) V.status[b] = status[b];
    }
    timer.lap(6);
}

// ---------------------------------------------------------------------------------------------
// zkgpu_prove_batch on ONE device: pipeline workers over sub-batches
// ---------------------------------------------------------------------------------------------
struct BatchArgs {
    const fr_t* advice; bool advice_on_device;
    const fr_t* instance; size_t num_pi; size_t m;
    const RngRef* rng; uint8_t* proofs; int32_t* status;
};
static void prove_batch_on_device(PkEntry& pk, const BatchArgs& a) {
    Context& C = *pk.C;
    C.bind();
    std::lock_guard<std::mutex> lk(pk.batch_mu);
    const size_t m = a.m;
    size_t Bmax = default_batch(pk);
    // Pipeline workers (host thread + stream + workspace each) take alternate sub-batches, so one worker's
    // transcript hashing and bookkeeping overlap the other's kernels.  Kernel-class timing and stage tracing
    // use a single worker (per-kernel durations are only meaningful when launches do not share the GPU).
    size_t nsub = (m + Bmax - 1) / Bmax;
    // Three workers from six sub-batches on (measured 646 -> 655 proofs/s: the digit sort and the other kernels that do not live on
    // the IMAD pipe find more IMAD-bound work of other streams to overlap with), two from two on; a fourth when the advice is already
    // resident (round 2: 699.8 -> 705.9 proofs/s; with host advice its extra background upload costs more than it gives, 680 -> 675).
    unsigned workers = (g_ktime_on || g_trace) ? 1 : (nsub >= 8 && a.advice_on_device) ? 4 : nsub >= 6 ? 3 : nsub >= 2 ? 2 : 1;
    if (const char* e = getenv("ZKGPU_PROVER_WORKERS")) { int v = atoi(e); if (v >= 1 && v <= (int)BATCH_WORKERS && (unsigned)v < workers) workers = (unsigned)v; }
    static const unsigned per_worker = [] { const char* e = getenv("ZKGPU_SUBBATCHES_PER_WORKER"); int v = e ? atoi(e) : 0; return (unsigned)(v >= 1 && v <= 8 ? v : 2); }();
    if (workers >= 2 && Bmax > 1) { Bmax = std::max<size_t>(1, std::min(Bmax, (m + per_worker * workers - 1) / (per_worker * workers))); nsub = (m + Bmax - 1) / Bmax; }
    // sub-batches as (offset, count).  With host advice the very first one is a quarter of the usual size: nothing can hide its
    // upload, so the GPU should start on a short one while the other worker's full-size upload is still in flight.
    std::vector<std::pair<size_t, size_t>> segs;
    {
        size_t off = 0;
        if (workers >= 2 && !a.advice_on_device && m > Bmax && Bmax >= 8) { segs.push_back({0, Bmax / 4}); off = Bmax / 4; }
        while (off < m) { size_t B = std::min(Bmax, m - off); segs.push_back({off, B}); off += B; }
    }
    const size_t adv_per = (size_t)pk.A * pk.n;
    auto run = [&](unsigned w) {
        C.bind();
        pk.ws[w].prefetched_src = nullptr;
        for (size_t i = w; i < segs.size(); i += workers) {
            const size_t off = segs[i].first;
            const bool has_next = i + workers < segs.size();
            BatchView V;
            V.B = segs[i].second; V.num_pi = a.num_pi;
            V.advice = a.advice + off * adv_per; V.advice_on_device = a.advice_on_device;
            V.instance = a.instance + off * a.num_pi; V.rng = a.rng + off;
            V.proofs = a.proofs + off * pk.proof_len; V.status = a.status ? a.status + off : nullptr;
            V.next_advice = has_next ? a.advice + segs[i + workers].first * adv_per : nullptr;
            V.next_B = has_next ? segs[i + workers].second : 0;
            prove_sub_batch(pk, pk.ws[w], V);
        }
    };
    if (workers == 1) { run(0); return; }
    std::exception_ptr err[BATCH_WORKERS];
    std::thread th[BATCH_WORKERS];
    for (unsigned w = 0; w < workers; ++w) th[w] = std::thread([&, w] { try { run(w); } catch (...) { err[w] = std::current_exception(); } });
    for (unsigned w = 0; w < workers; ++w) th[w].join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

// ---------------------------------------------------------------------------------------------
// zkgpu_prove_batch: shard the batch over the selected devices (replicas of the key material, contiguous
// shards of the proofs, one host thread per GPU, no collective)
// ---------------------------------------------------------------------------------------------
static void check_batch_args(const PkEntry& pk, const void* advice, const void* instance, size_t num_pi, size_t m, const void* rng,
                             const void* proofs, size_t proof_len, int rng_mode) {
    ZK_REQUIRE(proof_len == pk.proof_len, "prove: proof_len does not match the circuit (see zkgpu_pk_info)");
    ZK_REQUIRE(m == 0 || (advice && rng && proofs), "null pointer");
    ZK_REQUIRE(num_pi == 0 || instance, "null pointer");
    ZK_REQUIRE(num_pi <= pk.ustart, "prove: InstanceTooLarge");
    ZK_REQUIRE(rng_mode == RNG_SEED_U64 || rng_mode == RNG_XOSHIRO_STATE || rng_mode == RNG_CHACHA20_SEED, "prove: unknown rng_mode");
}

static void prove_batch(uint64_t handle, const uint64_t* advice, bool advice_on_device, const uint64_t* instance, size_t num_pi, size_t m,
                        int rng_mode, void* rng_data, uint8_t* proofs, size_t proof_len, int32_t* status) {
    Runtime& R = rt(); R.require();
    std::shared_ptr<PkShared> P = find_pk(handle);
    check_batch_args(*P->dev[0], advice, instance, num_pi, m, rng_data, proofs, proof_len, rng_mode);
    if (m == 0) return;
    std::vector<RngRef> rng(m);
    const size_t stride = rng_data_stride(rng_mode);
    for (size_t i = 0; i < m; ++i) { rng[i].mode = rng_mode; rng[i].data = static_cast<uint8_t*>(rng_data) + i * stride; }
    BatchArgs all{reinterpret_cast<const fr_t*>(advice), advice_on_device, reinterpret_cast<const fr_t*>(instance), num_pi, m, rng.data(), proofs, status};
    if (advice_on_device) {   // resident advice: the device that holds it proves the whole batch
        Context& C = R.of_pointer(advice);
        prove_batch_on_device(*P->dev[C.slot], all);
        return;
    }
    const size_t G = std::min<size_t>(P->dev.size(), m);
    if (G == 1) { prove_batch_on_device(*P->dev[0], all); return; }
    std::vector<std::thread> th(G);
    std::vector<std::exception_ptr> err(G);
    const size_t adv_per = (size_t)P->dev[0]->A * P->dev[0]->n;
    for (size_t g = 0; g < G; ++g) {
        const size_t lo = m * g / G, hi = m * (g + 1) / G;
        BatchArgs a = all;
        a.advice = all.advice + lo * adv_per; a.instance = all.instance + lo * num_pi; a.m = hi - lo; a.rng = all.rng + lo;
        a.proofs = proofs + lo * proof_len; a.status = status ? status + lo : nullptr;
        th[g] = std::thread([&, g, a] { try { prove_batch_on_device(*P->dev[g], a); } catch (...) { err[g] = std::current_exception(); } });
    }
    for (auto& t : th) t.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

// ---------------------------------------------------------------------------------------------
// zkgpu_prove: blocking single-proof call, coalesced
// ---------------------------------------------------------------------------------------------
static size_t coalesce_cap(const PkEntry& pk) {
    static const size_t env = [] { const char* e = getenv("ZKGPU_COALESCE_MAX"); long v = e ? atol(e) : 0; return v > 0 ? (size_t)v : (size_t)0; }();
    return std::max<size_t>(1, std::min(env ? env : (size_t)64, default_batch(pk)));
}
static void dispatcher_loop(PkShared* P, unsigned dev_slot, unsigned worker) {
    PkEntry& pk = *P->dev[dev_slot];
    Coalescer& co = P->co;
    ProverWs& W = pk.ws[BATCH_WORKERS + worker];
    try { pk.C->bind(); } catch (...) {}
    const size_t cap = coalesce_cap(pk);
    std::unique_lock<std::mutex> lk(co.mu);
    for (;;) {
        ++co.idle;
        co.cv_work.wait(lk, [&] { return co.stop || !co.q.empty(); });
        if (co.stop && co.q.empty()) { --co.idle; return; }
        // Take a fair share of what is waiting: with several idle dispatchers the queue is split between them, so their
        // sub-batches run concurrently (one's transcript hashing overlaps the other's kernels) instead of one after the other.
        const size_t share = (co.q.size() + co.idle - 1) / std::max(1u, co.idle);
        --co.idle;
        std::vector<ProveReq*> reqs;
        const size_t num_pi = co.q.front()->num_pi;
        while (!co.q.empty() && reqs.size() < std::min(cap, std::max<size_t>(1, share)) && co.q.front()->num_pi == num_pi) {
            reqs.push_back(co.q.front()); co.q.pop_front();
        }
        co.batches++; co.requests += reqs.size(); co.max_batch_seen = std::max<uint64_t>(co.max_batch_seen, reqs.size());
        lk.unlock();
        const size_t B = reqs.size();
        int rc = ZKGPU_OK; std::string msg;
        std::vector<int32_t> status(B, 0);
        try {
            std::vector<const fr_t*> adv(B);
            std::vector<fr_t> inst(B * num_pi);
            std::vector<RngRef> rng(B);
            std::vector<uint8_t> proofs(B * pk.proof_len);
            for (size_t b = 0; b < B; ++b) {
                adv[b] = reqs[b]->advice;
                if (num_pi) memcpy(&inst[b * num_pi], reqs[b]->instance, num_pi * sizeof(fr_t));
                rng[b].mode = reqs[b]->rng_mode; rng[b].data = reqs[b]->rng_data;
            }
            BatchView V;
            V.B = B; V.num_pi = num_pi; V.advice_ptrs = adv.data(); V.instance = inst.data(); V.rng = rng.data();
            V.proofs = proofs.data(); V.status = status.data();
            prove_sub_batch(pk, W, V);
            for (size_t b = 0; b < B; ++b) memcpy(reqs[b]->proof_out, &proofs[b * pk.proof_len], pk.proof_len);
        } catch (const zk::Error& e) { rc = e.code; msg = e.what(); }
        catch (const std::exception& e) { rc = ZKGPU_ERR_INTERNAL; msg = e.what(); }
        lk.lock();
        for (size_t b = 0; b < B; ++b) { reqs[b]->rc = rc; reqs[b]->err = msg; reqs[b]->status = status[b]; reqs[b]->done = true; }
        co.cv_done.notify_all();
    }
}
static void start_dispatchers(PkShared* P) {
    for (unsigned d = 0; d < P->dev.size(); ++d)
        for (unsigned w = 0; w < COALESCE_WORKERS; ++w) P->co.threads.emplace_back(dispatcher_loop, P, d, w);
}
PkShared::~PkShared() {
    { std::lock_guard<std::mutex> lk(co.mu); co.stop = true; }
    co.cv_work.notify_all();
    for (auto& t : co.threads) if (t.joinable()) t.join();
    for (void* p : co.free_slots) cudaFreeHost(p);
}
// a pinned staging slot for one request (blocks while all MAX_SLOTS are in use)
static void* acquire_slot(Coalescer& co, size_t bytes) {
    std::unique_lock<std::mutex> lk(co.mu);
    for (;;) {
        if (!co.free_slots.empty()) { void* p = co.free_slots.back(); co.free_slots.pop_back(); return p; }
        if (co.slots_allocated < Coalescer::MAX_SLOTS) {
            ++co.slots_allocated;
            lk.unlock();
            void* p = nullptr;
            cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
            if (e != cudaSuccess) { cudaGetLastError(); lk.lock(); --co.slots_allocated; throw Error(ZK_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
            return p;
        }
        co.cv_slot.wait(lk);
    }
}
static void release_slot(Coalescer& co, void* p) {
    { std::lock_guard<std::mutex> lk(co.mu); co.free_slots.push_back(p); }
    co.cv_slot.notify_one();
}

}  // namespace zk

using namespace zk;

extern "C" {

void zkgpu_prover_step_seconds(double out[8], int reset) {
    std::lock_guard<std::mutex> lk(g_step_mu);
    for (int i = 0; i < 8; ++i) { out[i] = g_step_s[i]; if (reset) g_step_s[i] = 0; }
}
void zkgpu_set_trace(void (*fn)(const char*, const void*, size_t)) { g_trace = fn; }
int zkgpu_set_rayon_threads(unsigned num_threads) {
    API_TRY
    ZK_REQUIRE(num_threads >= 1 && num_threads <= 65536, "rayon thread count out of range");
    g_rayon_threads.store(num_threads);
    API_END
}

static int pk_create_or_load(uint64_t srs, const uint8_t* circuit_blob, size_t blob_len, const uint8_t* pk_bin, size_t pk_bin_len, uint64_t* pk_out) {
    API_TRY
    ZK_REQUIRE(circuit_blob && pk_out, "null pointer");
    Runtime& R = rt(); R.require();
    std::shared_ptr<PkShared> P(new PkShared);
    const size_t G = R.devs.size();
    P->dev.resize(G);
    // keygen on every selected device (replicas of the proving key), one host thread each
    std::vector<std::exception_ptr> err(G);
    auto one = [&](size_t g) {
        try {
            DeviceScope scope(*R.devs[g]);
            P->dev[g] = keygen(*R.devs[g], srs, circuit_blob, blob_len, pk_bin, pk_bin_len);
        } catch (...) { err[g] = std::current_exception(); }
    };
    if (G == 1) one(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) th.emplace_back(one, g);
        for (auto& t : th) t.join();
    }
    for (auto& e : err) if (e) std::rethrow_exception(e);
    std::unique_lock<std::shared_mutex> tl(R.tab_mu);
    uint64_t h = g_next_pk++;
    g_pks[h] = std::move(P);
    *pk_out = h;
    API_END
}
int zkgpu_pk_create(uint64_t srs, const uint8_t* circuit_blob, size_t blob_len, uint64_t* pk_out) {
    return pk_create_or_load(srs, circuit_blob, blob_len, nullptr, 0, pk_out);
}
int zkgpu_pk_load(uint64_t srs, const uint8_t* cs_blob, size_t cs_blob_len, const uint8_t* pk_bin, size_t pk_bin_len, uint64_t* pk_out) {
    if (!pk_bin) { g_last_error = "null pointer"; return ZKGPU_ERR_ARG; }
    return pk_create_or_load(srs, cs_blob, cs_blob_len, pk_bin, pk_bin_len, pk_out);
}
int zkgpu_pk_release(uint64_t pk) {
    API_TRY
    std::shared_ptr<PkShared> P;
    {
        std::unique_lock<std::shared_mutex> tl(rt().tab_mu);
        auto it = g_pks.find(pk);
        ZK_REQUIRE(it != g_pks.end(), "unknown proving key handle");
        P = it->second;
        g_pks.erase(it);
    }
    P.reset();   // the last user (a call still in flight keeps its own reference) frees the device memory
    API_END
}
int zkgpu_pk_info(uint64_t pk, uint64_t info[16]) {
    API_TRY
    ZK_REQUIRE(info, "null pointer");
    std::shared_ptr<PkShared> P = find_pk(pk);
    const PkEntry& p = *P->dev[0];
    p.C->bind();
    uint64_t v[16] = {p.k, p.n, p.A, p.F, p.cs.degree(), p.bf, p.P, p.Q, p.num_evals, p.proof_len, p.ek, p.S, p.plan.sets.size(), default_batch(p),
                      P->dev.size(), 0};
    memcpy(info, v, sizeof v);
    API_END
}
int zkgpu_pk_vk(uint64_t pk, uint64_t* fixed_commitments, uint64_t* perm_commitments, uint64_t digest[4]) {
    API_TRY
    std::shared_ptr<PkShared> P = find_pk(pk);
    const PkEntry& p = *P->dev[0];
    if (fixed_commitments && !p.fixed_commitments.empty()) memcpy(fixed_commitments, p.fixed_commitments.data(), p.fixed_commitments.size() * 64);
    if (perm_commitments && !p.perm_commitments.empty()) memcpy(perm_commitments, p.perm_commitments.data(), p.perm_commitments.size() * 64);
    if (digest) memcpy(digest, p.digest.l, 32);
    API_END
}
int zkgpu_prove_batch_rng(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, size_t m,
                          int rng_mode, void* rng_data, uint8_t* proofs_out, size_t proof_len, int32_t* status_out) {
    API_TRY
    prove_batch(pk, advice, false, instance, num_instance, m, rng_mode, rng_data, proofs_out, proof_len, status_out);
    API_END
}
int zkgpu_prove_batch_rng_dev(uint64_t pk, const void* d_advice, const uint64_t* instance, size_t num_instance, size_t m,
                              int rng_mode, void* rng_data, uint8_t* proofs_out, size_t proof_len, int32_t* status_out) {
    API_TRY
    prove_batch(pk, reinterpret_cast<const uint64_t*>(d_advice), true, instance, num_instance, m, rng_mode, rng_data, proofs_out, proof_len, status_out);
    API_END
}
// Test / parity form: fresh `SmallRng::seed_from_u64` per proof, all-or-nothing result
static int prove_batch_seeded(uint64_t pk, const uint64_t* advice, bool on_device, const uint64_t* instance, size_t num_instance, size_t m,
                              const uint64_t* rng_seeds, uint8_t* proofs_out, size_t proof_len) {
    API_TRY
    std::vector<int32_t> status(m, 0);
    prove_batch(pk, advice, on_device, instance, num_instance, m, RNG_SEED_U64, const_cast<uint64_t*>(rng_seeds), proofs_out, proof_len, status.data());
    for (size_t i = 0; i < m; ++i)
        if (status[i] == PROOF_LOOKUP_FAILED)
            throw Error(ZK_ERR_ARG, "create_proof: a lookup input is not in its table (ConstraintSystemFailure) in proof " + std::to_string(i) +
                                        "; the other proofs of the batch are valid (zkgpu_prove_batch_rng reports a status per proof)");
    API_END
}
int zkgpu_prove_batch(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, size_t m,
                      const uint64_t* rng_seeds, uint8_t* proofs_out, size_t proof_len) {
    return prove_batch_seeded(pk, advice, false, instance, num_instance, m, rng_seeds, proofs_out, proof_len);
}
int zkgpu_prove_batch_dev(uint64_t pk, const void* d_advice, const uint64_t* instance, size_t num_instance, size_t m,
                          const uint64_t* rng_seeds, uint8_t* proofs_out, size_t proof_len) {
    return prove_batch_seeded(pk, reinterpret_cast<const uint64_t*>(d_advice), true, instance, num_instance, m, rng_seeds, proofs_out, proof_len);
}

int zkgpu_prove(uint64_t pk, const uint64_t* advice, const uint64_t* instance, size_t num_instance, int rng_mode, void* rng_data,
                uint8_t* proof_out, size_t proof_len) {
    API_TRY
    rt().require();
    std::shared_ptr<PkShared> P = find_pk(pk);
    check_batch_args(*P->dev[0], advice, instance, num_instance, 1, rng_data, proof_out, proof_len, rng_mode);
    std::call_once(P->co_once, start_dispatchers, P.get());
    Coalescer& co = P->co;
    const PkEntry& pk0 = *P->dev[0];
    const size_t adv_bytes = (size_t)pk0.A * pk0.n * sizeof(fr_t);
    pk0.C->bind();
    void* slot = acquire_slot(co, adv_bytes);
    memcpy(slot, advice, adv_bytes);
    ProveReq req;
    req.advice = static_cast<const fr_t*>(slot); req.instance = reinterpret_cast<const fr_t*>(instance); req.num_pi = num_instance;
    req.rng_mode = rng_mode; req.rng_data = static_cast<uint8_t*>(rng_data); req.proof_out = proof_out;
    {
        std::unique_lock<std::mutex> lk(co.mu);
        if (co.stop) { lk.unlock(); release_slot(co, slot); throw Error(ZK_ERR_STATE, "proving key is being released"); }
        co.q.push_back(&req);
        co.cv_work.notify_all();
        co.cv_done.wait(lk, [&] { return req.done; });
    }
    release_slot(co, slot);
    if (req.rc != ZKGPU_OK) throw Error(req.rc, req.err);
    if (req.status == PROOF_LOOKUP_FAILED)
        throw Error(ZKGPU_ERR_WITNESS, "create_proof: a lookup input is not in its table (ConstraintSystemFailure)");
    API_END
}
int zkgpu_prove_stats(uint64_t pk, uint64_t out[4]) {
    API_TRY
    ZK_REQUIRE(out, "null pointer");
    std::shared_ptr<PkShared> P = find_pk(pk);
    std::lock_guard<std::mutex> lk(P->co.mu);
    out[0] = P->co.requests; out[1] = P->co.batches; out[2] = P->co.max_batch_seen; out[3] = P->co.threads.size();
    API_END
}

}  // extern "C"
