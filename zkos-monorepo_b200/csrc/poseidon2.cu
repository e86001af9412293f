// Poseidon2 over BN254 Fr, t = 8 (rate 7), x^7, 8 full + 48 partial rounds — the hash of Shielder's note tree and of
// `shielder_bindings::hash::poseidon_hash` — and the Merkle-path root of the withdraw / deposit witness.
//
// SURVEY.md §8f-4 (witness generation): the reference computes these on the host
//   /root/reference/crates/shielder_bindings/src/hash.rs:16-27       poseidon_hash(inputs) = hash_variable_length
//   /root/reference/crates/shielder_bindings/src/utils.rs:14-30      lengths 1..7, anything else panics
//   /root/reference/contracts/MerkleTree.sol:88-113,121-152          path layout (13 levels x 7 siblings), parent = hash(children)
// and on chain with the code /root/reference/poseidon2-solidity/generate_t8.py emits; the parameter tables below are dumped
// from that generator (gen_poseidon2.py) and the kernels are checked against its output (tests/test_poseidon2.py).
//
// One thread per hash: the 8-element state lives in registers (64 x u32), round constants and the internal diagonal in
// constant memory (3.8 KB, every lane reads the same word: broadcast).  Work per hash: 8 full rounds x 8 S-boxes x 4 products
// + 48 partial rounds x (4 + 8) products = 832 Montgomery products — IMAD-bound like the rest of the prover; a batch of
// 1024 withdraw witnesses needs 13 x 1024 hashes = 11 M products, microseconds next to the 64 G products of their proofs.
#include "api_util.hpp"
#include "poseidon2_consts.inc"

namespace zk {

__device__ __forceinline__ fr_t p2_const(const uint32_t (*tab)[8], unsigned i) {
    fr_t r;
#pragma unroll
    for (int l = 0; l < 8; ++l) r.l[l] = tab[i][l];
    return r;
}
__device__ __forceinline__ fr_t p2_pow7(const fr_t& x) { fr_t x2 = sqr(x), x4 = sqr(x2); return x4 * x2 * x; }
// generate_t8.py:480-498: (a, b, c, d) <- M4 (a, b, c, d), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
__device__ __forceinline__ void p2_mm4(fr_t& a, fr_t& b, fr_t& c, fr_t& d) {
    fr_t t0 = a + b, t1 = c + d, t2 = dbl(b) + t1, t3 = dbl(d) + t0;
    fr_t t4 = dbl(dbl(t1)) + t3, t5 = dbl(dbl(t0)) + t2;
    a = t3 + t5; b = t5; c = t2 + t4; d = t4;
}
// generate_t8.py:500-516: circ(2 M4, M4)
__device__ __forceinline__ void p2_external(fr_t s[8]) {
    p2_mm4(s[0], s[1], s[2], s[3]);
    p2_mm4(s[4], s[5], s[6], s[7]);
#pragma unroll
    for (int i = 0; i < 4; ++i) { fr_t u = s[i] + s[i + 4]; s[i] = s[i] + u; s[i + 4] = s[i + 4] + u; }
}
__device__ __forceinline__ void p2_permute(fr_t s[8]) {
    p2_external(s);
    for (unsigned r = 0; r < P2_RF / 2; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = p2_pow7(s[i] + p2_const(P2_RC_FULL, 8 * r + i));
        p2_external(s);
    }
    for (unsigned r = 0; r < P2_RP; ++r) {
        s[0] = p2_pow7(s[0] + p2_const(P2_RC_PART, r));
        fr_t sum = s[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) sum = sum + s[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = p2_const(P2_DIAG, i) * s[i] + sum;
    }
    for (unsigned r = P2_RF / 2; r < P2_RF; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = p2_pow7(s[i] + p2_const(P2_RC_FULL, 8 * r + i));
        p2_external(s);
    }
}
// capacity element len * 2^64 in Montgomery form
__device__ __forceinline__ fr_t p2_tag(unsigned len) {
    fr_t t = fr_t::zero();
    t.l[2] = len;
    return to_mont(t);
}

// out[i] = hash(in[i][0..len)), inputs at stride `stride` elements
__global__ void __launch_bounds__(128) k_poseidon2_hash(const fr_t* __restrict__ in, size_t stride, unsigned len, size_t m, fr_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    fr_t s[8];
#pragma unroll
    for (unsigned j = 0; j < 7; ++j) s[j] = j < len ? fe_load(in + i * stride + j) : fr_t::zero();
    s[7] = p2_tag(len);
    p2_permute(s);
    fe_store(out + i, s[0]);
}
// Merkle paths [m][height][7]: one hash per (path, level) was written to hashes[m][height]; the root is the top one and
// level l+1 must contain hashes[l]
__global__ void k_merkle_check(const fr_t* __restrict__ paths, const fr_t* __restrict__ hashes, unsigned height, size_t m,
                               fr_t* __restrict__ roots, uint8_t* __restrict__ consistent) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    bool ok = true;
    for (unsigned l = 0; l + 1 < height; ++l) {
        fr_t h = fe_load(hashes + i * height + l);
        bool found = false;
        for (unsigned j = 0; j < 7; ++j) found = found || (fe_load(paths + (i * height + l + 1) * 7 + j) == h);
        ok = ok && found;
    }
    fe_store(roots + i, fe_load(hashes + i * height + height - 1));
    consistent[i] = ok ? 1 : 0;
}

}  // namespace zk

using namespace zk;


extern "C" {

int zkgpu_poseidon2_hash_batch(const uint64_t* inputs, size_t len, size_t m, uint64_t* out) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(len >= 1 && len <= 7, "poseidon2: input length must be between 1 and 7 (POSEIDON_RATE)");
    ZK_REQUIRE(m == 0 || (inputs && out), "null pointer");
    if (m) {
        cudaStream_t st = C.stream;
        C.fr_buf.ensure(m * len + m);
        fr_t* d_in = C.fr_buf.p; fr_t* d_out = C.fr_buf.p + m * len;
        ZK_CUDA(cudaMemcpyAsync(d_in, inputs, m * len * 32, cudaMemcpyHostToDevice, st));
        ZK_LAUNCH(k_poseidon2_hash, ceil_div(m, 128), 128, 0, st, d_in, len, (unsigned)len, m, d_out);
        ZK_CUDA(cudaMemcpyAsync(out, d_out, m * 32, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
    }
    API_END
}

int zkgpu_poseidon2_hash_batch_dev(const void* d_inputs, size_t len, size_t m, void* d_out, void* stream) {
    API_TRY
    if (!m) return ZKGPU_OK;
    ZK_REQUIRE(d_inputs && d_out, "null pointer");
    DeviceScope scope(rt().of_pointer(d_inputs));
    Context& C = scope.C;
    ZK_REQUIRE(len >= 1 && len <= 7, "poseidon2: input length must be between 1 and 7 (POSEIDON_RATE)");
    ZK_REQUIRE(m == 0 || (d_inputs && d_out), "null pointer");
    cudaStream_t st = stream ? (cudaStream_t)stream : C.stream;
    if (m) ZK_LAUNCH(k_poseidon2_hash, ceil_div(m, 128), 128, 0, st, (const fr_t*)d_inputs, len, (unsigned)len, m, (fr_t*)d_out);
    if (!stream) ZK_CUDA(cudaStreamSynchronize(st));
    API_END
}

int zkgpu_merkle_root_batch(const uint64_t* paths, size_t height, size_t m, uint64_t* roots, uint8_t* consistent) {
    API_BEGIN
    Context& C = ctx(); C.require();
    ZK_REQUIRE(height >= 1 && height <= 64, "merkle: height must be between 1 and 64");
    ZK_REQUIRE(m == 0 || (paths && roots), "null pointer");
    if (m) {
        cudaStream_t st = C.stream;
        const size_t per = height * 7;
        C.fr_buf.ensure(m * per + m * height + m);
        fr_t* d_paths = C.fr_buf.p; fr_t* d_hash = d_paths + m * per; fr_t* d_roots = d_hash + m * height;
        DevBuf<uint8_t> d_ok(m);
        ZK_CUDA(cudaMemcpyAsync(d_paths, paths, m * per * 32, cudaMemcpyHostToDevice, st));
        ZK_LAUNCH(k_poseidon2_hash, ceil_div(m * height, 128), 128, 0, st, d_paths, (size_t)7, 7u, m * height, d_hash);
        ZK_LAUNCH(k_merkle_check, ceil_div(m, 128), 128, 0, st, d_paths, d_hash, (unsigned)height, m, d_roots, d_ok.p);
        ZK_CUDA(cudaMemcpyAsync(roots, d_roots, m * 32, cudaMemcpyDeviceToHost, st));
        std::vector<uint8_t> ok(m);
        ZK_CUDA(cudaMemcpyAsync(ok.data(), d_ok.p, m, cudaMemcpyDeviceToHost, st));
        ZK_CUDA(cudaStreamSynchronize(st));
        if (consistent) memcpy(consistent, ok.data(), m);
    }
    API_END
}

}  // extern "C"
