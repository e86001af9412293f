// Correctness + throughput of the FP64-pipe Montgomery product (tools/dead_ends/fp_dfma.cuh) against the IMAD product.
//   correctness: N random pairs + edge values per field, bit-for-bit equality with fr_mul / fq_mul
//   throughput : dependent product chains — all warps IMAD, all warps DFMA, and a mix (DFMA on `mix` of 8 warps)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fp.cuh"
#include "dead_ends/fp_dfma.cuh"
using namespace zk;

template <class F>
__global__ void k_check(const F* a, const F* b, size_t n, unsigned long long* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = fe_load(a + i), y = fe_load(b + i);
    F r1 = x * y, r2 = mul_dfma(x, y);
    if (!(r1 == r2)) atomicAdd(bad, 1ull);
}

#define MUL_ITERS 256
template <class F, int MODE>   // MODE: 0 IMAD, 1 DFMA, 2..7: DFMA on warps with (warp % 8) < MODE
__global__ void __launch_bounds__(256) k_chain(uint32_t* out, uint32_t seed, int dfma_iters = MUL_ITERS) {
    F x[2], y;
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 8; ++i) x[c].l[i] = seed + threadIdx.x * 8 + i + c;
    for (int i = 0; i < 8; ++i) y.l[i] = seed * 7 + blockIdx.x + i;
    x[0].l[7] &= 0x0fffffff; x[1].l[7] &= 0x0fffffff; y.l[7] &= 0x0fffffff;
    const bool use_dfma = MODE == 1 || (MODE >= 2 && ((threadIdx.x >> 5) & 7) < MODE);
    if (use_dfma) {
        for (int it = 0; it < dfma_iters; ++it) { x[0] = mul_dfma(x[0], y); x[1] = mul_dfma(x[1], y); }
    } else {
        for (int it = 0; it < MUL_ITERS; ++it) { x[0] = x[0] * y; x[1] = x[1] * y; }
    }
    uint32_t r = 0;
    for (int c = 0; c < 2; ++c) for (int i = 0; i < 8; ++i) r ^= x[c].l[i];
    if (r == 0x12345679u) out[0] = r;
}

template <class L>
static double time_ms(L launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <class F>
static unsigned long long check(const uint32_t* mod, size_t n) {
    std::vector<F> a(n), b(n);
    auto reduce = [&](F& v) {   // bring below p: clear the top bits, then subtract p while >= p
        v.l[7] &= 0x3fffffffu;
        for (;;) {
            bool ge = true;
            for (int i = 7; i >= 0; --i) if (v.l[i] != mod[i]) { ge = v.l[i] > mod[i]; break; }
            if (!ge) break;
            uint64_t br = 0;
            for (int i = 0; i < 8; ++i) { uint64_t d = (uint64_t)v.l[i] - mod[i] - br; v.l[i] = (uint32_t)d; br = (d >> 63) & 1; }
        }
    };
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 8; ++k) { a[i].l[k] = (uint32_t)rnd(); b[i].l[k] = (uint32_t)rnd(); }
        if (i < 64) {   // edge values: 0, 1, p-1, all-ones halves, single limbs
            for (int k = 0; k < 8; ++k) { a[i].l[k] = (i & 1) ? mod[k] : 0; b[i].l[k] = (i & 2) ? mod[k] : 0xffffffffu * ((i >> 2) & 1); }
            if (i & 1) a[i].l[0] -= 1 + (i >> 3);
            if (i & 2) b[i].l[0] -= 1 + (i >> 4);
            if (i >= 32) { for (int k = 0; k < 8; ++k) a[i].l[k] = 0; a[i].l[(i - 32) % 8] = 0xffffffffu; }
        }
        reduce(a[i]); reduce(b[i]);
    }
    F *da, *db; unsigned long long* dbad; unsigned long long bad = 0;
    cudaMalloc(&da, n * sizeof(F)); cudaMalloc(&db, n * sizeof(F)); cudaMalloc(&dbad, 8);
    cudaMemcpy(da, a.data(), n * sizeof(F), cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * sizeof(F), cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 8);
    k_check<F><<<(unsigned)((n + 255) / 256), 256>>>(da, db, n, dbad);
    cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dbad);
    return bad;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t n = 1 << 22;
    unsigned long long bad_r = check<fr_t>(fr_params::MOD, n), bad_q = check<fq_t>(fq_params::MOD, n);
    uint32_t* d; cudaMalloc(&d, 4);
    int blocks = p.multiProcessorCount * 8, threads = 256;
    double muls = (double)blocks * threads * MUL_ITERS * 2;
    double t[8];
    t[0] = time_ms([&] { k_chain<fq_t, 0><<<blocks, threads>>>(d, 1); });
    t[1] = time_ms([&] { k_chain<fq_t, 1><<<blocks, threads>>>(d, 1); });
    t[2] = time_ms([&] { k_chain<fq_t, 2><<<blocks, threads>>>(d, 1); });
    t[3] = time_ms([&] { k_chain<fq_t, 3><<<blocks, threads>>>(d, 1); });
    t[4] = time_ms([&] { k_chain<fq_t, 4><<<blocks, threads>>>(d, 1); });
    // unequal work: DFMA warps run `di` iterations while IMAD warps run MUL_ITERS, so both kinds finish together
    for (int k = 2; k <= 4; ++k)
        for (int di = 64; di <= 224; di += 32) {
            double tm = 0;
            if (k == 2) tm = time_ms([&] { k_chain<fq_t, 2><<<blocks, threads>>>(d, 1, di); });
            if (k == 3) tm = time_ms([&] { k_chain<fq_t, 3><<<blocks, threads>>>(d, 1, di); });
            if (k == 4) tm = time_ms([&] { k_chain<fq_t, 4><<<blocks, threads>>>(d, 1, di); });
            double work = (double)blocks * threads / 8.0 * 2 * ((8 - k) * (double)MUL_ITERS + k * (double)di);
            printf("mix %d of 8 warps DFMA, %3d vs %d iterations: %.2f Gmul/s\n", k, di, MUL_ITERS, work / tm / 1e6);
        }
    printf("{\"gpu\": \"%s\", \"pairs_checked\": %zu, \"mismatch_fr\": %llu, \"mismatch_fq\": %llu, \"imad_Gmul_s\": %.2f, \"dfma_Gmul_s\": %.2f, "
           "\"mix_2of8_Gmul_s\": %.2f, \"mix_3of8_Gmul_s\": %.2f, \"mix_4of8_Gmul_s\": %.2f}\n",
           p.name, n, bad_r, bad_q, muls / t[0] / 1e6, muls / t[1] / 1e6, muls / t[2] / 1e6, muls / t[3] / 1e6, muls / t[4] / 1e6);
    return (bad_r || bad_q) ? 1 : 0;
}
