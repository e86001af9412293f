// Microbenchmark: integer-pipe issue peaks on this GPU — the denominator for the MSM
// "field-mul/s vs IMAD peak" roofline.  Measures, with all SMs busy and ILP=8 independent chains:
//   mad.lo.u32 (IMAD), mad.hi.u32 (IMAD.HI), mad.wide.u32 (IMAD.WIDE.U32), and the generated
//   254-bit Montgomery product (fr_mul, 128 IMAD.WIDE-class ops) as dependent chains per thread.
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../zkos-monorepo_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "fp.cuh"

#define ITERS 4096
template <int MODE>
__global__ void k_imad(uint32_t* out, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    unsigned long long w0 = a, w1 = a + 1, w2 = a + 2, w3 = a + 3, w4 = a + 4, w5 = a + 5, w6 = a + 6, w7 = a + 7;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {
            asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                         "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(b), "r"(a));
        } else if (MODE == 1) {
            asm volatile("mad.hi.u32 %0, %0, %8, %9; mad.hi.u32 %1, %1, %8, %9; mad.hi.u32 %2, %2, %8, %9; mad.hi.u32 %3, %3, %8, %9;"
                         "mad.hi.u32 %4, %4, %8, %9; mad.hi.u32 %5, %5, %8, %9; mad.hi.u32 %6, %6, %8, %9; mad.hi.u32 %7, %7, %8, %9;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(b), "r"(a));
        } else {
            // multiplicand = low half of the running accumulator, so the product cannot be hoisted
            asm volatile("mad.wide.u32 %0, %8, %16, %0; mad.wide.u32 %1, %9, %16, %1; mad.wide.u32 %2, %10, %16, %2; mad.wide.u32 %3, %11, %16, %3;"
                         "mad.wide.u32 %4, %12, %16, %4; mad.wide.u32 %5, %13, %16, %5; mad.wide.u32 %6, %14, %16, %6; mad.wide.u32 %7, %15, %16, %7;"
                         : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7)
                         : "r"((uint32_t)w0), "r"((uint32_t)w1), "r"((uint32_t)w2), "r"((uint32_t)w3), "r"((uint32_t)w4), "r"((uint32_t)w5), "r"((uint32_t)w6), "r"((uint32_t)w7), "r"(b));
        }
    }
    uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^ (uint32_t)((w0 ^ w7) >> 32);
    if (r == 0x12345679u) out[0] = r;
}

#define MUL_ITERS 512
template <int CHAINS>
__global__ void k_frmul(uint32_t* out, uint32_t seed) {
    zk::fr_t x[CHAINS], y;
    for (int c = 0; c < CHAINS; ++c)
        for (int i = 0; i < 8; ++i) x[c].l[i] = seed + threadIdx.x * 8 + i + c;
    for (int i = 0; i < 8; ++i) y.l[i] = seed * 7 + blockIdx.x + i;
    x[0].l[7] &= 0x0fffffff; y.l[7] &= 0x0fffffff;
    for (int it = 0; it < MUL_ITERS; ++it)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = x[c] * y;
    uint32_t r = 0;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 8; ++i) r ^= x[c].l[i];
    if (r == 0x12345679u) out[0] = r;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, 4);
    int blocks = sms * 8, threads = 256;
    double ops = (double)blocks * threads * ITERS * 8;
    double t0 = time_ms([&] { k_imad<0><<<blocks, threads>>>(d, 1); });
    double t1 = time_ms([&] { k_imad<1><<<blocks, threads>>>(d, 1); });
    double t2 = time_ms([&] { k_imad<2><<<blocks, threads>>>(d, 1); });
    double muls1 = (double)blocks * threads * MUL_ITERS * 1, muls2 = muls1 * 2;
    double m1 = time_ms([&] { k_frmul<1><<<blocks, threads>>>(d, 1); });
    double m2 = time_ms([&] { k_frmul<2><<<blocks, threads>>>(d, 1); });
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, \"imad_lo_Gops\": %.1f, \"imad_hi_Gops\": %.1f, \"imad_wide_Gops\": %.1f, "
           "\"fr_mul_G_per_s_1chain\": %.2f, \"fr_mul_G_per_s_2chain\": %.2f}\n",
           p.name, sms, clk, ops / t0 / 1e6, ops / t1 / 1e6, ops / t2 / 1e6, muls1 / m1 / 1e6, muls2 / m2 / 1e6);
    return 0;
}
