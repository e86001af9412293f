// Microbenchmark: FP64 DFMA issue rate on this GPU, alone and co-issued with 64-bit integer adds (the op mix
// of a floating-point-assisted Montgomery product).  Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k_dfma(double* out, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0 + seed * 1e-12;
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    long long s0 = threadIdx.x, s1 = 1, s2 = 2, s3 = 3;
    for (int i = 0; i < ITERS; ++i) {
        x0 = __fma_rz(x0, b, a); x1 = __fma_rz(x1, b, a); x2 = __fma_rz(x2, b, a); x3 = __fma_rz(x3, b, a);
        x4 = __fma_rz(x4, b, a); x5 = __fma_rz(x5, b, a); x6 = __fma_rz(x6, b, a); x7 = __fma_rz(x7, b, a);
        if (MODE == 1) {  // one 64-bit integer add of the raw bits per DFMA
            s0 += __double_as_longlong(x0); s1 += __double_as_longlong(x1); s2 += __double_as_longlong(x2); s3 += __double_as_longlong(x3);
            s0 += __double_as_longlong(x4); s1 += __double_as_longlong(x5); s2 += __double_as_longlong(x6); s3 += __double_as_longlong(x7);
        }
    }
    double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (double)(s0 ^ s1 ^ s2 ^ s3);
    if (r == 1.2345) out[0] = r;
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
    double* d; cudaMalloc(&d, 8);
    int blocks = p.multiProcessorCount * 8, threads = 256;
    double ops = (double)blocks * threads * ITERS * 8;
    double t0 = time_ms([&] { k_dfma<0><<<blocks, threads>>>(d, 1.5); });
    double t1 = time_ms([&] { k_dfma<1><<<blocks, threads>>>(d, 1.5); });
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_Gops\": %.1f, \"dfma_plus_iadd64_Gops\": %.1f}\n", p.name, p.multiProcessorCount,
           ops / t0 / 1e6, ops / t1 / 1e6);
    return 0;
}
