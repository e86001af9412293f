// Internal interface of the G1 FFT module (g1fft.cu).
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace zk {
void g1_fft(g1_xyzz_t* d_a, unsigned log_n, const fr_t& omega, cudaStream_t st);        // in place, natural order
void g1_scale(g1_xyzz_t* d_a, size_t n, const fr_t& scalar_mont, cudaStream_t st);      // a[i] *= scalar
void g1_from_affine(const g1_affine_t* d_in, g1_xyzz_t* d_out, size_t n, cudaStream_t st);
}  // namespace zk
