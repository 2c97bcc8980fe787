// Shared types and complex helpers for libtnqs_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;   // interleaved ComplexF64: .x = re, .y = im (binary compatible with Julia ComplexF64)

#define TNQS_MAX_LEGS 8          // site tensor rank <= 1 + 8

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { cplx r; r.x = re; r.y = im; return r; }
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cmake(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += a * conj(b)
__device__ __forceinline__ void cfma_cb(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y); acc.y = fma(-a.x, b.y, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfma_ca(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic (fixed-order) block sum; `red` holds >= blockDim.x/32 doubles; result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double s = 0;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}
