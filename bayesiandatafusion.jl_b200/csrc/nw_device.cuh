// Single-CTA dense helpers and samplers shared by the Normal-Wishart draw (nw_kernels.cuh) and the beta sampler (features.cu).
#pragma once
#include "philox.cuh"

namespace bdf {

// In-place lower Cholesky of a column-major n×n matrix (single CTA, blockDim.x a multiple of 16). Returns false on a non-positive
// pivot. Right-looking; the trailing update walks 16×(blockDim/16) patches of the lower triangle with a 2-D thread mapping (no
// integer division in the loop — the kernel that calls this sits on the critical path of every half-sweep).
__device__ inline bool cta_chol_lower(double* A, int n) {
  bool ok = true;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, nty = blockDim.x >> 4;
  for (int j = 0; j < n; j++) {
    __syncthreads();
    const double d = A[j + (size_t)j * n];
    if (!(d > 0.0)) ok = false;
    const double sd = sqrt(d);
    const double is = 1.0 / sd;
    __syncthreads();
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) A[i + (size_t)j * n] = (i == j) ? sd : A[i + (size_t)j * n] * is;
    __syncthreads();
    // trailing update, columns k > j, rows i >= k
    const double* cj = A + (size_t)j * n;
    for (int k = j + 1 + ty; k < n; k += nty) {
      const double akj = cj[k];
      double* ck = A + (size_t)k * n;
      for (int i0 = k + tx; i0 < n; i0 += 64) {  // four independent elements per pass: loads first, then the stores
        double a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + 16 * u;
          a[u] = i < n ? cj[i] : 0.0;
          c[u] = i < n ? ck[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + 16 * u;
          if (i < n) ck[i] = fma(-a[u], akj, c[u]);
        }
      }
    }
  }
  __syncthreads();
  return ok;
}

__device__ inline double gamma_mt(double a, uint64_t seed, uint64_t sweep, uint32_t stream, uint64_t row) {
  // Marsaglia–Tsang; Gamma(a, 1). For a < 1: Gamma(a+1)·U^(1/a).
  double boost = 1.0;
  uint32_t ctr = 0;
  if (a < 1.0) {
    double u1, u2;
    philox_uniform2(seed, sweep, stream, row, ctr++, u1, u2);
    boost = pow(u1, 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 64; it++) {
    double u1, u2, u3, u4;
    philox_uniform2(seed, sweep, stream, row, ctr++, u1, u2);
    philox_uniform2(seed, sweep, stream, row, ctr++, u3, u4);
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    const double x = sqrt(-2.0 * log(u1)) * co;
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u3) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
  }
  return boost * d;
}


}  // namespace bdf
