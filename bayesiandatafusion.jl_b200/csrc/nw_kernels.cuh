// Normal-Wishart hyper-parameter update on the device.
//   statistics  N, NU = Σu, NS = Σuuᵀ            — ConditionalNormalWishart, src/sampling.jl:116-119
//   posterior parameters + draw (mu, Lambda)      — src/sampling.jl:121-126, src/normal_wishart.jl:38-42
// The statistics are a streaming DMMA syrk over the rank's factor rows (same tile machinery as the row draw);
// the draw is a single-CTA kernel on D×D matrices.
#pragma once
#include "nw_device.cuh"
#include "row_kernel.cuh"

namespace bdf {

// Once per half-sweep: Λ in the tile order the row kernel factors in (identity on the padding), and Λ·μ for a shared μ.
__global__ void prep_lambda_kernel(const double* __restrict__ Lambda, const double* __restrict__ mu, int D, int DP, double* __restrict__ LT,
                                   double* __restrict__ lmu) {
  const int NB = DP / 8, NT = NB * (NB + 1) / 2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 64 * NT; e += gridDim.x * blockDim.x) {
    int I, J;
    tri_coords(e >> 6, I, J);
    const int i = 8 * I + ((e >> 3) & 7), j = 8 * J + (e & 7);
    LT[e] = (i < D && j < D) ? Lambda[i > j ? i + (size_t)j * D : j + (size_t)i * D] : (i == j ? 1.0 : 0.0);
  }
  if (mu)
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < DP; j += gridDim.x * blockDim.x) {
      double s = 0.0;
      if (j < D)
        for (int i = 0; i < D; i++) s = fma(Lambda[j + (size_t)i * D], mu[i], s);
      lmu[j] = s;
    }
}

// sum the partials in block order; emit [N, NU(D), NS(D×D, symmetric, column-major)]
__global__ void stats_reduce_kernel(const double* __restrict__ ws, int nblk, int D, double count, double* __restrict__ stats) {
  const int ne = tri(D + 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
    int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
    while (tri(i + 1) <= e) i++;
    while (tri(i) > e) i--;
    const int j = e - tri(i);
    if (j >= D) continue;  // (D, D) = Σ1·1 is not part of the statistics
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += ws[(size_t)b * ne + e];
    if (i < D) {
      stats[1 + D + i + (size_t)j * D] = s;
      stats[1 + D + j + (size_t)i * D] = s;
    } else {
      stats[1 + j] = s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] = count;
}

// ---- the draw -------------------------------------------------------------------------------------------------
struct NWDrawParams {
  int D;
  const double* stats;  // [N, NU, NS]
  const double* mu0;    // D
  const double* Tinv;   // D×D
  double b0, nu;
  const double* A_inj;  // injected Bartlett factor (D×D lower, col-major) or nullptr
  const double* z_inj;  // injected normals (D) or nullptr
  uint64_t seed, sweep;
  uint32_t stream;
  double* scratch;  // 4·D·D doubles
  double* mu_out;   // D
  double* Lam_out;  // D×D
  int* err_flag;
};

__global__ void __launch_bounds__(256) nw_draw_kernel(const NWDrawParams p) {
  const int D = p.D, tid = threadIdx.x, nt = blockDim.x;
  const size_t dd = (size_t)D * D;
  double* Sm = p.scratch;       // S reversed → L
  double* Y = p.scratch + dd;   // J·A → Y = L⁻ᵀ(J·A)
  double* Lm = p.scratch + 2 * dd;  // Lambda reversed → L2
  double* v = p.scratch + 3 * dd;   // vectors: mu_N [0,D), w [D,2D)
  const double N = p.stats[0];
  const double* NU = p.stats + 1;
  const double* NS = p.stats + 1 + D;
  const double betaN = p.b0 + N, nuN = p.nu + N;
  for (int i = tid; i < D; i += nt) v[i] = (p.b0 * p.mu0[i] + NU[i]) / (p.b0 + N);
  __syncthreads();
  // S = Tinv + NS + b0·mu0·mu0ᵀ − betaN·muN·muNᵀ (upper triangle is authoritative, src/sampling.jl:125), index-reversed
  for (size_t e = tid; e < dd; e += nt) {
    int i = (int)(e % D), j = (int)(e / D);
    if (i > j) { const int t = i; i = j; j = t; }
    const double s = p.Tinv[i + (size_t)j * D] + NS[i + (size_t)j * D] + p.b0 * p.mu0[i] * p.mu0[j] - betaN * v[i] * v[j];
    const int a = D - 1 - (int)(e % D), b = D - 1 - (int)(e / D);
    Sm[a + (size_t)b * D] = s;
  }
  // Bartlett factor A (lower): diag sqrt(chi2(nuN − i)), below-diagonal N(0,1); stored row-reversed: Y = J·A
  for (size_t e = tid; e < dd; e += nt) {
    const int i = (int)(e % D), j = (int)(e / D);
    double aij = 0.0;
    if (p.A_inj) {
      aij = i >= j ? p.A_inj[e] : 0.0;
    } else if (i == j) {
      aij = sqrt(2.0 * gamma_mt(0.5 * (nuN - i), p.seed, p.sweep, p.stream + 1, (uint64_t)i));
    } else if (i > j) {
      aij = philox_normal(p.seed, p.sweep, p.stream + 2, (uint64_t)i, j);
    }
    Y[(D - 1 - i) + (size_t)j * D] = aij;
  }
  bool ok = cta_chol_lower(Sm, D);  // J·S·J = L·Lᵀ  ⇒  chol_lower(inv(S)) = J·L⁻ᵀ·J
  // Y ← L⁻ᵀ·Y : thread per column, back substitution
  for (int c = tid; c < D; c += nt) {
    double* y = Y + (size_t)c * D;
    for (int i = D - 1; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < D; k++) s -= Sm[k + (size_t)i * D] * y[k];
      y[i] = s / Sm[i + (size_t)i * D];
    }
  }
  __syncthreads();
  // Z = J·Y ; Lambda = Z·Zᵀ ; store Lambda and its index-reversed copy
  for (size_t e = tid; e < dd; e += nt) {
    const int i = (int)(e % D), j = (int)(e / D);
    if (i < j) continue;
    double s = 0.0;
    for (int k = 0; k < D; k++) s = fma(Y[(D - 1 - i) + (size_t)k * D], Y[(D - 1 - j) + (size_t)k * D], s);
    p.Lam_out[i + (size_t)j * D] = s;
    p.Lam_out[j + (size_t)i * D] = s;
    Lm[(D - 1 - i) + (size_t)(D - 1 - j) * D] = s;
    Lm[(D - 1 - j) + (size_t)(D - 1 - i) * D] = s;
  }
  ok = cta_chol_lower(Lm, D) && ok;  // J·Lambda·J = L2·L2ᵀ ⇒ chol_lower(inv(Lambda)) = J·L2⁻ᵀ·J
  // w = L2⁻ᵀ (J z): column-oriented back substitution
  double* w = v + D;
  for (int i = tid; i < D; i += nt) {
    const int io = D - 1 - i;
    w[i] = p.z_inj ? p.z_inj[io] : philox_normal(p.seed, p.sweep, p.stream + 3, 0, io);
  }
  for (int i = D - 1; i >= 0; i--) {
    __syncthreads();
    const double wi = w[i] / Lm[i + (size_t)i * D];
    __syncthreads();
    if (tid == 0) w[i] = wi;
    for (int k = tid; k < i; k += nt) w[k] -= Lm[i + (size_t)k * D] * wi;
  }
  __syncthreads();
  const double sc = 1.0 / sqrt(betaN);
  for (int i = tid; i < D; i += nt) p.mu_out[i] = v[i] + sc * w[D - 1 - i];
  if (!ok && tid == 0) atomicOr(p.err_flag, 2);
}

// standard normals of the row-noise stream, D×N column-major (debug / parity hook)
__global__ void row_noise_kernel(double* out, int D, int64_t N, uint64_t seed, uint64_t sweep, uint32_t stream) {
  const int64_t n = N * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = philox_normal(seed, sweep, stream, (uint64_t)(e / D), (int)(e % D));
}

// ŷ_t = Σ_k Π_m U_m[slot(id_m(t))][k] + mean — udot/pred, src/sampling.jl:9-51
__global__ void predict_kernel(int K, const double* U0, const double* U1, const double* U2, const int32_t* s0, const int32_t* s1,
                               const int32_t* s2, int ld, int D, int64_t nt, double mean, double* out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x) {
    const double* a = U0 + (size_t)s0[t] * ld;
    const double* b = U1 + (size_t)s1[t] * ld;
    const double* c = K > 2 ? U2 + (size_t)s2[t] * ld : nullptr;
    double s = 0.0;
    for (int k = 0; k < D; k++) {
      double pr = a[k] * b[k];
      if (c) pr *= c[k];
      s += pr;
    }
    out[t] = s + mean;
  }
}

}  // namespace bdf
