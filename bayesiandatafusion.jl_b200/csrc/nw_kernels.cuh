// Normal-Wishart hyper-parameter update on the device.
//   statistics  N, NU = Σu, NS = Σuuᵀ            — ConditionalNormalWishart, src/sampling.jl:116-119
//   posterior parameters + draw (mu, Lambda)      — src/sampling.jl:121-126, src/normal_wishart.jl:38-42
// The statistics are a streaming DMMA syrk over the rank's factor rows (same tile machinery as the row draw);
// the draw is a single-CTA kernel on D×D matrices.
#pragma once
#include <cstdio>
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
#pragma unroll 8
    for (int b = 0; b < nblk; b++) s += ws[(size_t)b * ne + e];  // block order: deterministic
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
  int debug;        // print phase clocks (BDF_DEBUG_NW)
  int nsm;          // how many D×D work matrices live in dynamic shared memory (2, 1 or 0); the rest use `scratch`
  double* mu_out;   // D
  double* Lam_out;  // D×D
  int* err_flag;
};

template <int NSM>  // D×D work matrices in dynamic shared memory: 2, 1 or 0 (compile-time, so that their accesses are LDS/STS, not generic)
__global__ void __launch_bounds__(256, NSM == 0 ? 4 : 1) nw_draw_kernel(const NWDrawParams p) {
  const int D = p.D, tid = threadIdx.x, nt = blockDim.x;
  const size_t dd = (size_t)D * D;
  // The two factorisations are dependent chains of D column steps; on-chip operands (≈30-cycle loads instead of an L2 round trip
  // per step) are what makes this single-CTA kernel short — it sits on the critical path of every half-sweep.
  extern __shared__ __align__(16) double nw_smem[];
  double* Sm = NSM >= 1 ? nw_smem : p.scratch;            // S reversed → L
  double* Y = NSM >= 2 ? nw_smem + dd : p.scratch + dd;   // J·A → Y = L⁻ᵀ(J·A)
  double* Lm = Sm;                  // Lambda reversed → L2; Sm is dead once Y has been formed
  double* v = p.scratch + 3 * dd;   // vectors: mu_N [0,D), w [D,2D)
  long long tk[10];
  int nk = 0;
#define NW_STAMP() do { __syncthreads(); if (p.debug && nk < 10) tk[nk++] = clock64(); } while (0)
  NW_STAMP();
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
      continue;  // the chi-square diagonal is drawn below, one thread per entry, so that no warp runs both samplers per element
    } else if (i > j) {
      aij = philox_normal(p.seed, p.sweep, p.stream + 2, (uint64_t)i, j);
    }
    Y[(D - 1 - i) + (size_t)j * D] = aij;
  }
  if (!p.A_inj)
    for (int i = tid; i < D; i += nt) Y[(D - 1 - i) + (size_t)i * D] = sqrt(2.0 * gamma_mt(0.5 * (nuN - i), p.seed, p.sweep, p.stream + 1, (uint64_t)i));
  NW_STAMP();
  bool ok = cta_chol_lower(Sm, D);  // J·S·J = L·Lᵀ  ⇒  chol_lower(inv(S)) = J·L⁻ᵀ·J
  NW_STAMP();
  // Y ← L⁻ᵀ·Y, right-looking: row i of Y is final once divided by L[i][i]; it is then eliminated from all rows k < i at once
  // (i·D independent updates per step over the whole CTA instead of one dependent chain per column)
  {
    const int tx = tid & 15, ty = tid >> 4, nty = nt >> 4;
    for (int i = D - 1; i >= 0; i--) {
      const double inv = 1.0 / Sm[i + (size_t)i * D];
      for (int c = tid; c < D; c += nt) Y[i + (size_t)c * D] *= inv;
      __syncthreads();
      for (int c = ty; c < D; c += nty) {
        double* yc = Y + (size_t)c * D;
        const double yic = yc[i];
        for (int k0 = tx; k0 < i; k0 += 64) {  // four independent elements per pass
          double l[4], y[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = k0 + 16 * u;
            l[u] = k < i ? Sm[i + (size_t)k * D] : 0.0;
            y[u] = k < i ? yc[k] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = k0 + 16 * u;
            if (k < i) yc[k] = fma(-l[u], yic, y[u]);
          }
        }
      }
      __syncthreads();
    }
  }
  NW_STAMP();
  // Z = J·Y ; Lambda = Z·Zᵀ ; store Lambda and its index-reversed copy
  for (size_t e = tid; e < dd; e += nt) {
    const int i = (int)(e % D), j = (int)(e / D);
    if (i < j) continue;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four partial sums: the dot product is not one dependent chain
    const double* yi = Y + (D - 1 - i);
    const double* yj = Y + (D - 1 - j);
    int k = 0;
    for (; k + 4 <= D; k += 4) {
      s0 = fma(yi[(size_t)k * D], yj[(size_t)k * D], s0);
      s1 = fma(yi[(size_t)(k + 1) * D], yj[(size_t)(k + 1) * D], s1);
      s2 = fma(yi[(size_t)(k + 2) * D], yj[(size_t)(k + 2) * D], s2);
      s3 = fma(yi[(size_t)(k + 3) * D], yj[(size_t)(k + 3) * D], s3);
    }
    for (; k < D; k++) s0 = fma(yi[(size_t)k * D], yj[(size_t)k * D], s0);
    const double s = (s0 + s1) + (s2 + s3);
    p.Lam_out[i + (size_t)j * D] = s;
    p.Lam_out[j + (size_t)i * D] = s;
    Lm[(D - 1 - i) + (size_t)(D - 1 - j) * D] = s;
    Lm[(D - 1 - j) + (size_t)(D - 1 - i) * D] = s;
  }
  NW_STAMP();
  ok = cta_chol_lower(Lm, D) && ok;  // J·Lambda·J = L2·L2ᵀ ⇒ chol_lower(inv(Lambda)) = J·L2⁻ᵀ·J
  NW_STAMP();
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
  NW_STAMP();
  if (p.debug && tid == 0)
    printf("nw_draw D=%d cycles: build+noise %lld, chol1 %lld, backsub %lld, ZZt %lld, chol2 %lld, mu %lld, total %lld\n", D, tk[1] - tk[0], tk[2] - tk[1],
           tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[6] - tk[0]);
#undef NW_STAMP
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
