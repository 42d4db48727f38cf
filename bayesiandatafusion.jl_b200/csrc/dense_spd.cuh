// Dense symmetric-positive-definite solve on the device: (FF + λI) \ rhs of solve_full (src/sampling.jl:314-320) and the relation-feature
// solve (α·FᵀF + λI) \ rhs (src/sampling.jl:322-334). The reference calls LAPACK's LU through `\`; the matrix is SPD, so this is a
// blocked right-looking Cholesky A = L·Lᵀ (block 64) followed by blocked forward / backward substitutions on all right-hand sides at once.
//
//   A   n × n, column-major, leading dimension n; only the LOWER triangle is read and it is overwritten with L;
//   B   n × nrhs right-hand sides, ROW-major with pitch ld (row i = the nrhs values of unknown i — the layout of the β matrix on the
//       device); overwritten with the solution;
//   info  device int: 0, or 1 + the first column whose pivot was not positive.
//
// Per 64-column block: potf2 (one CTA, the 64×64 diagonal block in shared memory) → trsm_panel (one thread per row below the block, the
// row in registers, L_kk broadcast from shared memory) → syrk_update (64×64 tiles of the trailing lower triangle, 4×4 register tiles,
// the two 64×16 panel slices staged in shared memory). The substitutions follow the same pattern with one thread per right-hand side in
// the diagonal step and 8×4 register tiles in the update. Plain FP64 FMAs throughout: numF ≤ compute_ff_size = 6500 (≤ 0.1 TFLOP per
// factorisation) — this solve is outside the bandwidth path, it only has to be exact and not slow.
//
// The kernels use nothing but threadIdx / blockIdx / __shared__ / __syncthreads, and every launch goes through BDF_LAUNCH, so that
// tests/cpu_shim can compile this very file for the host (threads = std::thread, __syncthreads = a barrier) and check it against LAPACK
// without a GPU.
#pragma once
#include <cstdint>

#ifndef BDF_LAUNCH
#define BDF_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

namespace bdf {
namespace spd {

constexpr int NB = 64;    // block size of the factorisation and of the substitutions
constexpr int KC = 16;    // depth of one staged slice in the update kernels
constexpr int RC = 128;   // right-hand sides per CTA in the substitution kernels

// L_kk (lower triangle of the diagonal block at k0, bs = min(NB, n − k0) rows) → shared memory, padded with the identity
template <int NT>
__device__ __forceinline__ void load_diag_block(const double* __restrict__ A, int64_t n, int64_t k0, double (*L)[NB + 1]) {
  const int bs = (int)(n - k0 < NB ? n - k0 : NB);
  for (int e = threadIdx.x; e < NB * NB; e += NT) {
    const int i = e % NB, j = e / NB;
    double v = i == j ? 1.0 : 0.0;
    if (i < bs && j <= i) v = A[(size_t)(k0 + i) + (size_t)(k0 + j) * n];
    L[i][j] = v;
  }
}

// Cholesky of the diagonal block: right-looking, one column per step
static __global__ void __launch_bounds__(256) potf2_kernel(double* __restrict__ A, int64_t n, int64_t k0, int* __restrict__ info) {
  __shared__ double T[NB][NB + 1];
  const int bs = (int)(n - k0 < NB ? n - k0 : NB);
  load_diag_block<256>(A, n, k0, T);
  for (int j = 0; j < NB; j++) {
    __syncthreads();
    const double ajj = T[j][j];
    if (threadIdx.x == 0 && !(ajj > 0.0)) atomicCAS(info, 0, (int)(k0 + j + 1));
    const double d = sqrt(ajj);
    __syncthreads();  // every thread has read the pivot
    for (int i = j + threadIdx.x; i < NB; i += 256) T[i][j] = i == j ? d : T[i][j] / d;
    __syncthreads();
    const int m = NB - 1 - j;
    for (int e = threadIdx.x; e < m * m; e += 256) {
      const int i = j + 1 + e % m, l = j + 1 + e / m;
      if (i >= l) T[i][l] -= T[i][j] * T[l][j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int i = e % NB, j = e / NB;
    if (i < bs && j <= i) A[(size_t)(k0 + i) + (size_t)(k0 + j) * n] = T[i][j];
  }
}

// rows below the diagonal block: A[i, k0:k0+NB] ← A[i, k0:k0+NB] · L_kk⁻ᵀ, one thread per row (blockIdx.x = row block below the diagonal one)
static __global__ void __launch_bounds__(NB) trsm_panel_kernel(double* __restrict__ A, int64_t n, int64_t k0) {
  __shared__ double L[NB][NB + 1];
  load_diag_block<NB>(A, n, k0, L);
  __syncthreads();
  const int64_t row = k0 + (int64_t)NB * (1 + blockIdx.x) + threadIdx.x;
  if (row >= n) return;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) x[j] = A[(size_t)row + (size_t)(k0 + j) * n];
#pragma unroll
  for (int j = 0; j < NB; j++) {
    double s = x[j];
#pragma unroll
    for (int l = 0; l < j; l++) s -= x[l] * L[j][l];
    x[j] = s / L[j][j];
  }
#pragma unroll
  for (int j = 0; j < NB; j++) A[(size_t)row + (size_t)(k0 + j) * n] = x[j];
}

// trailing update: tile (I, J), J ≤ I, of the lower triangle below / right of block k:  A_IJ −= P_I · P_Jᵀ  (P = the panel just scaled)
static __global__ void __launch_bounds__(256) syrk_update_kernel(double* __restrict__ A, int64_t n, int64_t k0) {
  if (blockIdx.y > blockIdx.x) return;
  __shared__ double As[KC][NB];
  __shared__ double Bs[KC][NB];
  const int64_t i0 = k0 + (int64_t)NB * (1 + blockIdx.x), j0 = k0 + (int64_t)NB * (1 + blockIdx.y);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
  for (int kc = 0; kc < NB; kc += KC) {
    __syncthreads();
    for (int e = threadIdx.x; e < KC * NB; e += 256) {
      const int i = e % NB, kk = e / NB;
      const size_t col = (size_t)(k0 + kc + kk) * n;
      As[kk][i] = i0 + i < n ? A[(size_t)(i0 + i) + col] : 0.0;
      Bs[kk][i] = j0 + i < n ? A[(size_t)(j0 + i) + col] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; kk++) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; r++) a[r] = As[kk][tx + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; c++) b[c] = Bs[kk][ty + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] += a[r] * b[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int64_t i = i0 + tx + 16 * r, j = j0 + ty + 16 * c;
      if (i < n && j <= i) A[(size_t)i + (size_t)j * n] -= acc[r][c];
    }
}

// diagonal step of the substitutions, one thread per right-hand side: FWD solves L_kk·y = b, else L_kkᵀ·x = y
template <bool FWD>
__global__ void __launch_bounds__(RC) trsv_block_kernel(const double* __restrict__ A, int64_t n, int64_t k0, double* __restrict__ B, int ld, int nrhs) {
  __shared__ double L[NB][NB + 1];
  load_diag_block<RC>(A, n, k0, L);
  __syncthreads();
  const int c = blockIdx.x * RC + threadIdx.x;
  if (c >= nrhs) return;
  const int bs = (int)(n - k0 < NB ? n - k0 : NB);
  double y[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) y[j] = j < bs ? B[(size_t)(k0 + j) * ld + c] : 0.0;
  if (FWD) {
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double s = y[j];
#pragma unroll
      for (int l = 0; l < j; l++) s -= L[j][l] * y[l];
      y[j] = s / L[j][j];
    }
  } else {
#pragma unroll
    for (int j = NB - 1; j >= 0; j--) {
      double s = y[j];
#pragma unroll
      for (int l = j + 1; l < NB; l++) s -= L[l][j] * y[l];
      y[j] = s / L[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < NB; j++)
    if (j < bs) B[(size_t)(k0 + j) * ld + c] = y[j];
}

// update step of the substitutions, 64 rows × RC right-hand sides per CTA (blockIdx.y = group of right-hand sides):
//   FWD: rows of the blocks BELOW block k:  B[i, :] −= Σ_j A[i, k0+j]·B[k0+j, :]        (blockIdx.x = row block below the diagonal one)
//   else: rows of the blocks ABOVE block k: B[i, :] −= Σ_j A[k0+j, i]·B[k0+j, :]        (blockIdx.x = row block, from the top)
template <bool FWD>
__global__ void __launch_bounds__(256) subst_update_kernel(const double* __restrict__ A, int64_t n, int64_t k0, double* __restrict__ B, int ld, int nrhs) {
  __shared__ double Ls[KC][NB];
  __shared__ double Ys[KC][RC];
  const int64_t i0 = FWD ? k0 + (int64_t)NB * (1 + blockIdx.x) : (int64_t)NB * blockIdx.x;
  const int c0 = blockIdx.y * RC;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  double acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int q = 0; q < 4; q++) acc[r][q] = 0.0;
  for (int kc = 0; kc < NB; kc += KC) {
    __syncthreads();
    for (int e = threadIdx.x; e < KC * NB; e += 256) {
      int i, kk;
      if (FWD) { i = e % NB; kk = e / NB; } else { kk = e % KC; i = e / KC; }
      const int64_t kr = k0 + kc + kk, ir = i0 + i;
      double v = 0.0;
      if (kr < n && ir < n) v = FWD ? A[(size_t)ir + (size_t)kr * n] : A[(size_t)kr + (size_t)ir * n];
      Ls[kk][i] = v;
    }
    for (int e = threadIdx.x; e < KC * RC; e += 256) {
      const int c = e % RC, kk = e / RC;
      const int64_t kr = k0 + kc + kk;
      Ys[kk][c] = kr < n && c0 + c < nrhs ? B[(size_t)kr * ld + c0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; kk++) {
      double a[8], b[4];
#pragma unroll
      for (int r = 0; r < 8; r++) a[r] = Ls[kk][ty + 8 * r];
#pragma unroll
      for (int q = 0; q < 4; q++) b[q] = Ys[kk][tx + 32 * q];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[r][q] += a[r] * b[q];
    }
  }
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int64_t i = i0 + ty + 8 * r;
      const int c = c0 + tx + 32 * q;
      if (i < n && c < nrhs) B[(size_t)i * ld + c] -= acc[r][q];
    }
}

// A ← L (lower triangle), B ← A⁻¹·B. Everything is enqueued on `stream`; *info (device) must be 0 on entry. Returns the number of launches.
template <class Stream>
inline int solve(Stream stream, double* A, int64_t n, double* B, int ld, int nrhs, int* info) {
  if (n <= 0 || nrhs <= 0) return 0;
  const int nb = (int)((n + NB - 1) / NB);
  const int cg = (nrhs + RC - 1) / RC;
  int launches = 0;
  for (int k = 0; k < nb; k++) {
    const int64_t k0 = (int64_t)k * NB;
    const int below = nb - 1 - k;
    BDF_LAUNCH(potf2_kernel, dim3(1), dim3(256), stream, A, n, k0, info);
    launches++;
    if (below > 0) {
      BDF_LAUNCH(trsm_panel_kernel, dim3(below), dim3(NB), stream, A, n, k0);
      BDF_LAUNCH(syrk_update_kernel, dim3(below, below), dim3(256), stream, A, n, k0);
      launches += 2;
    }
  }
  for (int k = 0; k < nb; k++) {
    const int64_t k0 = (int64_t)k * NB;
    const int below = nb - 1 - k;
    BDF_LAUNCH(trsv_block_kernel<true>, dim3(cg), dim3(RC), stream, A, n, k0, B, ld, nrhs);
    launches++;
    if (below > 0) {
      BDF_LAUNCH(subst_update_kernel<true>, dim3(below, cg), dim3(256), stream, A, n, k0, B, ld, nrhs);
      launches++;
    }
  }
  for (int k = nb - 1; k >= 0; k--) {
    const int64_t k0 = (int64_t)k * NB;
    BDF_LAUNCH(trsv_block_kernel<false>, dim3(cg), dim3(RC), stream, A, n, k0, B, ld, nrhs);
    launches++;
    if (k > 0) {
      BDF_LAUNCH(subst_update_kernel<false>, dim3(k, cg), dim3(256), stream, A, n, k0, B, ld, nrhs);
      launches++;
    }
  }
  return launches;
}

}  // namespace spd
}  // namespace bdf
