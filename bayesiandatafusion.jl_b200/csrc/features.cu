// Macau link-matrix (beta) path on the device — SURVEY §8a A8–A15:
//   sparse-binary feature matrix F as CSR + CSC index lists  (SparseBinMatrix src/parallel_matrix.jl:9-24,
//                                                              SparseBinMatrixCSR src/sparsebin_csr.jl:6-37)
//   Y = F·X, Y = Fᵀ·X for D right-hand sides at once          (A_mul_B! / At_mul_B! :242-267, sparsebin_csr.jl:49-63)
//   (FᵀF + λI)·X                                               (AtA_mul_B!, src/parallel_cg.jl:7-14)
//   batched conjugate gradients with per-column scalars/masks   (cg_AtA src/parallel_cg.jl:63-94, solve_cg2 parallel_matrix.jl:488-507)
//   sample_beta / update_beta! / sample_lambda_beta             (src/sampling.jl:291-312, 361-370, 136-142)
//   F_mul_beta → uhat, mu .+ uhat                               (src/RelationData.jl:314-320, src/macau.jl:102-104)
// A 0/1 matrix needs no multiplies: both products are gathers of D-vectors summed in stored order (the reference's
// order), one warp per output row, indices read 128 bits at a time, the D columns spread over the lanes.
#include <cub/cub.cuh>
#include <cublas_v2.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/bdf_b200.h"
#include "engine.cuh"
#include "dense_spd.cuh"
#include "nw_device.cuh"
#include "row_kernel.cuh"

using namespace bdf;

int bdf_stats_of(bdf_t* h, const double* X, const double* sub, int64_t slot0, int64_t nrows, double* stats);
int bdf_check_err_flag(bdf_t* h);
int bdf_copy_rows_h2d_impl(bdf_t* h, int entity, const double* host, double* dev);
int bdf_relation_residuals(bdf_t* h, int rel);
int bdf_refresh_relation_offsets(bdf_t* h, int rel);
int bdf_sample_entity_impl(bdf_t* h, int entity, const double* mu_dev, int64_t mu_ld, const double* Lambda_dev, const double* Z_dev);
int bdf_join_side(bdf_t* h);

namespace {

inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- Y[r,:] = Σ_{e in row r} X[idx[e],:] (+ lam·P[r,:])  — one warp per work item, lane l owns columns l, l+32, … -----------
// A work item is a whole row (≤ SPLIT indices: summed strictly in stored order = the reference's order, bit-exact) or one
// SPLIT-index chunk of a longer row (popular feature bits reach 10⁵ entries); chunk partials are added in chunk order by
// spbin_reduce_kernel, so the result is deterministic. Indices are fetched 32 at a time (one coalesced 128-byte load per
// warp), broadcast by shuffle, and 8 gathers are kept in flight before their in-order accumulation.
struct SpItem {
  int32_t row;    // output row
  int32_t slot;   // -1: write Y[row] directly; else partial slot in the workspace
  int64_t beg, end;
};

// VAL: general sparse F with stored values (SparseMatrixCSC features, e.g. test/parallel_latent_basic.jl:4): y += a·x with a
// separate multiply and add, like Julia's un-contracted `tmp += nzval[j]*x[...]`, in stored order.
template <int NC, bool VAL>
__global__ void __launch_bounds__(256) spbin_gather_kernel(const SpItem* __restrict__ items, int n_items, const int32_t* __restrict__ idx,
                                                           const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                                                           double* __restrict__ part, int ld, double lam, const double* __restrict__ P) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int it = warp0; it < n_items; it += nwarps) {
    const SpItem w = items[it];
    double acc[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) acc[k] = 0.0;
    for (int64_t o = w.beg; o < w.end; o += 32) {
      const int n = (int)min((int64_t)32, w.end - o);
      const int mine = lane < n ? __ldg(idx + o + lane) : 0;
      const double myval = (VAL && lane < n) ? __ldg(vals + o + lane) : 0.0;
      for (int j0 = 0; j0 < n; j0 += 8) {
        double v[8][NC];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = __shfl_sync(0xffffffffu, mine, (j0 + u) & 31);
          const double a = VAL ? __shfl_sync(0xffffffffu, myval, (j0 + u) & 31) : 1.0;
          const double* src = X + (size_t)c * ld + lane;
#pragma unroll
          for (int k = 0; k < NC; k++) {
            const double x = (j0 + u < n && lane + 32 * k < ld) ? __ldg(src + 32 * k) : 0.0;
            v[u][k] = VAL ? __dmul_rn(a, x) : x;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int k = 0; k < NC; k++) acc[k] = __dadd_rn(acc[k], v[u][k]);  // +0.0 for the padded slots leaves the sum bit-identical
      }
    }
    if (w.slot < 0) {
      double* dst = Y + (size_t)w.row * ld + lane;
#pragma unroll
      for (int k = 0; k < NC; k++)
        if (lane + 32 * k < ld) {
          double s = acc[k];
          if (P) s += lam * P[(size_t)w.row * ld + lane + 32 * k];
          dst[32 * k] = s;
        }
    } else {
      double* dst = part + (size_t)w.slot * ld + lane;
#pragma unroll
      for (int k = 0; k < NC; k++)
        if (lane + 32 * k < ld) dst[32 * k] = acc[k];
    }
  }
}

// The same product with as few instructions per index as the hardware allows — the kernel above is issue-bound (ncu: 69 % of the issue slots
// busy, ≈20 warp instructions per index, L2 at 21–38 % of its throughput). Here LPR lanes cover one gathered row with 128-bit loads (two columns per
// lane) and the warp handles 32/LPR indices per step: a load instruction always moves 512 bytes, one shuffle serves 32/LPR indices. The 32/LPR
// interleaved partial sums are combined by a shuffle tree at the end, so the summation order is deterministic but NOT the strictly sequential
// order of the reference — this kernel is used where that order is unobservable (inside the CG iteration, whose iterates are chaotic anyway,
// and on the packed column windows of a column-split solve); bdf_spmm, F·beta and the rhs of sample_beta keep the in-order kernel.
template <int LPR, int NC2, bool VAL>
__global__ void __launch_bounds__(256) spbin_gather_fast_kernel(const SpItem* __restrict__ items, int n_items, const int32_t* __restrict__ idx,
                                                                const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                                                                double* __restrict__ part, int ld, double lam, const double* __restrict__ P) {
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, c = lane % LPR;
  bool ok[NC2];
#pragma unroll
  for (int k = 0; k < NC2; k++) ok[k] = 2 * (c + LPR * k) < ld;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int it = warp0; it < n_items; it += nwarps) {
    const SpItem w = items[it];
    double2 acc[NC2];
#pragma unroll
    for (int k = 0; k < NC2; k++) acc[k] = make_double2(0.0, 0.0);
    for (int64_t o = w.beg; o < w.end; o += 32) {
      const int n = (int)min((int64_t)32, w.end - o);
      const int mine = lane < n ? __ldg(idx + o + lane) : 0;
      const double myval = (VAL && lane < n) ? __ldg(vals + o + lane) : 0.0;
      constexpr int U = G >= 8 ? 2 : (G >= 4 ? 4 : 8);  // index groups in flight (16 gathered rows per warp)
      for (int j0 = 0; j0 < n; j0 += U * G) {
        double2 v[U][NC2];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int j = j0 + u * G + g;
          const int ci = __shfl_sync(0xffffffffu, mine, j & 31);
          const double a = VAL ? __shfl_sync(0xffffffffu, myval, j & 31) : 1.0;
          const double2* src = reinterpret_cast<const double2*>(X + (size_t)ci * ld) + c;
#pragma unroll
          for (int k = 0; k < NC2; k++) {
            double2 x = (j < n && ok[k]) ? __ldg(src + LPR * k) : make_double2(0.0, 0.0);
            if (VAL) { x.x = __dmul_rn(a, x.x); x.y = __dmul_rn(a, x.y); }
            v[u][k] = x;
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int k = 0; k < NC2; k++) { acc[k].x = __dadd_rn(acc[k].x, v[u][k].x); acc[k].y = __dadd_rn(acc[k].y, v[u][k].y); }
      }
    }
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
      for (int k = 0; k < NC2; k++) { acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, off); acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, off); }
    if (g == 0) {
#pragma unroll
      for (int k = 0; k < NC2; k++)
        if (ok[k]) {
          const int col = 2 * (c + LPR * k);
          if (w.slot < 0) {
            double2 r = acc[k];
            if (P) { const double2 pv = *reinterpret_cast<const double2*>(P + (size_t)w.row * ld + col); r.x += lam * pv.x; r.y += lam * pv.y; }
            *reinterpret_cast<double2*>(Y + (size_t)w.row * ld + col) = r;
          } else {
            *reinterpret_cast<double2*>(part + (size_t)w.slot * ld + col) = acc[k];
          }
        }
    }
  }
}

// rows that were split: Y[row,:] = Σ_chunks partial (chunk order) (+ lam·P[row,:])
struct SpLong {
  int32_t row, nchunks;
  int64_t slot0;
};
__global__ void spbin_reduce_kernel(const SpLong* __restrict__ rows, int nlong, const double* __restrict__ part, double* __restrict__ Y, int ld,
                                    double lam, const double* __restrict__ P) {
  for (int r = blockIdx.x; r < nlong; r += gridDim.x) {
    const SpLong w = rows[r];
    for (int d = threadIdx.x; d < ld; d += blockDim.x) {
      double s = 0.0;
      for (int c = 0; c < w.nchunks; c++) s += part[(size_t)(w.slot0 + c) * ld + d];
      if (P) s += lam * P[(size_t)w.row * ld + d];
      Y[(size_t)w.row * ld + d] = s;
    }
  }
}

// ---- column-wise dot products over a (rows × ld) pair, deterministic two-stage reduction ----------------------------
__global__ void __launch_bounds__(256) coldot_partial_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t rows, int ld,
                                                             int D, double* __restrict__ part) {
  // thread (ty, d): d = threadIdx.x % 32 … covers columns d, d+32, d+64, d+96; ty = threadIdx.x / 32 strides rows
  const int d0 = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t r = (int64_t)blockIdx.x * 8 + ty; r < rows; r += (int64_t)gridDim.x * 8) {
    const double* a = A + (size_t)r * ld;
    const double* b = B + (size_t)r * ld;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int d = d0 + 32 * k;
      if (d < D) s[k] = fma(a[d], b[d], s[k]);
    }
  }
  __shared__ double sh[8][128];
#pragma unroll
  for (int k = 0; k < 4; k++) sh[ty][d0 + 32 * k] = s[k];
  __syncthreads();
  if (threadIdx.x < 128) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < 8; y++) t += sh[y][threadIdx.x];
    part[(size_t)blockIdx.x * 128 + threadIdx.x] = t;
  }
}

// CG scalars for every column — one block. mode 0: bknum = r·r, convergence test, bk, bkden (src/parallel_cg.jl:74-84);
// mode 1: ak = bknum / (z·p) (:89).
struct CGState {
  double* bknum;   // [128]
  double* bkden;   // [128]
  double* coef;    // [128]  bk (mode 0) or ak (mode 1); 0 for inactive columns
  double* tolv;    // [128]  tol·‖b‖ per column
  int* active;     // [128]  column still iterating
  int* iters;      // [128]  operator applications so far
  int* nactive;    // [1]
};

__global__ void cg_scalars_kernel(const double* __restrict__ part, int nblk, int D, int mode, int iter, CGState st) {
  const int d = threadIdx.x;
  if (d < D) {
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += part[(size_t)b * 128 + d];
    if (mode == 2) {  // ‖b‖ → tolerance (src/parallel_cg.jl:65)
      st.tolv[d] = st.tolv[d] * sqrt(s);
    } else if (mode == 0) {
      double bk = 0.0;
      if (st.active[d]) {
        if (sqrt(s) < st.tolv[d]) {
          st.active[d] = 0;  // err < tol && return x
        } else {
          if (iter > 1) bk = s / st.bkden[d];
          st.bkden[d] = s;
          st.bknum[d] = s;
          st.iters[d] += 1;
        }
      }
      st.coef[d] = bk;
    } else {
      st.coef[d] = st.active[d] ? st.bknum[d] / s : 0.0;
    }
  }
  __syncthreads();
  if (mode == 0 && threadIdx.x == 0) {
    int n = 0;
    for (int k = 0; k < D; k++) n += st.active[k];
    *st.nactive = n;
  }
}

// coldot_partial_kernel + cg_scalars_kernel in one launch: the block that finishes last (arrival counter) adds the partials up in block order —
// the same fixed order as the two-kernel form — and updates the CG scalars. Two launches less per CG iteration.
__global__ void __launch_bounds__(256) coldot_scalars_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t rows, int ld, int D,
                                                             double* __restrict__ part, int* __restrict__ counter, int mode, int iter, CGState st) {
  const int d0 = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t r = (int64_t)blockIdx.x * 8 + ty; r < rows; r += (int64_t)gridDim.x * 8) {
    const double* a = A + (size_t)r * ld;
    const double* b = B + (size_t)r * ld;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int d = d0 + 32 * k;
      if (d < D) s[k] = fma(a[d], b[d], s[k]);
    }
  }
  __shared__ double sh[8][128];
  __shared__ int s_last;
#pragma unroll
  for (int k = 0; k < 4; k++) sh[ty][d0 + 32 * k] = s[k];
  __syncthreads();
  if (threadIdx.x < 128) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < 8; y++) t += sh[y][threadIdx.x];
    part[(size_t)blockIdx.x * 128 + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(counter, 1);
    s_last = old == (int)gridDim.x - 1;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int d = threadIdx.x;
  if (d < D) {
    double t = 0.0;
    for (int b = 0; b < (int)gridDim.x; b++) t += __ldcg(part + (size_t)b * 128 + d);
    if (mode == 2) {
      st.tolv[d] = st.tolv[d] * sqrt(t);
    } else if (mode == 0) {
      double bk = 0.0;
      if (st.active[d]) {
        if (sqrt(t) < st.tolv[d]) {
          st.active[d] = 0;
        } else {
          if (iter > 1) bk = t / st.bkden[d];
          st.bkden[d] = t;
          st.bknum[d] = t;
          st.iters[d] += 1;
        }
      }
      st.coef[d] = bk;
    } else {
      st.coef[d] = st.active[d] ? st.bknum[d] / t : 0.0;
    }
  }
  __syncthreads();
  if (mode == 0 && threadIdx.x == 0) {
    int n = 0;
    for (int k = 0; k < D; k++) n += st.active[k];
    *st.nactive = n;
  }
}

// p = bk·p + r on the active columns (prod_add!, src/parallel_cg.jl:28-32); iter 1 keeps p = r
__global__ void cg_update_p_kernel(double* __restrict__ P, const double* __restrict__ R, int64_t rows, int ld, int D, const double* __restrict__ bk,
                                   const int* __restrict__ active, int iter) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(e % ld);
    if (d < D && active[d] && iter > 1) P[e] = bk[d] * P[e] + R[e];
  }
}

// x += ak·p ; r −= ak·z on the active columns (add_prod!/sub_prod!, src/parallel_cg.jl:34-46)
__global__ void cg_update_xr_kernel(double* __restrict__ X, double* __restrict__ R, const double* __restrict__ P, const double* __restrict__ Z,
                                    int64_t rows, int ld, int D, const double* __restrict__ ak, const int* __restrict__ active) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(e % ld);
    if (d < D && active[d]) {
      const double a = ak[d];
      X[e] += a * P[e];
      R[e] -= a * Z[e];
    }
  }
}

// column-major host layout (rows × ncol) ↔ row-major device layout (rows × ld)
__global__ void to_rowmajor_kernel(const double* __restrict__ cm, int64_t rows, int ncol, int ld, double* __restrict__ rm) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ld;
    const int d = (int)(e % ld);
    rm[e] = d < ncol ? cm[r + (size_t)d * rows] : 0.0;
  }
}
__global__ void to_colmajor_kernel(const double* __restrict__ rm, int64_t rows, int ncol, int ld, double* __restrict__ cm) {
  const int64_t n = rows * ncol;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % rows;
    const int d = (int)(e / rows);
    cm[e] = rm[(size_t)r * ld + d];
  }
}

// column window [c0, c0+nc) of a (rows × ld) matrix → packed (rows × ldw) buffer, zero padding
__global__ void window_extract_kernel(const double* __restrict__ A, int64_t rows, int ld, int c0, int nc, int ldw, double* __restrict__ W) {
  const int64_t n = rows * ldw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldw;
    const int c = (int)(e % ldw);
    W[e] = c < nc ? A[(size_t)r * ld + c0 + c] : 0.0;
  }
}
// … and back: the solved window goes into columns [c0, c0+nc) of this rank's beta AND of every peer's replica (IPC-mapped peer memory, plain
// NVLink stores) — the all-gather of the beta columns of solve_cg2's column split (src/parallel_matrix.jl:488-507) without a collective
struct PeerPtrs { double* p[8]; };
__global__ void window_scatter_kernel(const double* __restrict__ W, int64_t rows, int ld, int c0, int nc, int ldw, double* __restrict__ A, PeerPtrs peers) {
  const int64_t n = rows * nc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / nc;
    const int c = (int)(e % nc);
    const double v = W[(size_t)r * ldw + c];
    const size_t o = (size_t)r * ld + c0 + c;
    A[o] = v;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (peers.p[q]) peers.p[q][o] = v;
  }
}

// C = chol_lower(inv(Λ)) (the colouring matrix of MvNormal(0, inv(PDMat(Λ))), src/sampling.jl:298): J·Λ·J = L·Lᵀ ⇒
// C = J·L⁻ᵀ·J. Single CTA; scratch = 2·D·D doubles. Output C column-major.
__global__ void __launch_bounds__(256) color_matrix_kernel(const double* __restrict__ Lambda, int D, double* scratch, double* __restrict__ Cout,
                                                           int* err_flag) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double* Lm = scratch;
  double* Y = scratch + (size_t)D * D;
  for (int e = tid; e < D * D; e += nt) {
    int i = e % D, j = e / D;
    const int a = D - 1 - i, b = D - 1 - j;
    if (i > j) { const int t = i; i = j; j = t; }
    Lm[a + (size_t)b * D] = Lambda[i + (size_t)j * D];  // Symmetric(Λ) reads the upper triangle
    Y[e] = (e % D == e / D) ? 1.0 : 0.0;
  }
  const bool ok = cta_chol_lower(Lm, D);
  // Y = L⁻ᵀ (columns of the identity, back substitution), thread per column
  for (int c = tid; c < D; c += nt) {
    double* y = Y + (size_t)c * D;
    for (int i = D - 1; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < D; k++) s -= Lm[k + (size_t)i * D] * y[k];
      y[i] = s / Lm[i + (size_t)i * D];
    }
  }
  __syncthreads();
  for (int e = tid; e < D * D; e += nt) {
    const int i = e % D, j = e / D;
    Cout[e] = Y[(D - 1 - i) + (size_t)(D - 1 - j) * D];
  }
  if (!ok && tid == 0) atomicOr(err_flag, 4);
}

// T[r,:] = (U[r,:] − mu) + C·e_r   (sample_u_c' + rand(mv,N)', src/sampling.jl:300) or scale·C·e_r when U == nullptr;
// e_r: injected standard normals (row-major rows × ld) or the Philox stream `stream`.
__global__ void __launch_bounds__(128) colored_rows_kernel(const double* __restrict__ U, const double* __restrict__ mu, const double* __restrict__ Cm,
                                                           const double* __restrict__ E, int64_t rows, int ld, int D, double scale,
                                                           uint64_t seed, uint64_t sweep, uint32_t stream, double* __restrict__ T, int accumulate,
                                                           int world = 1, int64_t nper = 0, const int32_t* __restrict__ slot_tab = nullptr) {
  extern __shared__ double sh[];  // C (D×D col-major) + per-warp noise vectors
  double* Cs = sh;
  double* es = sh + (size_t)D * D;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) Cs[e] = Cm[e];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* ev = es + (size_t)w * D;
  for (int64_t r = (int64_t)blockIdx.x * nw + w; r < rows; r += (int64_t)gridDim.x * nw) {
    for (int k = lane; k < D; k += 32) ev[k] = E ? E[(size_t)r * ld + k] : philox_normal(seed, sweep, stream, (uint64_t)r, k);
    __syncwarp();
    for (int i = lane; i < ld; i += 32) {
      double s = 0.0;
      if (i < D) {
        for (int k = 0; k <= i; k++) s = fma(Cs[i + (size_t)k * D], ev[k], s);  // C lower triangular
        s *= scale;
        if (U) {
          // U is slot-major (the rank's replica of the whole factor matrix); T, E and the noise key are in row order
          const int64_t us = slot_tab ? slot_tab[r] : (world > 1 ? (r % world) * nper + r / world : r);
          s += U[(size_t)us * ld + i] - mu[i];
        }
        if (accumulate) s += T[(size_t)r * ld + i];
      }
      T[(size_t)r * ld + i] = s;
    }
    __syncwarp();
  }
}

// lambda_beta ~ Gamma(shape νx/2, scale 2μx/νx), νx = ν + numF·D, μx = μ·νx/(ν + μ·tr((βᵀβ)Λ)) — src/sampling.jl:136-142
__global__ void lambda_beta_kernel(const double* __restrict__ btb_stats, const double* __restrict__ Lambda, int D, double numF, double nu, double mu,
                                   double g_inj, uint64_t seed, uint64_t sweep, uint32_t stream, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double* BtB = btb_stats + 1 + D;
    double tr = 0.0;
    for (int i = 0; i < D; i++)
      for (int k = 0; k < D; k++) tr += BtB[i + (size_t)k * D] * Lambda[k + (size_t)i * D];
    const double nux = nu + numF * D;
    const double mux = mu * nux / (nu + mu * tr);
    const double b = nux / 2.0, c = 2.0 * mux / nux;
    const double g = g_inj == g_inj ? g_inj : gamma_mt(b, seed, sweep, stream, 0);
    out[0] = c * g;
    out[1] = b;
  }
}

__global__ void add_mu_kernel(const double* __restrict__ uhat, const double* __restrict__ mu, int64_t rows, int ld, int D, double* __restrict__ out) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(e % ld);
    out[e] = d < D ? uhat[e] + mu[d] : 0.0;
  }
}

// the same through the row → slot map of a sharded entity: uhat (row order, from the feature product) → uhat and mu .+ uhat in slot order
__global__ void add_mu_scatter_kernel(const double* __restrict__ uhat_rows, const double* __restrict__ mu, int64_t rows, int ld, int D, int world, int64_t nper,
                                      const int32_t* __restrict__ slot_tab, double* __restrict__ uhat_slots, double* __restrict__ out) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ld;
    const int d = (int)(e % ld);
    const int64_t us = slot_tab ? slot_tab[r] : (r % world) * nper + r / world;
    const double v = uhat_rows[e];
    uhat_slots[(size_t)us * ld + d] = d < D ? v : 0.0;
    out[(size_t)us * ld + d] = d < D ? v + mu[d] : 0.0;
  }
}

__global__ void sort_keys_kernel(const int32_t* a, int64_t n, int32_t maxv, uint32_t* keys, uint32_t* idx, int* bad) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = a[o];
    if (v < 1 || v > maxv) { *bad = 1; keys[o] = 0; } else keys[o] = (uint32_t)(v - 1);
    idx[o] = (uint32_t)o;
  }
}
__global__ void count_keys_kernel(const uint32_t* keys, int64_t n, unsigned long long* counts) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) atomicAdd(counts + keys[o], 1ULL);
}
__global__ void gather_other_kernel(const uint32_t* perm, int64_t n, const int32_t* other, int32_t* out) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) out[o] = other[perm[o]] - 1;
}

__global__ void gather_fval_kernel(const uint32_t* perm, int64_t n, const double* v, double* out) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) out[o] = v[perm[o]];
}
__global__ void csc_expand_kernel(const int64_t* colptr, const int64_t* rowval, int64_t ncols, int64_t nnz, int32_t* rows, int32_t* cols, int* bad) {
  // one thread per column: COO (row, col) pairs in CSC order (1-based in, 1-based out)
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = colptr[c] - 1, e = colptr[c + 1] - 1;
    if (b < 0 || e < b || e > nnz) { *bad = 1; continue; }
    for (int64_t o = b; o < e; o++) { rows[o] = (int32_t)rowval[o]; cols[o] = (int32_t)(c + 1); }
  }
}

template <class T>
int dalloc(bdf_t* h, T** p, size_t n) {
  *p = nullptr;
  CU(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return BDF_OK;
}

// one orientation of F: stable sort of the COO list by `key` (rows for CSR, cols for CSC) — the reference's
// sortperm(rows) in SparseBinMatrixCSR (src/sparsebin_csr.jl:23) — then the pointer array by counting.
int build_orientation(bdf_t* h, const int32_t* d_key, const int32_t* d_other, int64_t nnz, int64_t nkeys, int32_t other_max, int64_t** ptr_out,
                      int32_t** ind_out, const double* d_val = nullptr, double** val_out = nullptr) {
  uint32_t *keys = nullptr, *keys2 = nullptr, *idx = nullptr, *idx2 = nullptr;
  int* d_bad = nullptr; unsigned long long* cnt = nullptr; void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(keys); cudaFree(keys2); cudaFree(idx); cudaFree(idx2); cudaFree(d_bad); cudaFree(cnt); cudaFree(tmp); };
#define TRYC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return BDF_ERR_CUDA; } } while (0)
  const size_t n1 = std::max<int64_t>(nnz, 1);
  TRYC(cudaMalloc((void**)&keys, n1 * 4)); TRYC(cudaMalloc((void**)&keys2, n1 * 4)); TRYC(cudaMalloc((void**)&idx, n1 * 4)); TRYC(cudaMalloc((void**)&idx2, n1 * 4));
  TRYC(cudaMalloc((void**)&d_bad, 4)); TRYC(cudaMalloc((void**)&cnt, (nkeys + 1) * 8));
  TRYC(cudaMemsetAsync(d_bad, 0, 4, h->stream));
  TRYC(cudaMemsetAsync(cnt, 0, (nkeys + 1) * 8, h->stream));
  sort_keys_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(d_key, nnz, (int32_t)nkeys, keys, idx, d_bad);
  // `other` range check rides on the same flag
  sort_keys_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(d_other, nnz, other_max, keys2, idx2, d_bad);
  int bits = 1;
  while ((1LL << bits) < nkeys) bits++;
  size_t tb = 0, sb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, idx2, (int)n1, 0, bits, h->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, sb, cnt, (int64_t*)nullptr, (int)(nkeys + 1), h->stream);
  tb = std::max(tb, sb);
  TRYC(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
  sort_keys_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(d_key, nnz, (int32_t)nkeys, keys, idx, d_bad);
  if (nnz) TRYC(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, idx2, (int)nnz, 0, bits, h->stream));
  count_keys_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(keys2, nnz, cnt);
  int rc;
  if ((rc = dalloc(h, ptr_out, (size_t)nkeys + 1))) { cleanup(); return rc; }
  if ((rc = dalloc(h, ind_out, n1))) { cleanup(); return rc; }
  TRYC(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, *ptr_out, (int)(nkeys + 1), h->stream));
  gather_other_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(idx2, nnz, d_other, *ind_out);
  if (d_val && val_out) {
    if ((rc = dalloc(h, val_out, n1))) { cleanup(); return rc; }
    gather_fval_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(idx2, nnz, d_val, *val_out);
  }
  int bad = 0;
  TRYC(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream));
  TRYC(cudaStreamSynchronize(h->stream));
  cleanup();
#undef TRYC
  if (bad) FAIL(BDF_ERR_INVALID, "feature row/column index outside 1..m / 1..n");
  return BDF_OK;
}

// ---- dense feature matrices and the FF direct solve (A9: solve_full, src/sampling.jl:314-320; FF = full(FᵀF), src/RelationData.jl:337-339)
// Plain library calls: cuBLAS dgemm for F·X / Fᵀ·X / FᵀF with a dense F, cuSOLVER potrf/potrs for (FF + λI)·β = rhs.
int dense_handles(bdf_t* h) {
  if (!h->cublas) {
    cublasHandle_t cb = nullptr;
    if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) FAIL(BDF_ERR_CUDA, "cublasCreate failed");
    h->cublas = cb;
  }
  cublasSetStream((cublasHandle_t)h->cublas, h->stream);
  return BDF_OK;
}

__global__ void identity_block_kernel(double* __restrict__ X, int64_t rows, int ld, int64_t c0, int ncol) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * ld; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ld;
    const int c = (int)(e % ld);
    X[e] = (c < ncol && r == c0 + c) ? 1.0 : 0.0;
  }
}
// FF[:, c0 + c] = Y[:, c] for c < ncol (Y row-major rows × ld, FF column-major rows × rows)
__global__ void scatter_ff_kernel(const double* __restrict__ Y, int64_t rows, int ld, int64_t c0, int ncol, double* __restrict__ FF) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * ncol; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % rows;
    const int c = (int)(e / rows);
    FF[r + (size_t)(c0 + c) * rows] = Y[(size_t)r * ld + c];
  }
}
__global__ void add_diag_kernel(double* __restrict__ A, int64_t n, double lam) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) A[i + (size_t)i * n] += lam;
}

// Y = F·X (transpose == false, X numF × ld → Y N × ld) or Y = Fᵀ·X (+ lam·P) for a dense column-major F; all operands row-major with
// pitch ld, i.e. column-major ld × rows, so both products are one dgemm on the transposed problem.
int dense_mm(bdf_t* h, const EntityS& e, bool transpose, const double* X, double* Y, double lam, const double* P) {
  int rc = dense_handles(h);
  if (rc) return rc;
  cublasHandle_t cb = (cublasHandle_t)h->cublas;
  const double one = 1.0, zero = 0.0;
  const int ld = h->ld;
  cublasStatus_t st;
  if (!transpose) {
    // Yᵀ (ld × N) = Xᵀ (ld × numF) · Fᵀ (numF × N)
    st = cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, ld, (int)e.N, (int)e.numF, &one, X, ld, e.f_dense, (int)e.N, &zero, Y, ld);
  } else {
    if (P && lam != 0.0) CU(cudaMemcpyAsync(Y, P, sizeof(double) * (size_t)e.numF * ld, cudaMemcpyDeviceToDevice, h->stream));
    const double beta = (P && lam != 0.0) ? lam : 0.0;
    // Yᵀ (ld × numF) = Xᵀ (ld × N) · F (N × numF)
    st = cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, ld, (int)e.numF, (int)e.N, &one, X, ld, e.f_dense, (int)e.N, &beta, Y, ld);
  }
  h->launches++;
  if (st != CUBLAS_STATUS_SUCCESS) FAIL(BDF_ERR_CUDA, "cublasDgemm failed");
  return BDF_OK;
}

int need_features(bdf_t* h, int entity) {
  if (h->ents[entity].numF <= 0) FAIL(BDF_ERR_STATE, "entity has no feature matrix (bdf_set_features_sbm / _csc / _dense)");
  return BDF_OK;
}

// `ldx` = pitch of X / Y / P in doubles (0 = the handle's): a column window of the right-hand sides (the rank's share of a sharded CG,
// src/parallel_matrix.jl:488-507) is a packed buffer with a smaller pitch, so a gathered row is only as wide as the window
int spmm(bdf_t* h, const EntityS& e, bool transpose, const double* X, double* Y, double lam = 0.0, const double* P = nullptr, int ldx = 0, bool exact = true) {
  const int ld = ldx > 0 ? ldx : h->ld;
  if (e.f_dense) {
    if (ld != h->ld) FAIL(BDF_ERR_STATE, "column windows are not used with dense feature matrices");
    return dense_mm(h, e, transpose, X, Y, lam, P);
  }
  const int o = transpose ? 1 : 0;
  const SpItem* items = reinterpret_cast<const SpItem*>(e.sp_items[o]);
  const int ni = e.sp_nitems[o];
  int64_t g = ((int64_t)ni * 32 + 255) / 256;
  g = std::min<int64_t>(std::max<int64_t>(g, 1), 148 * 8);
  const int32_t* idx = transpose ? e.f_rowind : e.f_colind;
  const int nc = (ld + 31) / 32;
  double* part = reinterpret_cast<double*>(e.sp_part);
  const double* vals = transpose ? e.f_val_csc : e.f_val_csr;
  if (!exact) {  // order-free variant (CG iterations, column windows): fewest instructions per index
#define SPF(LPR_, NC2_)                                                                                                                          \
  if (vals) spbin_gather_fast_kernel<LPR_, NC2_, true><<<(int)g, 256, 0, h->stream>>>(items, ni, idx, vals, X, Y, part, ld, lam, P);           \
  else spbin_gather_fast_kernel<LPR_, NC2_, false><<<(int)g, 256, 0, h->stream>>>(items, ni, idx, nullptr, X, Y, part, ld, lam, P);
    if (ld <= 4) { SPF(2, 1) } else if (ld <= 8) { SPF(4, 1) } else if (ld <= 16) { SPF(8, 1) } else if (ld <= 32) { SPF(16, 1) } else if (ld <= 64) { SPF(32, 1) } else { SPF(32, 2) }
#undef SPF
    h->launches++;
    if (e.sp_nlong[o] > 0) {
      spbin_reduce_kernel<<<std::min(e.sp_nlong[o], 148 * 4), 128, 0, h->stream>>>(reinterpret_cast<const SpLong*>(e.sp_long[o]), e.sp_nlong[o], part, Y, ld, lam, P);
      h->launches++;
    }
    CU(cudaGetLastError());
    return BDF_OK;
  }
#define SPL(NC_)                                                                                                                   \
  if (vals) spbin_gather_kernel<NC_, true><<<(int)g, 256, 0, h->stream>>>(items, ni, idx, vals, X, Y, part, ld, lam, P);         \
  else spbin_gather_kernel<NC_, false><<<(int)g, 256, 0, h->stream>>>(items, ni, idx, nullptr, X, Y, part, ld, lam, P);
  switch (nc) {
    case 1: SPL(1) break;
    case 2: SPL(2) break;
    case 3: SPL(3) break;
    default: SPL(4) break;
  }
#undef SPL
  h->launches++;
  if (e.sp_nlong[o] > 0) {
    spbin_reduce_kernel<<<std::min(e.sp_nlong[o], 148 * 4), 128, 0, h->stream>>>(reinterpret_cast<const SpLong*>(e.sp_long[o]), e.sp_nlong[o], part, Y, ld, lam, P);
    h->launches++;
  }
  CU(cudaGetLastError());
  return BDF_OK;
}

// work list of one orientation: rows with more than SPLIT indices are cut into SPLIT-index chunks
int build_sp_items(bdf_t* h, EntityS& e, int o, const int64_t* d_ptr, int64_t nrows, int64_t* part_slots) {
  const int64_t SPLIT = 1024;
  std::vector<int64_t> ptr((size_t)nrows + 1);
  CU(cudaMemcpyAsync(ptr.data(), d_ptr, 8 * ptr.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<SpItem> items;
  std::vector<SpLong> longs;
  int64_t slots = 0;
  items.reserve((size_t)nrows);
  for (int64_t r = 0; r < nrows; r++) {
    const int64_t b = ptr[r], n = ptr[r + 1] - b;
    if (n <= SPLIT) {
      items.push_back({(int32_t)r, -1, b, b + n});
    } else {
      const int64_t nch = (n + SPLIT - 1) / SPLIT;
      longs.push_back({(int32_t)r, (int32_t)nch, slots});
      for (int64_t c = 0; c < nch; c++) items.push_back({(int32_t)r, (int32_t)(slots + c), b + c * SPLIT, std::min(b + n, b + (c + 1) * SPLIT)});
      slots += nch;
    }
  }
  std::stable_sort(items.begin(), items.end(), [](const SpItem& a, const SpItem& b) { return (a.end - a.beg) > (b.end - b.beg); });
  e.sp_nitems[o] = (int)items.size();
  e.sp_nlong[o] = (int)longs.size();
  CU(cudaMalloc(&e.sp_items[o], std::max<size_t>(items.size(), 1) * sizeof(SpItem)));
  CU(cudaMalloc(&e.sp_long[o], std::max<size_t>(longs.size(), 1) * sizeof(SpLong)));
  if (!items.empty()) CU(cudaMemcpyAsync(e.sp_items[o], items.data(), items.size() * sizeof(SpItem), cudaMemcpyHostToDevice, h->stream));
  if (!longs.empty()) CU(cudaMemcpyAsync(e.sp_long[o], longs.data(), longs.size() * sizeof(SpLong), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (slots > *part_slots) *part_slots = slots;
  return BDF_OK;
}

// batched CG on (FᵀF + λI)X = B, all D columns at once; B, X device row-major (numF × ld). Returns per-column iteration counts.
// Dc / ldc: the number of right-hand sides in B / X and their pitch (0 = all num_latent columns at the handle's pitch)
int cg_solve_dev(bdf_t* h, EntityS& e, const double* B, double* X, double lambda, double tol, int64_t maxiter, int* iters_host, int Dc = 0, int ldc = 0) {
  const int D = Dc > 0 ? Dc : h->D, ld = ldc > 0 ? ldc : h->ld;
  const int64_t n = e.numF, m = e.N;
  const size_t vn = (size_t)n * ld, vm = (size_t)m * ld;
  const int NBLK = 296;
  if (!e.cgbuf) {  // sized for the full width; a column window uses a prefix of each vector
    int rc = dalloc(h, &e.cgbuf, 3 * (size_t)n * h->ld + (size_t)m * h->ld + (size_t)NBLK * 128 + 4 * 128 + 4 * 128 + 8);
    if (rc) return rc;
    CU(cudaMemsetAsync(e.cgbuf + 3 * (size_t)n * h->ld + (size_t)m * h->ld + (size_t)NBLK * 128 + 8 * 128, 0, 64, h->stream));  // arrival counter
  }
  double* R = e.cgbuf; double* P = R + (size_t)n * h->ld; double* Z = P + (size_t)n * h->ld; double* T = Z + (size_t)n * h->ld; double* part = T + (size_t)m * h->ld;
  CGState st;
  st.bknum = part + (size_t)NBLK * 128; st.bkden = st.bknum + 128; st.coef = st.bkden + 128; st.tolv = st.coef + 128;
  st.active = reinterpret_cast<int*>(st.tolv + 128); st.iters = st.active + 128; st.nactive = st.iters + 128;
  int* counter = reinterpret_cast<int*>(part + (size_t)NBLK * 128 + 8 * 128);  // self-resetting arrival counter of coldot_scalars_kernel
  std::vector<double> tolh(128, tol);
  std::vector<int> acth(128, 0), zero(128, 0);
  for (int d = 0; d < D; d++) acth[d] = 1;
  CU(cudaMemcpyAsync(st.tolv, tolh.data(), 128 * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(st.active, acth.data(), 128 * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(st.iters, zero.data(), 128 * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(X, 0, vn * 8, h->stream));                                        // x = 0
  CU(cudaMemcpyAsync(R, B, vn * 8, cudaMemcpyDeviceToDevice, h->stream));              // r = b
  CU(cudaMemcpyAsync(P, B, vn * 8, cudaMemcpyDeviceToDevice, h->stream));              // p = r
  coldot_scalars_kernel<<<NBLK, 256, 0, h->stream>>>(B, B, n, ld, D, part, counter, 2, 0, st);   // tol ← tol·‖b‖
  h->launches += 1;
  int nact = D;
  // one CG iteration (src/parallel_cg.jl:73-92); `iter` only matters as iter == 1 (p = r) versus iter > 1
  auto body = [&](int64_t iter) -> int {
    coldot_scalars_kernel<<<NBLK, 256, 0, h->stream>>>(R, R, n, ld, D, part, counter, 0, (int)std::min<int64_t>(iter, 2), st);
    cg_update_p_kernel<<<grid_for(vn), 256, 0, h->stream>>>(P, R, n, ld, D, st.coef, st.active, (int)std::min<int64_t>(iter, 2));
    int rc = spmm(h, e, false, P, T, 0.0, nullptr, ld, false);    // T = F·P       (order-free gather: see spbin_gather_fast_kernel)
    if (!rc) rc = spmm(h, e, true, T, Z, lambda, P, ld, false);   // Z = Fᵀ·T + λ·P
    coldot_scalars_kernel<<<NBLK, 256, 0, h->stream>>>(Z, P, n, ld, D, part, counter, 1, (int)std::min<int64_t>(iter, 2), st);
    cg_update_xr_kernel<<<grid_for(vn), 256, 0, h->stream>>>(X, R, P, Z, n, ld, D, st.coef, st.active);
    h->launches += 4;
    return rc;
  };
  auto check = [&]() -> int {  // the all-converged test costs a drain of the stream
    CU(cudaMemcpyAsync(&nact, st.nactive, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BDF_OK;
  };
  int64_t iter = 1;
  int rcb;
  if (maxiter >= 1) { if ((rcb = body(1))) return rcb; iter = 2; }
  // The iteration is launch-bound (ten short kernels, ≈1000 iterations per draw): iterations 2… are captured ONCE into a CUDA graph of KB
  // iterations and replayed; converged columns are masked on the device, so the few iterations that run past convergence inside a
  // replay change nothing (neither the solution nor the per-column iteration counts).
  constexpr int KB = 8;
  if (maxiter - iter + 1 >= 2 * KB) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const int64_t l0 = h->launches;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rcc = BDF_OK;
    for (int k = 0; k < KB && !rcc; k++) rcc = body(2);
    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    const int64_t per_replay = h->launches - l0;
    h->launches = l0;
    if (rcc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (rcc) return rcc; FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce)); }
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) { cudaGraphDestroy(graph); FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce)); }
    while (nact > 0 && maxiter - iter + 1 >= KB) {
      ce = cudaGraphLaunch(exec, h->stream);
      if (ce != cudaSuccess) break;
      h->launches += per_replay;
      iter += KB;
      if ((rcb = check())) { cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); return rcb; }
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  }
  for (; iter <= maxiter && nact > 0; iter++) {
    if ((rcb = body(iter))) return rcb;
    if ((iter & 3) == 0 || iter == maxiter) { if ((rcb = check())) return rcb; }
  }
  CU(cudaGetLastError());
  if (iters_host) {
    std::vector<int> it(128);
    CU(cudaMemcpyAsync(it.data(), st.iters, 128 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(iters_host, it.data(), sizeof(int) * D);
  }
  return BDF_OK;
}

int upload_rowmajor(bdf_t* h, const double* host_cm, int64_t rows, int ncol, double* dev_rm) {
  int rc = bdf_ensure_arena2(h, sizeof(double) * (size_t)rows * ncol);  // staging in the second arena: no cudaMalloc per call
  if (rc) return rc;
  double* stage = reinterpret_cast<double*>(h->arena2);
  CU(cudaMemcpyAsync(stage, host_cm, sizeof(double) * (size_t)rows * ncol, cudaMemcpyHostToDevice, h->stream));
  to_rowmajor_kernel<<<grid_for(rows * h->ld), 256, 0, h->stream>>>(stage, rows, ncol, h->ld, dev_rm);
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}
int download_colmajor(bdf_t* h, const double* dev_rm, int64_t rows, int ncol, double* host_cm) {
  int rc = bdf_ensure_arena2(h, sizeof(double) * (size_t)rows * ncol);
  if (rc) return rc;
  double* stage = reinterpret_cast<double*>(h->arena2);
  to_colmajor_kernel<<<grid_for(rows * ncol), 256, 0, h->stream>>>(dev_rm, rows, ncol, h->ld, stage);
  CU(cudaMemcpyAsync(host_cm, stage, sizeof(double) * (size_t)rows * ncol, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

}  // namespace

// =====================================================================================================================
static void free_feature_state(EntityS& e) {
  for (int r = 0; r < 8; r++) if (e.peer_beta[r]) { cudaIpcCloseMemHandle(e.peer_beta[r]); e.peer_beta[r] = nullptr; }
  cudaFree(e.f_rowptr); cudaFree(e.f_colind); cudaFree(e.f_colptr); cudaFree(e.f_rowind); cudaFree(e.beta); cudaFree(e.uhat); cudaFree(e.cgbuf); cudaFree(e.btb);
  cudaFree(e.f_val_csr); cudaFree(e.f_val_csc); cudaFree(e.f_dense); cudaFree(e.FF);
  e.f_rowptr = e.f_colptr = nullptr; e.f_colind = e.f_rowind = nullptr; e.beta = e.uhat = e.cgbuf = e.btb = nullptr; e.f_val_csr = e.f_val_csc = nullptr;
  e.f_dense = e.FF = nullptr;
  e.use_ff = false;
  for (int o = 0; o < 2; o++) { cudaFree(e.sp_items[o]); cudaFree(e.sp_long[o]); e.sp_items[o] = e.sp_long[o] = nullptr; e.sp_nitems[o] = e.sp_nlong[o] = 0; }
  cudaFree(e.sp_part); e.sp_part = nullptr;
  e.numF = 0; e.fnnz = 0;
}

// beta = zeros(numF, num_latent) (src/RelationData.jl:76), uhat, the betaᵀbeta statistics block, the per-row mean buffer
static int alloc_feature_state(bdf_t* h, EntityS& e, int64_t n, int64_t nnz) {
  int rc;
  const size_t bn = (size_t)n * h->ld, un = (size_t)e.Nper * h->world * h->ld;
  if ((rc = dalloc(h, &e.beta, bn)) || (rc = dalloc(h, &e.uhat, un)) || (rc = dalloc(h, &e.btb, (size_t)1 + h->D + (size_t)h->D * h->D))) return rc;
  CU(cudaMemsetAsync(e.beta, 0, bn * 8, h->stream));
  CU(cudaMemsetAsync(e.uhat, 0, un * 8, h->stream));
  if (!e.mu_rows) { if ((rc = dalloc(h, &e.mu_rows, un))) return rc; CU(cudaMemsetAsync(e.mu_rows, 0, un * 8, h->stream)); }
  e.numF = n; e.fnnz = nnz;
  const size_t need = sizeof(double) * 296 * (size_t)tri(h->D + 1);
  if (h->ws_bytes < need) FAIL(BDF_ERR_STATE, "workspace too small");
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

static int set_features_dev(bdf_t* h, int entity, int64_t m, int64_t n, int64_t nnz, int32_t* d_rows, int32_t* d_cols, const double* d_val) {
  EntityS& e = h->ents[entity];
  free_feature_state(e);
  int rc = build_orientation(h, d_rows, d_cols, nnz, m, (int32_t)n, &e.f_rowptr, &e.f_colind, d_val, &e.f_val_csr);
  if (!rc) rc = build_orientation(h, d_cols, d_rows, nnz, n, (int32_t)m, &e.f_colptr, &e.f_rowind, d_val, &e.f_val_csc);
  if (rc) return rc;
  {
    int64_t slots = 0;
    if ((rc = build_sp_items(h, e, 0, e.f_rowptr, m, &slots)) || (rc = build_sp_items(h, e, 1, e.f_colptr, n, &slots))) return rc;
    CU(cudaMalloc(&e.sp_part, std::max<size_t>((size_t)slots * h->ld, 1) * sizeof(double)));
  }
  return alloc_feature_state(h, e, n, nnz);
}

void bdf_dense_teardown(bdf_t* h) {
  if (h->cublas) cublasDestroy((cublasHandle_t)h->cublas);
  h->cublas = nullptr;
}

// FF = full(At_mul_B(F, F)) (src/RelationData.jl:337-339). Sparse F: num_latent columns of the identity at a time through the two
// gather products (exact integer counts for a 0/1 matrix); dense F: one dgemm.
static int compute_ff_dev(bdf_t* h, EntityS& e) {
  int rc;
  const int64_t n = e.numF;
  if (!e.FF && (rc = dalloc(h, &e.FF, (size_t)n * n))) return rc;
  if (e.f_dense) {
    if ((rc = dense_handles(h))) return rc;
    const double one = 1.0, zero = 0.0;
    if (cublasDgemm((cublasHandle_t)h->cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)n, (int)n, (int)e.N, &one, e.f_dense, (int)e.N, e.f_dense, (int)e.N, &zero, e.FF, (int)n) !=
        CUBLAS_STATUS_SUCCESS)
      FAIL(BDF_ERR_CUDA, "cublasDgemm failed");
    h->launches++;
  } else {
    double *X = nullptr, *T = nullptr, *Y = nullptr;
    if ((rc = dalloc(h, &X, (size_t)n * h->ld)) || (rc = dalloc(h, &T, (size_t)e.N * h->ld)) || (rc = dalloc(h, &Y, (size_t)n * h->ld))) { cudaFree(X); cudaFree(T); return rc; }
    for (int64_t c0 = 0; c0 < n; c0 += h->D) {
      const int nc = (int)std::min<int64_t>(h->D, n - c0);
      identity_block_kernel<<<grid_for(n * h->ld), 256, 0, h->stream>>>(X, n, h->ld, c0, nc);
      spmm(h, e, false, X, T);
      spmm(h, e, true, T, Y);
      scatter_ff_kernel<<<grid_for(n * nc), 256, 0, h->stream>>>(Y, n, h->ld, c0, nc, e.FF);
      h->launches += 2;
    }
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    cudaFree(X); cudaFree(T); cudaFree(Y);
    if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  }
  e.use_ff = true;
  return BDF_OK;
}

// solve_full(FF, rhs, lambda) — src/sampling.jl:314-320: (FF + λI) \ rhs for num_latent right-hand sides; B and X row-major numF × ld.
// The regularised matrix is symmetric positive definite: blocked Cholesky + substitutions of dense_spd.cuh, in place on X.
__global__ void copy_rows_zero_pad_kernel(const double* __restrict__ B, int64_t rows, int ncol, int ld, double* __restrict__ X) {
  const int64_t n = rows * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) X[e] = (int)(e % ld) < ncol ? B[e] : 0.0;
}

static int solve_full_dev(bdf_t* h, EntityS& e, const double* B, double* X, double lambda) {
  int rc;
  if (!e.FF) FAIL(BDF_ERR_STATE, "FF has not been computed (bdf_compute_ff)");
  if (B == X) FAIL(BDF_ERR_INVALID, "solve_full: the right-hand sides and the solution must not alias");
  const int64_t n = e.numF;
  const int D = h->D;
  // the factor lives in the handle's second grow-only arena (the first one holds the caller's B / X): no cudaMalloc per draw
  const size_t nA = ((size_t)n * n + 31) / 32 * 32;
  if ((rc = bdf_ensure_arena2(h, sizeof(double) * (nA + 32)))) return rc;
  double* A = reinterpret_cast<double*>(h->arena2);
  int* info = reinterpret_cast<int*>(A + nA);
  cudaMemcpyAsync(A, e.FF, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, h->stream);
  cudaMemsetAsync(info, 0, sizeof(int), h->stream);
  add_diag_kernel<<<grid_for(n), 256, 0, h->stream>>>(A, n, lambda);
  copy_rows_zero_pad_kernel<<<grid_for(n * h->ld), 256, 0, h->stream>>>(B, n, D, h->ld, X);
  h->launches += 2 + spd::solve(h->stream, A, n, X, h->ld, D, info);
  int hinfo = 0;
  cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t ce = cudaStreamSynchronize(h->stream);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  if (hinfo != 0) FAIL(BDF_ERR_NUMERIC, "solve_full: FF + lambda*I is not positive definite");
  return BDF_OK;
}

// pred_all(r) = udot_all(r) + mean_value — src/sampling.jl:72-97, matrix relations: sample_1' * sample_2 as one dgemm
__global__ void add_scalar_kernel(double* __restrict__ A, int64_t n, double v) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) A[e] += v;
}

// every cell of a 3-mode relation, column-major N1 × N2 × N3 (first index fastest)
__global__ void pred_all3_kernel(const double* __restrict__ U1, const double* __restrict__ U2, const double* __restrict__ U3, int64_t n1, int64_t n2,
                                 int64_t n3, int ld, int D, double mean, double* __restrict__ out) {
  const int64_t n = n1 * n2 * n3;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const double* a = U1 + (size_t)(e % n1) * ld;
    const double* b = U2 + (size_t)((e / n1) % n2) * ld;
    const double* c = U3 + (size_t)(e / (n1 * n2)) * ld;
    double s = 0.0;
    for (int k = 0; k < D; k++) s += a[k] * b[k] * c[k];
    out[e] = s + mean;
  }
}

// factor rows in the caller's row order (pitch ld) out of the slot-ordered replica of a sharded / partitioned entity
__global__ void natural_rows_kernel(const double* __restrict__ U, const int32_t* __restrict__ slot_tab, int world, int64_t nper, int64_t N, int ld,
                                    double* __restrict__ out) {
  const int64_t n = N * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ld;
    const int64_t sl = slot_tab ? slot_tab[i] : (i % world) * nper + i / world;
    out[e] = U[(size_t)sl * ld + e % ld];
  }
}

extern "C" int bdf_predict_all(bdf_t* h, int rel, double* out) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  if (!out) FAIL(BDF_ERR_INVALID, "null argument");
  RelationS& r = h->rels[rel];
  EntityS& a = h->ents[r.entity_of_mode[0]];
  EntityS& b = h->ents[r.entity_of_mode[1]];
  CU(cudaSetDevice(h->device));
  // every rank holds all factor rows (slot order): with several ranks or an explicit shard map the rows are first put back into the
  // caller's order in the second arena
  const double* Un[3] = {a.U, b.U, r.K == 3 ? h->ents[r.entity_of_mode[2]].U : nullptr};
  {
    size_t need = 0;
    for (int m = 0; m < r.K; m++) {
      const EntityS& e = h->ents[r.entity_of_mode[m]];
      if (h->world > 1 || e.slot_of_row) need += (size_t)e.N * h->ld;
    }
    if (need) {
      int rcn = bdf_ensure_arena2(h, sizeof(double) * need);
      if (rcn) return rcn;
      double* dst = reinterpret_cast<double*>(h->arena2);
      for (int m = 0; m < r.K; m++) {
        const EntityS& e = h->ents[r.entity_of_mode[m]];
        if (h->world > 1 || e.slot_of_row) {
          natural_rows_kernel<<<grid_for(e.N * h->ld), 256, 0, h->stream>>>(e.U, e.slot_of_row, h->world, e.Nper, e.N, h->ld, dst);
          h->launches++;
          Un[m] = dst;
          dst += (size_t)e.N * h->ld;
        }
      }
      CU(cudaGetLastError());
    }
  }
  if (r.K == 3) {  // the reference enumerates every cell (src/sampling.jl:78-89); so does this kernel, one thread per cell
    EntityS& c = h->ents[r.entity_of_mode[2]];
    const size_t n3 = (size_t)a.N * b.N * c.N;
    int rc3 = bdf_ensure_arena(h, sizeof(double) * n3);
    if (rc3) return rc3;
    double* Y3 = reinterpret_cast<double*>(h->arena);
    pred_all3_kernel<<<grid_for((int64_t)n3), 256, 0, h->stream>>>(Un[0], Un[1], Un[2], a.N, b.N, c.N, h->ld, h->D, r.mean, Y3);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, Y3, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BDF_OK;
  }
  int rc = dense_handles(h);
  if (rc) return rc;
  const size_t n = (size_t)a.N * b.N;
  if ((rc = bdf_ensure_arena(h, sizeof(double) * n))) return rc;
  double* Y = reinterpret_cast<double*>(h->arena);
  const double one = 1.0, zero = 0.0;
  // factor buffers are row-major N × ld = column-major ld × N: Y (N1 × N2, column-major) = U1·U2ᵀ
  if (cublasDgemm((cublasHandle_t)h->cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)a.N, (int)b.N, h->D, &one, Un[0], h->ld, Un[1], h->ld, &zero, Y, (int)a.N) != CUBLAS_STATUS_SUCCESS)
    FAIL(BDF_ERR_CUDA, "cublasDgemm failed");
  add_scalar_kernel<<<grid_for((int64_t)n), 256, 0, h->stream>>>(Y, (int64_t)n, r.mean);
  h->launches += 2;
  CU(cudaMemcpyAsync(out, Y, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

// ---- relation-level features (N3): sample_beta_rel, src/sampling.jl:322-337; linear_values, src/macau.jl:89-92 --------------------------
//   res    = values − udot(r) − mean_value
//   aFt_y  = α·Fᵀ(res + α^(-1/2)·z1) + sqrt(λ)·z2         z1 ~ N(0, I_nnz), z2 ~ N(0, I_nF) injected or Philox
//   beta   = (α·FF + λ·I) \ aFt_y                        FF = FᵀF precomputed (reset!, src/RelationData.jl:350-352)
//   linear_values = mean_value + F·beta                   the per-observation offset of the row draws and of pred(r)
__global__ void relfeat_noise_kernel(double* __restrict__ res, const double* __restrict__ z1, int64_t n, double sc, uint64_t seed, uint64_t sweep,
                                     uint32_t stream) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
    res[o] += sc * (z1 ? z1[o] : philox_normal(seed, sweep, stream, (uint64_t)o, 0));
}
__global__ void relfeat_rhs_kernel(double* __restrict__ rhs, const double* __restrict__ z2, int64_t nF, double alpha, double sl, uint64_t seed,
                                   uint64_t sweep, uint32_t stream) {
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += (int64_t)gridDim.x * blockDim.x)
    rhs[f] = alpha * rhs[f] + sl * (z2 ? z2[f] : philox_normal(seed, sweep, stream, (uint64_t)f, 1));
}
__global__ void relfeat_k_kernel(const double* __restrict__ FF, int64_t nF, double alpha, double lambda, double* __restrict__ K) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nF * nF; e += (int64_t)gridDim.x * blockDim.x)
    K[e] = alpha * FF[e] + ((e % nF == e / nF) ? lambda : 0.0);
}
__global__ void relfeat_adj_kernel(const double* __restrict__ val, const uint32_t* __restrict__ perm, const double* __restrict__ linear, double mean,
                                   int64_t n, double* __restrict__ out) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
    out[o] = val[o] - (mean + linear[perm[o]]);
}

// val_adj of every mode from the current beta: linear = F·beta, val_adj = val − (mean + linear)
int bdf_refresh_relation_offsets(bdf_t* h, int rel) {
  RelationS& r = h->rels[rel];
  int rc = dense_handles(h);
  if (rc) return rc;
  const double one = 1.0, zero = 0.0;
  if (r.nnz > 0 && cublasDgemv((cublasHandle_t)h->cublas, CUBLAS_OP_N, (int)r.nnz, (int)r.nF, &one, r.F, (int)r.nnz, r.beta, 1, &zero, r.linear, 1) != CUBLAS_STATUS_SUCCESS)
    FAIL(BDF_ERR_CUDA, "cublasDgemv failed");
  for (int m = 0; m < r.K; m++) {
    ModeIndex& mi = r.modes[m];
    relfeat_adj_kernel<<<grid_for(mi.nnz), 256, 0, h->stream>>>(mi.val, mi.perm, r.linear, r.mean, mi.nnz, mi.val_adj);
  }
  h->launches += 1 + r.K;
  CU(cudaGetLastError());
  return BDF_OK;
}

extern "C" int bdf_set_relation_features(bdf_t* h, int rel, int64_t nnz, int64_t nF, const double* F) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (!F || nF < 1 || nF > 32768) FAIL(BDF_ERR_INVALID, "bad feature matrix");
  if (nnz != r.nnz) FAIL(BDF_ERR_INVALID, "Number of rows in the relation's feature matrix must equal the number of training observations");
  if (h->world != 1) FAIL(BDF_ERR_INVALID, "relation-level features are single-GPU in this version");
  CU(cudaSetDevice(h->device));
  int rc = dense_handles(h);
  if (rc) return rc;
  cudaFree(r.F); cudaFree(r.FF); cudaFree(r.beta); cudaFree(r.linear); cudaFree(r.res);
  r.F = r.FF = r.beta = r.linear = r.res = nullptr;
  r.nF = nF;
  const size_t n1 = (size_t)std::max<int64_t>(nnz, 1);
  if ((rc = dalloc(h, &r.F, n1 * nF)) || (rc = dalloc(h, &r.FF, (size_t)nF * nF)) || (rc = dalloc(h, &r.beta, (size_t)nF)) || (rc = dalloc(h, &r.linear, n1)) ||
      (rc = dalloc(h, &r.res, n1)))
    return rc;
  CU(cudaMemcpyAsync(r.F, F, sizeof(double) * (size_t)nnz * nF, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(r.beta, 0, sizeof(double) * nF, h->stream));  // beta = zeros(nF)
  const double one = 1.0, zero = 0.0;
  if (nnz > 0 && cublasDgemm((cublasHandle_t)h->cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)nF, (int)nF, (int)nnz, &one, r.F, (int)nnz, r.F, (int)nnz, &zero, r.FF, (int)nF) !=
                     CUBLAS_STATUS_SUCCESS)
    FAIL(BDF_ERR_CUDA, "cublasDgemm failed");
  for (int m = 0; m < r.K; m++) {
    ModeIndex& mi = r.modes[m];
    cudaFree(mi.val_adj);
    mi.val_adj = nullptr;
    if ((rc = dalloc(h, &mi.val_adj, (size_t)std::max<int64_t>(mi.nnz, 1)))) return rc;
  }
  if ((rc = bdf_refresh_relation_offsets(h, rel))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

extern "C" int bdf_sample_beta_rel(bdf_t* h, int rel, double lambda_beta, const double* z1, const double* z2, double* beta_out) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (!r.F) FAIL(BDF_ERR_STATE, "the relation has no features (bdf_set_relation_features)");
  if (!(lambda_beta > 0.0)) FAIL(BDF_ERR_INVALID, "lambda_beta must be positive");
  CU(cudaSetDevice(h->device));
  int rc = dense_handles(h);
  if (rc) return rc;
  cublasHandle_t cb = (cublasHandle_t)h->cublas;
  const int nF = (int)r.nF;
  const int64_t nnz = r.nnz;
  // temporaries from the handle's second grow-only arena: no cudaMalloc per draw
  auto up32 = [](size_t x) { return (x + 31) / 32 * 32; };
  const size_t nK = up32((size_t)nF * nF), nR = up32((size_t)nF), nZ1 = z1 ? up32((size_t)std::max<int64_t>(nnz, 1)) : 0, nZ2 = z2 ? nR : 0;
  if ((rc = bdf_ensure_arena2(h, sizeof(double) * (nK + nR + nZ1 + nZ2 + 32)))) return rc;
  double* K = reinterpret_cast<double*>(h->arena2);
  double* rhs = K + nK;
  double* dz1 = z1 ? rhs + nR : nullptr;
  double* dz2 = z2 ? rhs + nR + nZ1 : nullptr;
  int* info = reinterpret_cast<int*>(rhs + nR + nZ1 + nZ2);
  if (z1) cudaMemcpyAsync(dz1, z1, sizeof(double) * nnz, cudaMemcpyHostToDevice, h->stream);
  if (z2) cudaMemcpyAsync(dz2, z2, sizeof(double) * nF, cudaMemcpyHostToDevice, h->stream);
  if ((rc = bdf_relation_residuals(h, rel))) return rc;
  const uint32_t st = philox_stream(PHILOX_RELFEAT, 2u * (uint32_t)rel);
  relfeat_noise_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(r.res, dz1, nnz, 1.0 / sqrt(r.alpha), h->seed, h->sweep, st);
  const double one = 1.0, zero = 0.0;
  cublasStatus_t s0 = nnz > 0 ? cublasDgemv(cb, CUBLAS_OP_T, (int)nnz, nF, &one, r.F, (int)nnz, r.res, 1, &zero, rhs, 1) : CUBLAS_STATUS_SUCCESS;
  if (nnz == 0) cudaMemsetAsync(rhs, 0, sizeof(double) * nF, h->stream);
  relfeat_rhs_kernel<<<grid_for(nF), 256, 0, h->stream>>>(rhs, dz2, nF, r.alpha, sqrt(lambda_beta), h->seed, h->sweep, st + 1);
  relfeat_k_kernel<<<grid_for((int64_t)nF * nF), 256, 0, h->stream>>>(r.FF, nF, r.alpha, lambda_beta, K);
  // (alpha·FF + lambda_beta·I) \ rhs: one right-hand side, pitch 1 (dense_spd.cuh)
  int hinfo = 0;
  cudaMemsetAsync(info, 0, sizeof(int), h->stream);
  h->launches += spd::solve(h->stream, K, (int64_t)nF, rhs, 1, 1, info);
  cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaMemcpyAsync(r.beta, rhs, sizeof(double) * nF, cudaMemcpyDeviceToDevice, h->stream);
  h->launches += 5;
  rc = bdf_refresh_relation_offsets(h, rel);
  if (!rc && beta_out) cudaMemcpyAsync(beta_out, r.beta, sizeof(double) * nF, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t ce = cudaStreamSynchronize(h->stream);
  if (rc) return rc;
  if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  if (s0 != CUBLAS_STATUS_SUCCESS) FAIL(BDF_ERR_CUDA, "cublasDgemv failed");
  if (hinfo != 0) FAIL(BDF_ERR_NUMERIC, "sample_beta_rel: alpha*FF + lambda*I is not positive definite");
  return BDF_OK;
}

extern "C" int bdf_get_relation_beta(bdf_t* h, int rel, double* beta) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (!r.F || !beta) FAIL(BDF_ERR_STATE, "the relation has no features");
  CU(cudaMemcpyAsync(beta, r.beta, sizeof(double) * r.nF, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

extern "C" int bdf_set_relation_beta(bdf_t* h, int rel, const double* beta) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (!r.F || !beta) FAIL(BDF_ERR_STATE, "the relation has no features");
  CU(cudaMemcpyAsync(r.beta, beta, sizeof(double) * r.nF, cudaMemcpyHostToDevice, h->stream));
  int rc = bdf_refresh_relation_offsets(h, rel);
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}

extern "C" int bdf_set_features_dense(bdf_t* h, int entity, int64_t m, int64_t n, const double* F) {
  CHECK_H(); CHECK_ENT(entity);
  if (!F || m < 1 || n < 1) FAIL(BDF_ERR_INVALID, "null or empty feature matrix");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (m != e.N) FAIL(BDF_ERR_INVALID, "Number of rows in the feature matrix must equal the entity count");
  if (m >= 2147483647LL || n >= 2147483647LL) FAIL(BDF_ERR_INVALID, "feature matrix too large");
  free_feature_state(e);
  int rc;
  if ((rc = dalloc(h, &e.f_dense, (size_t)m * n))) return rc;
  CU(cudaMemcpyAsync(e.f_dense, F, sizeof(double) * (size_t)m * n, cudaMemcpyHostToDevice, h->stream));
  return alloc_feature_state(h, e, n, m * n);
}

extern "C" int bdf_compute_ff(bdf_t* h, int entity, double* FF_out) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if ((rc = compute_ff_dev(h, e))) return rc;
  if (FF_out) CU(cudaMemcpyAsync(FF_out, e.FF, sizeof(double) * (size_t)e.numF * e.numF, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

extern "C" int bdf_set_use_ff(bdf_t* h, int entity, int use_ff) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  EntityS& e = h->ents[entity];
  if (use_ff && !e.FF) FAIL(BDF_ERR_STATE, "FF has not been computed (bdf_compute_ff)");
  e.use_ff = use_ff != 0;
  return BDF_OK;
}

extern "C" int bdf_solve_full(bdf_t* h, int entity, const double* rhs, int ncol, double lambda, double* x) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!rhs || !x || ncol != h->D) FAIL(BDF_ERR_INVALID, "DimensionMismatch: rhs must have num_latent columns");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  double *db = nullptr, *dx = nullptr;
  if ((rc = dalloc(h, &db, (size_t)e.numF * h->ld)) || (rc = dalloc(h, &dx, (size_t)e.numF * h->ld))) { cudaFree(db); return rc; }
  rc = upload_rowmajor(h, rhs, e.numF, ncol, db);
  if (!rc) rc = solve_full_dev(h, e, db, dx, lambda);
  if (!rc) rc = download_colmajor(h, dx, e.numF, ncol, x);
  cudaFree(db); cudaFree(dx);
  return rc;
}

extern "C" int bdf_set_features_sbm(bdf_t* h, int entity, int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols) {
  CHECK_H(); CHECK_ENT(entity);
  EntityS& e = h->ents[entity];
  if (m != e.N) FAIL(BDF_ERR_INVALID, "DimensionMismatch: number of feature rows must equal the entity count");  // src/RelationData.jl:263-268
  if (n < 1 || n > 2000000000LL || nnz < 0 || nnz >= 2147483647LL) FAIL(BDF_ERR_INVALID, "bad feature matrix size");
  if (nnz > 0 && (!rows || !cols)) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  int32_t *d_rows = nullptr, *d_cols = nullptr;
  int rc;
  if ((rc = dalloc(h, &d_rows, (size_t)nnz)) || (rc = dalloc(h, &d_cols, (size_t)nnz))) { cudaFree(d_rows); return rc; }
  if (nnz) {
    CU(cudaMemcpyAsync(d_rows, rows, 4 * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_cols, cols, 4 * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  }
  rc = set_features_dev(h, entity, m, n, nnz, d_rows, d_cols, nullptr);
  cudaFree(d_rows); cudaFree(d_cols);
  return rc;
}

/* Entity(F = ::SparseMatrixCSC{Float64,Int64}) — the general sparse feature matrix of the reference's own tests
 * (test/parallel_latent_basic.jl:4, test/parallel_mult.jl:4): Julia's CSC fields colptr (n+1), rowval (nnz), nzval (nnz),
 * 1-based. Products accumulate in Julia's order (F*x: ascending column per row; F'x: stored order per column). */
extern "C" int bdf_set_features_csc(bdf_t* h, int entity, int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval) {
  CHECK_H(); CHECK_ENT(entity);
  EntityS& e = h->ents[entity];
  if (m != e.N) FAIL(BDF_ERR_INVALID, "DimensionMismatch: number of feature rows must equal the entity count");
  if (n < 1 || n > 2000000000LL || !colptr) FAIL(BDF_ERR_INVALID, "bad feature matrix size");
  const int64_t nnz = colptr[n] - 1;
  if (nnz < 0 || nnz >= 2147483647LL || (nnz > 0 && (!rowval || !nzval))) FAIL(BDF_ERR_INVALID, "bad colptr / null argument");
  CU(cudaSetDevice(h->device));
  int64_t *d_cp = nullptr, *d_rv = nullptr; double* d_nz = nullptr; int32_t *d_rows = nullptr, *d_cols = nullptr; int* d_bad = nullptr;
  auto cleanup = [&]() { cudaFree(d_cp); cudaFree(d_rv); cudaFree(d_nz); cudaFree(d_rows); cudaFree(d_cols); cudaFree(d_bad); };
  int rc;
  if ((rc = dalloc(h, &d_cp, (size_t)n + 1)) || (rc = dalloc(h, &d_rv, (size_t)nnz)) || (rc = dalloc(h, &d_nz, (size_t)nnz)) ||
      (rc = dalloc(h, &d_rows, (size_t)nnz)) || (rc = dalloc(h, &d_cols, (size_t)nnz)) || (rc = dalloc(h, &d_bad, 1))) { cleanup(); return rc; }
  cudaMemsetAsync(d_bad, 0, 4, h->stream);
  cudaMemcpyAsync(d_cp, colptr, 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, h->stream);
  if (nnz) {
    cudaMemcpyAsync(d_rv, rowval, 8 * (size_t)nnz, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_nz, nzval, 8 * (size_t)nnz, cudaMemcpyHostToDevice, h->stream);
  }
  csc_expand_kernel<<<grid_for(n), 256, 0, h->stream>>>(d_cp, d_rv, n, nnz, d_rows, d_cols, d_bad);
  int bad = 0;
  cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t ce = cudaStreamSynchronize(h->stream);
  if (ce != cudaSuccess) { cleanup(); FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce)); }
  if (bad) { cleanup(); FAIL(BDF_ERR_INVALID, "colptr is not a valid CSC column pointer array"); }
  rc = set_features_dev(h, entity, m, n, nnz, d_rows, d_cols, d_nz);
  cleanup();
  return rc;
}

extern "C" {

/* debug/parity hook: the CSR the device built, in the reference's own representation (1-based Int32 row_ptr of length m+1
 * and col_ind, src/sparsebin_csr.jl:6-11) */
int bdf_debug_features_csr(bdf_t* h, int entity, int transpose, int32_t* ptr_out, int32_t* ind_out) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (e.f_dense) FAIL(BDF_ERR_STATE, "the feature matrix is dense: no CSR representation");
  const int64_t nk = transpose ? e.numF : e.N;
  std::vector<int64_t> p((size_t)nk + 1);
  std::vector<int32_t> ind((size_t)std::max<int64_t>(e.fnnz, 1));
  CU(cudaMemcpyAsync(p.data(), transpose ? e.f_colptr : e.f_rowptr, 8 * p.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(ind.data(), transpose ? e.f_rowind : e.f_colind, 4 * (size_t)e.fnnz, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i <= nk; i++) ptr_out[i] = (int32_t)(p[i] + 1);
  for (int64_t i = 0; i < e.fnnz; i++) ind_out[i] = ind[i] + 1;
  return BDF_OK;
}

int bdf_spmm(bdf_t* h, int entity, int transpose, const double* X, int ncol, double* Y) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!X || !Y || ncol < 1 || ncol > h->D) FAIL(BDF_ERR_INVALID, "DimensionMismatch: need 1 <= ncol <= num_latent");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  const int64_t rin = transpose ? e.N : e.numF, rout = transpose ? e.numF : e.N;
  double *dx = nullptr, *dy = nullptr;
  if ((rc = dalloc(h, &dx, (size_t)rin * h->ld)) || (rc = dalloc(h, &dy, (size_t)rout * h->ld))) { cudaFree(dx); return rc; }
  rc = upload_rowmajor(h, X, rin, ncol, dx);
  if (!rc) { spmm(h, e, transpose != 0, dx, dy); rc = download_colmajor(h, dy, rout, ncol, Y); }
  cudaFree(dx); cudaFree(dy);
  return rc;
}

int bdf_ata_mul(bdf_t* h, int entity, const double* x, double lambda, double* y) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!x || !y) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  double *dx = nullptr, *dt = nullptr, *dy = nullptr;
  if ((rc = dalloc(h, &dx, (size_t)e.numF * h->ld)) || (rc = dalloc(h, &dt, (size_t)e.N * h->ld)) || (rc = dalloc(h, &dy, (size_t)e.numF * h->ld))) { cudaFree(dx); cudaFree(dt); return rc; }
  rc = upload_rowmajor(h, x, e.numF, 1, dx);
  if (!rc) { spmm(h, e, false, dx, dt); spmm(h, e, true, dt, dy, lambda, dx); rc = download_colmajor(h, dy, e.numF, 1, y); }
  cudaFree(dx); cudaFree(dt); cudaFree(dy);
  return rc;
}

int bdf_cg_solve(bdf_t* h, int entity, const double* rhs, int ncol, double lambda, double tol, int64_t maxiter, double* x, int* iters) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!rhs || !x || ncol != h->D) FAIL(BDF_ERR_INVALID, "DimensionMismatch: rhs must have num_latent columns");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (maxiter <= 0) maxiter = e.numF;                        // maxiter=size(rhs,1), src/parallel_matrix.jl:488
  if (tol != tol) tol = 2.220446049250313e-16 * (double)e.numF;  // eps()·numF, src/sampling.jl:294-296
  double *db = nullptr, *dx = nullptr;
  if ((rc = dalloc(h, &db, (size_t)e.numF * h->ld)) || (rc = dalloc(h, &dx, (size_t)e.numF * h->ld))) { cudaFree(db); return rc; }
  rc = upload_rowmajor(h, rhs, e.numF, ncol, db);
  if (!rc) rc = cg_solve_dev(h, e, db, dx, lambda, tol, maxiter, iters);
  if (!rc) rc = download_colmajor(h, dx, e.numF, ncol, x);
  cudaFree(db); cudaFree(dx);
  return rc;
}

int bdf_set_beta(bdf_t* h, int entity, const double* beta) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  return upload_rowmajor(h, beta, h->ents[entity].numF, h->D, h->ents[entity].beta);
}
int bdf_get_beta(bdf_t* h, int entity, double* beta) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  return download_colmajor(h, h->ents[entity].beta, h->ents[entity].numF, h->D, beta);
}

/* uhat = (F·beta)' and the per-row prior mean mu .+ uhat (src/macau.jl:102-104); uhat_out (D×N) may be NULL */
int bdf_update_uhat(bdf_t* h, int entity, const double* mu, double* uhat_out) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!mu) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if ((rc = bdf_join_side(h))) return rc;
  CU(cudaMemcpyAsync(e.mu, mu, sizeof(double) * h->D, cudaMemcpyHostToDevice, h->stream));
  const double* uhat_rows = e.uhat;  // row order
  if (h->world == 1 && !e.slot_of_row) {
    spmm(h, e, false, e.beta, e.uhat);
    add_mu_kernel<<<grid_for(e.N * h->ld), 256, 0, h->stream>>>(e.uhat, e.mu, e.N, h->ld, h->D, e.mu_rows);
  } else {
    // sharded entity: every rank holds the full beta and F, computes F·beta for all rows (row order) and files it by slot
    if ((rc = bdf_ensure_arena(h, sizeof(double) * (size_t)e.N * h->ld))) return rc;
    double* tmp = reinterpret_cast<double*>(h->arena);
    spmm(h, e, false, e.beta, tmp);
    add_mu_scatter_kernel<<<grid_for(e.N * h->ld), 256, 0, h->stream>>>(tmp, e.mu, e.N, h->ld, h->D, h->world, e.Nper, e.slot_of_row, e.uhat, e.mu_rows);
    uhat_rows = tmp;
  }
  h->launches++;
  CU(cudaGetLastError());
  if (uhat_out) CU(cudaMemcpy2DAsync(uhat_out, sizeof(double) * h->D, uhat_rows, sizeof(double) * h->ld, sizeof(double) * h->D, (size_t)e.N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

/* sample_latent_all2! with the per-row mean mu .+ uhat left on the device by bdf_update_uhat (src/macau.jl:105) */
int bdf_sample_mode_uhat(bdf_t* h, int entity, const double* Lambda, const double* z) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!Lambda) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  const size_t un = (size_t)e.Nper * h->world * h->ld;
  if ((rc = bdf_join_side(h))) return rc;
  CU(cudaMemcpyAsync(e.Lambda, Lambda, sizeof(double) * h->D * h->D, cudaMemcpyHostToDevice, h->stream));
  const double* zd = nullptr;
  if (z) {
    if (!e.Z) { if ((rc = dalloc(h, &e.Z, un))) return rc; CU(cudaMemsetAsync(e.Z, 0, un * 8, h->stream)); }
    if ((rc = bdf_copy_rows_h2d_impl(h, entity, z, e.Z))) return rc;  // D×N host matrix → slot-major
    zd = e.Z;
  }
  if ((rc = bdf_sample_entity_impl(h, entity, e.mu_rows, h->ld, e.Lambda, zd))) return rc;
  if (h->async_mode && !z) return BDF_OK;  // deferred, see bdf_set_async
  return bdf_check_err_flag(h);
}

/* ConditionalNormalWishart statistics of U − uhat (src/macau.jl:124) */
int bdf_nw_stats_uhat(bdf_t* h, int entity, double* N, double* NU, double* NS) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if ((rc = bdf_stats_of(h, e.U, e.uhat, (int64_t)h->rank * e.Nper, e.nlocal, e.stats))) return rc;  // this rank's rows (all-reduce: caller)
  const int D = h->D;
  std::vector<double> buf((size_t)1 + D + (size_t)D * D);
  CU(cudaMemcpyAsync(buf.data(), e.stats, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (N) *N = buf[0];
  if (NU) memcpy(NU, buf.data() + 1, sizeof(double) * D);
  if (NS) memcpy(NS, buf.data() + 1 + D, sizeof(double) * D * D);
  return BDF_OK;
}

/* beta' * beta (D×D) — src/macau.jl:128, src/sampling.jl:138 */
int bdf_beta_gram(bdf_t* h, int entity, double* BtB) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if ((rc = bdf_stats_of(h, e.beta, nullptr, 0, e.numF, e.btb))) return rc;
  if (BtB) CU(cudaMemcpyAsync(BtB, e.btb + 1 + h->D, sizeof(double) * h->D * h->D, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

/* sample_beta(entity, sample_u_c, Lambda_u, lambda_beta, use_ff=false, tol) — src/sampling.jl:291-312 — with the CG solver
 * (solve_cg2). E1 (D×N) and E2 (D×numF) are the injected standard normals behind rand(mv, N) / rand(mv, numF), consumed in
 * that order; NULL = device Philox. Leaves beta on the device; beta_out (numF×D) / rhs_out (numF×D, the Ft_y the
 * reference also returns) / iters_out (D) may be NULL. */
int bdf_sample_beta(bdf_t* h, int entity, const double* mu, const double* Lambda, double lambda_beta, double tol, const double* E1,
                    const double* E2, double* beta_out, double* rhs_out, int* iters_out) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!mu || !Lambda) FAIL(BDF_ERR_INVALID, "null argument");
  if (!(lambda_beta > 0.0)) FAIL(BDF_ERR_INVALID, "lambda_beta must be positive");
  CU(cudaSetDevice(h->device));
  if ((rc = bdf_join_side(h))) return rc;  // e.mu / e.Lambda / h->scratch may still belong to an asynchronous draw
  EntityS& e = h->ents[entity];
  const int D = h->D, ld = h->ld;
  const size_t dd = (size_t)D * D;
  if (tol != tol) tol = 2.220446049250313e-16 * (double)e.numF;
  CU(cudaMemcpyAsync(e.mu, mu, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(e.Lambda, Lambda, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream));
  // temporaries from the handle's grow-only arena (no per-call cudaMalloc/cudaFree)
  const size_t nT = (size_t)e.N * ld, nR = (size_t)e.numF * ld;
  // column-split solve over the ranks (the reference's own strategy, solve_cg2): needs every peer's beta mapped (bdf_ipc_import_beta)
  int npeer = 0;
  for (int r = 0; r < 8; r++) npeer += e.peer_beta[r] != nullptr;
  const bool sharded = h->world > 1 && npeer == h->world - 1 && !e.use_ff && !e.f_dense;
  if (sharded && (beta_out || rhs_out)) FAIL(BDF_ERR_INVALID, "column-split beta solve: fetch beta with bdf_get_beta once the ranks have synchronised (beta_out / rhs_out must be NULL)");
  if ((rc = bdf_ensure_arena(h, sizeof(double) * (dd + 2 * nT + 4 * nR + 64)))) return rc;
  double* Cm = reinterpret_cast<double*>(h->arena);
  double* T = Cm + ((dd + 31) / 32) * 32;
  double* rhs = T + nT;
  double* E1d = E1 ? rhs + nR : nullptr;
  double* E2d = E2 ? rhs + nR + nT : nullptr;
  color_matrix_kernel<<<1, 256, 0, h->stream>>>(e.Lambda, D, h->scratch, Cm, h->err_flag);
  if (E1) cudaMemcpy2DAsync(E1d, sizeof(double) * ld, E1, sizeof(double) * D, sizeof(double) * D, (size_t)e.N, cudaMemcpyHostToDevice, h->stream);
  if (E2) cudaMemcpy2DAsync(E2d, sizeof(double) * ld, E2, sizeof(double) * D, sizeof(double) * D, (size_t)e.numF, cudaMemcpyHostToDevice, h->stream);
  const size_t smem = sizeof(double) * (dd + 4 * (size_t)D);
  if (!(h->smem_optin & BDF_OPTIN_COLORED)) {
    CU(cudaFuncSetAttribute(colored_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (128 * 128 + 4 * 128))));
    h->smem_optin |= BDF_OPTIN_COLORED;
  }
  const uint32_t s0 = philox_stream(PHILOX_BETA, 2u * (uint32_t)entity);  // +0: rand(mv, N), +1: rand(mv, numF)
  // T = (U − mu) + C·E1 ; rhs = Fᵀ·T ; rhs += sqrt(lambda_beta)·C·E2
  colored_rows_kernel<<<grid_for(e.N, 4), 128, smem, h->stream>>>(e.U, e.mu, Cm, E1d, e.N, ld, D, 1.0, h->seed, h->sweep, s0, T, 0, h->world, e.Nper, e.slot_of_row);
  spmm(h, e, true, T, rhs);
  colored_rows_kernel<<<grid_for(e.numF, 4), 128, smem, h->stream>>>(nullptr, nullptr, Cm, E2d, e.numF, ld, D, sqrt(lambda_beta), h->seed, h->sweep, s0 + 1, rhs, 1);
  h->launches += 3;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  if (e.use_ff) {  // use_ff: solve_full(entity.FF, Ft_y, lambda_beta), src/sampling.jl:303-304
    rc = solve_full_dev(h, e, rhs, e.beta, lambda_beta);
    if (iters_out) for (int d = 0; d < D; d++) iters_out[d] = 0;
  } else if (sharded) {
    const int c0 = (int)((int64_t)D * h->rank / h->world), c1 = (int)((int64_t)D * (h->rank + 1) / h->world);
    const int Dw = c1 - c0, ldw = (std::max(Dw, 1) + 3) / 4 * 4;
    if (iters_out) for (int d = 0; d < D; d++) iters_out[d] = 0;  // the other ranks' columns: the caller adds the vectors up
    if (Dw > 0) {
      double* Bw = rhs + nR + nT + nR;  // behind the E1 / E2 staging areas
      double* Xw = Bw + nR;
      window_extract_kernel<<<grid_for((int64_t)e.numF * ldw), 256, 0, h->stream>>>(rhs, e.numF, ld, c0, Dw, ldw, Bw);
      h->launches++;
      std::vector<int> itw(D, 0);
      rc = cg_solve_dev(h, e, Bw, Xw, lambda_beta, tol, e.numF, itw.data(), Dw, ldw);
      if (!rc) {
        PeerPtrs pp;
        for (int r = 0; r < 8; r++) pp.p[r] = e.peer_beta[r];
        window_scatter_kernel<<<grid_for((int64_t)e.numF * Dw), 256, 0, h->stream>>>(Xw, e.numF, ld, c0, Dw, ldw, e.beta, pp);
        h->launches++;
        if (iters_out) for (int d = 0; d < Dw; d++) iters_out[c0 + d] = itw[d];
      }
    }
  } else {
    rc = cg_solve_dev(h, e, rhs, e.beta, lambda_beta, tol, e.numF, iters_out);
  }
  if (!rc && beta_out) rc = download_colmajor(h, e.beta, e.numF, D, beta_out);
  if (!rc && rhs_out) rc = download_colmajor(h, rhs, e.numF, D, rhs_out);
  if (!rc) rc = bdf_check_err_flag(h); else cudaStreamSynchronize(h->stream);
  return rc;
}

/* CUDA IPC handle of this rank's beta buffer / mapping of a peer's: with every peer mapped, bdf_sample_beta solves only this rank's share of
 * the num_latent right-hand sides and stores the solved columns into all replicas (see window_scatter_kernel). */
int bdf_ipc_export_beta(bdf_t* h, int entity, unsigned char* handle64) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!handle64) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  CU(cudaIpcGetMemHandle(&mh, h->ents[entity].beta));
  memcpy(handle64, &mh, 64);
  return BDF_OK;
}
int bdf_ipc_import_beta(bdf_t* h, int entity, int peer_rank, const unsigned char* handle64) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!handle64 || peer_rank < 0 || peer_rank >= h->world || peer_rank >= 8 || peer_rank == h->rank) FAIL(BDF_ERR_INVALID, "bad peer rank (up to 8 ranks)");
  CU(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void* ptr = nullptr;
  CU(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
  h->ents[entity].peer_beta[peer_rank] = (double*)ptr;
  return BDF_OK;
}

/* Profiling hook: `reps` device-resident applications of (FᵀF + λI) to the current beta (D columns at once), CUDA-event
 * timed on the handle's stream; returns the mean milliseconds per application (two sparse-binary products). */
int bdf_debug_ata_time_window(bdf_t* h, int entity, int reps, int ncols, double* ms_per_apply);
int bdf_debug_ata_time(bdf_t* h, int entity, int reps, double* ms_per_apply) { return bdf_debug_ata_time_window(h, entity, reps, 0, ms_per_apply); }
int bdf_debug_ata_time_window(bdf_t* h, int entity, int reps, int ncols, double* ms_per_apply) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  double *T = nullptr, *Z = nullptr;
  if ((rc = dalloc(h, &T, (size_t)e.N * h->ld)) || (rc = dalloc(h, &Z, (size_t)e.numF * h->ld))) { cudaFree(T); return rc; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int ldw = ncols > 0 && ncols < h->D ? (ncols + 3) / 4 * 4 : 0;  // a packed column window like a rank's share of a column-split CG
  const bool exact = ncols < 0;  // ncols < 0: the in-order kernel pair of bdf_ata_mul; else the order-free pair the CG iteration runs
  spmm(h, e, false, e.beta, T, 0.0, nullptr, ldw, exact); spmm(h, e, true, T, Z, 1.0, e.beta, ldw, exact);
  cudaEventRecord(a, h->stream);
  for (int i = 0; i < reps; i++) { spmm(h, e, false, e.beta, T, 0.0, nullptr, ldw, exact); spmm(h, e, true, T, Z, 1.0, e.beta, ldw, exact); }
  cudaEventRecord(b, h->stream);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(T); cudaFree(Z);
  if (ms_per_apply) *ms_per_apply = ms / reps;
  return BDF_OK;
}

/* sample_lambda_beta(beta, Lambda_u, nu, mu) — src/sampling.jl:136-142. gamma_variate: the injected Gamma(shape, 1) draw
 * behind rand(Gamma(b, c)), or NaN for the device Philox stream. */
int bdf_sample_lambda_beta(bdf_t* h, int entity, const double* Lambda, double nu, double mu, double gamma_variate, double* lambda_beta_out,
                           double* shape_out) {
  CHECK_H(); CHECK_ENT(entity);
  int rc = need_features(h, entity);
  if (rc) return rc;
  if (!Lambda || !lambda_beta_out) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  if ((rc = bdf_join_side(h))) return rc;
  EntityS& e = h->ents[entity];
  CU(cudaMemcpyAsync(e.Lambda, Lambda, sizeof(double) * h->D * h->D, cudaMemcpyHostToDevice, h->stream));
  if ((rc = bdf_stats_of(h, e.beta, nullptr, 0, e.numF, e.btb))) return rc;
  double* out = h->scratch;  // 2 doubles
  lambda_beta_kernel<<<1, 32, 0, h->stream>>>(e.btb, e.Lambda, h->D, (double)e.numF, nu, mu, gamma_variate, h->seed, h->sweep, philox_stream(PHILOX_LAMBDA_BETA, (uint32_t)entity), out);
  h->launches++;
  double res[2];
  CU(cudaMemcpyAsync(res, out, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *lambda_beta_out = res[0];
  if (shape_out) *shape_out = res[1];
  return BDF_OK;
}

}  // extern "C"
