/*
 * bdf_oracle.c — CPU restatement of BayesianDataFusion.jl's latent-factor Gibbs hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU baseline. The product (libbdf_b200.so) never
 * links, loads or calls this file.
 *
 * Every function restates one reference function and cites it (paths relative to the
 * reference checkout). The reference's OPERATION ORDER is kept on purpose — explicit LU
 * inverse (Julia `inv` = LAPACK getrf+getri), Cholesky of the covariance's upper triangle
 * (Julia `chol(Hermitian(covar))` = potrf 'U'), column-major storage, 1-based indices at
 * the interface — so that "parity with the oracle" means "parity with the reference's
 * arithmetic", up to the rounding differences between these plain loops and OpenBLAS'
 * blocked kernels (κ(Λ*)·ε).
 *
 * Parity pinning: Julia is not installable here, so the reference itself cannot run.
 * The restatement is pinned by (a) the reference's own known-answer fixtures
 * (test/sparsebin_csr.jl, test/sparse_csr.jl, test/parallel_matrix.jl, test/solver.jl,
 * test/basic.jl, test/parallel_latent_basic.jl), and (b) a numpy/scipy twin (oracle/oracle.py) that calls
 * the SAME LAPACK routines Julia Base calls (dgetrf/dgetri/dpotrf). The Wishart / MvNormal /
 * Gamma streams of Distributions.jl are NOT pinned by any reference test ("parity unpinned"
 * for those draws): all randomness is injected as standard variates, and the restated
 * conventions (Bartlett factor, chol_lower colouring) are checked as DISTRIBUTIONS against
 * scipy.stats (tests/test_oracle.py), which is as far as they can be pinned without Julia.
 *
 * All matrices are column-major double. Index arrays that cross this interface are 1-based
 * exactly as Julia holds them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, ld) ((size_t)(i) + (size_t)(j) * (size_t)(ld))

/* ------------------------------------------------------------------------------------------
 * Dense helpers restating the Julia Base / LAPACK calls the reference makes [ext].
 * ---------------------------------------------------------------------------------------- */

/* LU with partial pivoting, unblocked (LAPACK dgetf2 order). Returns 0 or k+1 on zero pivot. */
static int lu_factor(int n, double* A, int* piv) {
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(A[IDX(k, k, n)]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(A[IDX(i, k, n)]);
      if (v > best) { best = v; p = i; }
    }
    piv[k] = p;
    if (A[IDX(p, k, n)] == 0.0) return k + 1;
    if (p != k)
      for (int j = 0; j < n; j++) {
        double t = A[IDX(k, j, n)]; A[IDX(k, j, n)] = A[IDX(p, j, n)]; A[IDX(p, j, n)] = t;
      }
    double inv = 1.0 / A[IDX(k, k, n)];
    for (int i = k + 1; i < n; i++) A[IDX(i, k, n)] *= inv;
    for (int j = k + 1; j < n; j++) {
      double akj = A[IDX(k, j, n)];
      for (int i = k + 1; i < n; i++) A[IDX(i, j, n)] -= A[IDX(i, k, n)] * akj;
    }
  }
  return 0;
}

/* Julia `inv(A)` for a dense Matrix = LAPACK getrf! + getri! [ext]. getri: invert U in place,
 * then solve inv(A)·L = inv(U) column by column from the right, then undo the pivoting. */
int orc_inv(int n, double* A /* in: A, out: inv(A) */) {
  int* piv = (int*)malloc(sizeof(int) * (size_t)n);
  double* work = (double*)malloc(sizeof(double) * (size_t)n);
  int info = lu_factor(n, A, piv);
  if (info) { free(piv); free(work); return info; }
  /* dtrti2 'U','N': invert upper triangle in place */
  for (int j = 0; j < n; j++) {
    A[IDX(j, j, n)] = 1.0 / A[IDX(j, j, n)];
    double ajj = -A[IDX(j, j, n)];
    /* x = U(0:j-1,0:j-1)^{-1}(already inverted) * A(0:j-1, j) : dtrmv upper */
    for (int i = 0; i < j; i++) {
      double s = 0.0;
      for (int k = i; k < j; k++) s += A[IDX(i, k, n)] * A[IDX(k, j, n)];
      work[i] = s;
    }
    for (int i = 0; i < j; i++) A[IDX(i, j, n)] = work[i] * ajj;
  }
  /* dgetri unblocked: for j = n-1..0: copy L(:,j) to work, zero it, A(:,j) -= A(:,j+1:n)*work(j+1:n) */
  for (int j = n - 2; j >= 0; j--) {
    for (int i = j + 1; i < n; i++) { work[i] = A[IDX(i, j, n)]; A[IDX(i, j, n)] = 0.0; }
    for (int k = j + 1; k < n; k++) {
      double w = work[k];
      for (int i = 0; i < n; i++) A[IDX(i, j, n)] -= A[IDX(i, k, n)] * w;
    }
  }
  for (int j = n - 2; j >= 0; j--) {
    int p = piv[j];
    if (p != j)
      for (int i = 0; i < n; i++) {
        double t = A[IDX(i, j, n)]; A[IDX(i, j, n)] = A[IDX(i, p, n)]; A[IDX(i, p, n)] = t;
      }
  }
  free(piv); free(work);
  return 0;
}

/* Julia `chol(Hermitian(A))` = potrf 'U' on the UPPER triangle: A = RᵀR, R upper [ext].
 * Writes R into the upper triangle of R_out (lower part zeroed). */
int orc_chol_upper(int n, const double* A, double* R) {
  memset(R, 0, sizeof(double) * (size_t)n * (size_t)n);
  for (int j = 0; j < n; j++) {
    double s = A[IDX(j, j, n)];
    for (int k = 0; k < j; k++) s -= R[IDX(k, j, n)] * R[IDX(k, j, n)];
    if (!(s > 0.0)) return j + 1;
    double rjj = sqrt(s);
    R[IDX(j, j, n)] = rjj;
    for (int i = j + 1; i < n; i++) {
      double t = A[IDX(j, i, n)]; /* upper triangle element (j,i) */
      for (int k = 0; k < j; k++) t -= R[IDX(k, j, n)] * R[IDX(k, i, n)];
      R[IDX(j, i, n)] = t / rjj;
    }
  }
  return 0;
}

/* Dense solve A\B by LU (Julia `\` for square dense) [ext]; A n×n destroyed, B n×nrhs overwritten. */
int orc_lu_solve(int n, int nrhs, double* A, double* B) {
  int* piv = (int*)malloc(sizeof(int) * (size_t)n);
  int info = lu_factor(n, A, piv);
  if (info) { free(piv); return info; }
  for (int c = 0; c < nrhs; c++) {
    double* b = B + (size_t)c * n;
    for (int k = 0; k < n; k++) { int p = piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int k = 0; k < n; k++) { double bk = b[k]; for (int i = k + 1; i < n; i++) b[i] -= A[IDX(i, k, n)] * bk; }
    for (int k = n - 1; k >= 0; k--) { b[k] /= A[IDX(k, k, n)]; double bk = b[k]; for (int i = 0; i < k; i++) b[i] -= A[IDX(i, k, n)] * bk; }
  }
  free(piv);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A1. Observation index — src/IndexedDF.jl:6-21 (IndexedDF ctor), :46-70 (FastIDF, getData).
 * index[mode][i] = positions (table order) of the observations whose id in `mode` is i.
 * Stored here as one CSR per mode: ptr (N_m+1, 0-based offsets) and pos (nnz, 0-based table rows).
 * ---------------------------------------------------------------------------------------- */
int orc_idf_build(int K, const int64_t* dims, int64_t nnz, const int64_t* ids /* nnz×K col-major, 1-based */,
                  int64_t** ptr_out /* K arrays, caller-allocated dims[m]+1 */, int64_t** pos_out /* K arrays nnz */) {
  for (int m = 0; m < K; m++) {
    int64_t N = dims[m];
    int64_t* ptr = ptr_out[m];
    int64_t* pos = pos_out[m];
    memset(ptr, 0, sizeof(int64_t) * (size_t)(N + 1));
    const int64_t* col = ids + (size_t)m * (size_t)nnz;
    for (int64_t r = 0; r < nnz; r++) {
      int64_t j = col[r];
      if (j < 1 || j > N) return -1;
      ptr[j]++;
    }
    for (int64_t i = 0; i < N; i++) ptr[i + 1] += ptr[i];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; i++) fill[i] = ptr[i];
    for (int64_t r = 0; r < nnz; r++) pos[fill[col[r] - 1]++] = r; /* push! in table order: IndexedDF.jl:13-18 */
    free(fill);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A4/A5/A5'. Per-entity conditional draw.
 *   sample_user_basic (matrix)  src/sampling.jl:200-212
 *   sample_user_basic (tensor)  src/sampling.jl:215-234
 *   sample_user2 (multi-rel)    src/sampling.jl:266-289
 * One routine covers all three: the caller hands, per relation, the row's observations
 * (partner ids per other mode, values, per-observation offset). Steps, in reference order:
 *   MM      = ∏_{j≠mode} U_j[:, id[:,j]]                       (:205 / :222-227 / :277-280)
 *   Λ*      = Λ + Σ_rel α·MM·MMᵀ                               (:207 / :281)
 *   rhs     = Λ·μᵢ + Σ_rel α·MM·rr,  rr = v − offset           (:208 / :268,:282)
 *   covar   = inv(Λ*)                                           (:207 / :284)
 *   mu      = covar·rhs                                         (:208 / :285)
 *   x       = chol(Hermitian(covar))ᵀ·z + mu                    (:211 / :288)
 * z is the injected standard-normal vector the reference would get from randn(D).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int nother;                 /* number of partner modes (K-1) */
  const double* const* U;     /* nother pointers: partner factor matrices, D×N_j col-major */
  const int64_t* const* ids;  /* nother pointers: nobs 1-based partner ids */
  const double* vals;         /* nobs values */
  const double* offsets;      /* nobs per-observation offsets (linear_values) or NULL */
  double offset;              /* mean_value, used when offsets == NULL */
  double alpha;
  int64_t nobs;
} orc_rowdata;

int orc_sample_row(int D, int nrel, const orc_rowdata* rels, const double* mu, const double* Lambda,
                   const double* z, double* x_out, double* work /* ≥ 4*D*D + 4*D doubles, or NULL */) {
  double* own = NULL;
  if (!work) { own = (double*)malloc(sizeof(double) * (size_t)(4 * D * D + 4 * D)); work = own; }
  double* Ls = work;             /* Λ* then covar */
  double* R = Ls + D * D;
  double* rhs = R + D * D;
  double* mm = rhs + D;
  double* mvec = mm + D;
  memcpy(Ls, Lambda, sizeof(double) * (size_t)D * D);
  /* Λ·μ (gemv, column-major) */
  for (int i = 0; i < D; i++) rhs[i] = 0.0;
  for (int j = 0; j < D; j++) { double mj = mu[j]; for (int i = 0; i < D; i++) rhs[i] += Lambda[IDX(i, j, D)] * mj; }
  for (int r = 0; r < nrel; r++) {
    const orc_rowdata* rd = &rels[r];
    double* G = R; /* MM·MMᵀ accumulated separately, then scaled by α as the reference does */
    memset(G, 0, sizeof(double) * (size_t)D * D);
    for (int i = 0; i < D; i++) mvec[i] = 0.0;
    for (int64_t o = 0; o < rd->nobs; o++) {
      const double* u0 = rd->U[0] + (size_t)(rd->ids[0][o] - 1) * D;
      for (int i = 0; i < D; i++) mm[i] = u0[i];
      for (int j = 1; j < rd->nother; j++) {
        const double* uj = rd->U[j] + (size_t)(rd->ids[j][o] - 1) * D;
        for (int i = 0; i < D; i++) mm[i] *= uj[i];
      }
      double rr = rd->vals[o] - (rd->offsets ? rd->offsets[o] : rd->offset);
      for (int j = 0; j < D; j++) { double mj = mm[j]; for (int i = 0; i <= j; i++) G[IDX(i, j, D)] += mm[i] * mj; }
      for (int i = 0; i < D; i++) mvec[i] += mm[i] * rr;
    }
    for (int j = 0; j < D; j++)
      for (int i = 0; i <= j; i++) {
        double g = rd->alpha * G[IDX(i, j, D)];
        Ls[IDX(i, j, D)] += g;
        if (i != j) Ls[IDX(j, i, D)] += g;
      }
    for (int i = 0; i < D; i++) rhs[i] += rd->alpha * mvec[i];
  }
  int info = orc_inv(D, Ls);
  if (info) { free(own); return info; }
  for (int i = 0; i < D; i++) mvec[i] = 0.0;
  for (int j = 0; j < D; j++) { double rj = rhs[j]; for (int i = 0; i < D; i++) mvec[i] += Ls[IDX(i, j, D)] * rj; }
  info = orc_chol_upper(D, Ls, R);
  if (info) { free(own); return 1000 + info; }
  /* x = Rᵀ z + mu */
  for (int i = 0; i < D; i++) {
    double s = 0.0;
    for (int k = 0; k <= i; k++) s += R[IDX(k, i, D)] * z[k];
    x_out[i] = s + mvec[i];
  }
  free(own);
  return 0;
}

/* A2/A3. sample_latent_all2! + sample_latent_range — src/sampling.jl:149-198. Half-sweep over one
 * mode of ONE relation (the only case the multi-worker path handles, src/macau.jl:44-47).
 * Rows are visited in the reference's cyclic shards i:P:N (:154); with injected noise the result is
 * independent of P, which the OpenMP schedule mirrors (shard p = thread p).
 *   ptr/pos: this mode's index from orc_idf_build; ids: nnz×K table; U[m]: D×N_m factors (U[mode] is output)
 *   mu: D (mu_ld == 0) or D×N matrix (mu_ld == D), Z: D×N injected normals.                       */
int orc_sample_latent_all(int D, int K, int mode /* 0-based */, const int64_t* dims, int64_t nnz,
                          const int64_t* ids, const double* vals, const int64_t* ptr, const int64_t* pos,
                          double* const* U, double alpha, double mean_value, const double* linear_values,
                          const double* mu, int64_t mu_ld, const double* Lambda, const double* Z, int nshards) {
  int64_t N = dims[mode];
  int err = 0;
  if (nshards < 1) nshards = 1;
#pragma omp parallel for schedule(static, 1) num_threads(nshards) reduction(| : err)
  for (int p = 0; p < nshards; p++) {
    double* work = (double*)malloc(sizeof(double) * (size_t)(4 * D * D + 4 * D));
    int64_t cap = 0;
    int64_t** idbuf = (int64_t**)calloc((size_t)K, sizeof(int64_t*));
    double* vbuf = NULL;
    double* obuf = NULL;
    const double* Uo[16];
    const int64_t* ido[16];
    for (int64_t i = p; i < N; i += nshards) { /* StepRange i:Nprocs:N, :154 */
      int64_t n = ptr[i + 1] - ptr[i];
      if (n > cap) {
        cap = n * 2;
        for (int j = 0; j < K - 1; j++) idbuf[j] = (int64_t*)realloc(idbuf[j], sizeof(int64_t) * (size_t)cap);
        vbuf = (double*)realloc(vbuf, sizeof(double) * (size_t)cap);
        obuf = (double*)realloc(obuf, sizeof(double) * (size_t)cap);
      }
      /* getData: copies ids[idx,:], values[idx] — src/IndexedDF.jl:67-70 */
      int no = 0;
      for (int m = 0; m < K; m++) {
        if (m == mode) continue;
        for (int64_t o = 0; o < n; o++) idbuf[no][o] = ids[(size_t)m * nnz + pos[ptr[i] + o]];
        Uo[no] = U[m];
        ido[no] = idbuf[no];
        no++;
      }
      for (int64_t o = 0; o < n; o++) vbuf[o] = vals[pos[ptr[i] + o]];
      if (linear_values) for (int64_t o = 0; o < n; o++) obuf[o] = linear_values[pos[ptr[i] + o]];
      orc_rowdata rd;
      rd.nother = K - 1; rd.U = Uo; rd.ids = ido; rd.vals = vbuf;
      rd.offsets = linear_values ? obuf : NULL; rd.offset = mean_value; rd.alpha = alpha; rd.nobs = n;
      const double* mui = mu_ld ? mu + (size_t)i * mu_ld : mu;
      int info = orc_sample_row(D, 1, &rd, mui, Lambda, Z + (size_t)i * D, U[mode] + (size_t)i * D, work);
      if (info) err |= 1;
    }
    for (int j = 0; j < K - 1; j++) free(idbuf[j]);
    free(idbuf); free(vbuf); free(obuf); free(work);
  }
  return err;
}

/* ------------------------------------------------------------------------------------------
 * A6. ConditionalNormalWishart — src/sampling.jl:116-127.
 *   N=size(U,2); NU=sum(U,2); NS=U*U'; nu_N=nu+N; beta_N=beta_0+N; mu_N=(beta_0*mu+NU)/(beta_0+N)
 *   T_N = inv(Symmetric(Tinv + NS + beta_0*mu*mu' - beta_N*mu_N*mu_N'))
 * Julia inv(Symmetric) = Bunch-Kaufman sytrf/sytri [ext]; restated with the LU inverse on the
 * symmetrised (upper-triangle-mirrored) matrix — same value up to rounding.
 * ---------------------------------------------------------------------------------------- */
int orc_nw_stats(int D, int64_t N, const double* U, const double* uhat /* or NULL: U - uhat, src/macau.jl:124 */,
                 double* NU, double* NS) {
  for (int i = 0; i < D; i++) NU[i] = 0.0;
  memset(NS, 0, sizeof(double) * (size_t)D * D);
  double* u = (double*)malloc(sizeof(double) * (size_t)D);
  for (int64_t c = 0; c < N; c++) {
    for (int i = 0; i < D; i++) u[i] = U[(size_t)c * D + i] - (uhat ? uhat[(size_t)c * D + i] : 0.0);
    for (int i = 0; i < D; i++) NU[i] += u[i];
    for (int j = 0; j < D; j++) { double uj = u[j]; for (int i = 0; i <= j; i++) NS[IDX(i, j, D)] += u[i] * uj; }
  }
  for (int j = 0; j < D; j++) for (int i = 0; i < j; i++) NS[IDX(j, i, D)] = NS[IDX(i, j, D)];
  free(u);
  return 0;
}

int orc_cond_normal_wishart(int D, double N, const double* NU, const double* NS, const double* mu0, double beta0,
                            const double* Tinv, double nu, double* mu_N, double* beta_N, double* T_N, double* nu_N) {
  *nu_N = nu + N;
  *beta_N = beta0 + N;
  for (int i = 0; i < D; i++) mu_N[i] = (beta0 * mu0[i] + NU[i]) / (beta0 + N);
  for (int j = 0; j < D; j++)
    for (int i = 0; i <= j; i++) {
      double v = Tinv[IDX(i, j, D)] + NS[IDX(i, j, D)] + beta0 * mu0[i] * mu0[j] - (*beta_N) * mu_N[i] * mu_N[j];
      T_N[IDX(i, j, D)] = v; /* Symmetric(...) reads the upper triangle */
      T_N[IDX(j, i, D)] = v;
    }
  return orc_inv(D, T_N);
}

/* A7. rand(::NormalWishart) — src/normal_wishart.jl:38-42:
 *   Lam = rand(Wishart(nu, T));  mu = rand(MvNormal(nw.mu, inv(Symmetric(Lam)) ./ kappa))
 * Distributions.jl [ext, Bartlett]: Z = chol_lower(T)·A, Lam = Z·Zᵀ, with A lower-triangular,
 * A[i,i] = sqrt(χ²(nu−i+1)) (1-based i), A[i>j] ~ N(0,1); MvNormal: mu + chol_lower(Σ)·z.
 * A (D×D, lower) and z (D) are injected. The NormalWishart ctor stores full(Symmetric(T)) (:27). */
int orc_nw_rand(int D, const double* mu_N, double kappa, const double* T, const double* A, const double* z,
                double* mu_out, double* Lam_out) {
  size_t dd = (size_t)D * D;
  double* Ts = (double*)malloc(sizeof(double) * dd * 4);
  double* R = Ts + dd; double* Zm = R + dd; double* S = Zm + dd;
  for (int j = 0; j < D; j++) for (int i = 0; i <= j; i++) { Ts[IDX(i, j, D)] = T[IDX(i, j, D)]; Ts[IDX(j, i, D)] = T[IDX(i, j, D)]; }
  int info = orc_chol_upper(D, Ts, R); /* T = RᵀR, chol_lower(T) = Rᵀ */
  if (info) { free(Ts); return info; }
  /* Zm = Rᵀ·A  (A lower triangular) */
  for (int j = 0; j < D; j++)
    for (int i = 0; i < D; i++) {
      double s = 0.0;
      int kmax = i; /* Rᵀ[i,k] = R[k,i] nonzero for k ≤ i; A[k,j] nonzero for k ≥ j */
      for (int k = j; k <= kmax; k++) s += R[IDX(k, i, D)] * A[IDX(k, j, D)];
      Zm[IDX(i, j, D)] = s;
    }
  for (int j = 0; j < D; j++)
    for (int i = 0; i < D; i++) {
      double s = 0.0;
      for (int k = 0; k < D; k++) s += Zm[IDX(i, k, D)] * Zm[IDX(j, k, D)];
      Lam_out[IDX(i, j, D)] = s;
    }
  /* Σ = inv(Symmetric(Lam)) ./ kappa ; mu = mu_N + chol_lower(Σ)·z */
  for (int j = 0; j < D; j++) for (int i = 0; i <= j; i++) { S[IDX(i, j, D)] = Lam_out[IDX(i, j, D)]; S[IDX(j, i, D)] = Lam_out[IDX(i, j, D)]; }
  info = orc_inv(D, S);
  if (info) { free(Ts); return 2000 + info; }
  for (size_t k = 0; k < dd; k++) S[k] /= kappa;
  info = orc_chol_upper(D, S, R);
  if (info) { free(Ts); return 3000 + info; }
  for (int i = 0; i < D; i++) {
    double s = 0.0;
    for (int k = 0; k <= i; k++) s += R[IDX(k, i, D)] * z[k];
    mu_out[i] = mu_N[i] + s;
  }
  free(Ts);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A12. SparseBinMatrix (COO) — src/parallel_matrix.jl:9-24, A_mul_B! :242-252, At_mul_B! :257-267.
 * rows/cols Int32 1-based, arbitrary order; sums in list order.
 * ---------------------------------------------------------------------------------------- */
void orc_sbm_mul(int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const double* x, double* y) {
  (void)n;
  for (int64_t i = 0; i < m; i++) y[i] = 0.0;
  for (int64_t i = 0; i < nnz; i++) y[rows[i] - 1] += x[cols[i] - 1];
}
void orc_sbm_tmul(int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const double* x, double* y) {
  (void)m;
  for (int64_t i = 0; i < n; i++) y[i] = 0.0;
  for (int64_t i = 0; i < nnz; i++) y[cols[i] - 1] += x[rows[i] - 1];
}

/* A13. SparseBinMatrixCSR ctor — src/sparsebin_csr.jl:22-37. Stable sort of rows (Julia sortperm's default
 * for integers is a stable merge/counting sort [ext]); m=max(rows), n=max(cols); row_ptr filled with
 * nnz+1 then row_ptr[r] = first 1-based position of row r (empty rows inherit the next start).
 * Outputs are 1-based Int32 exactly like the Julia fields. */
int orc_csr_build(int64_t nnz, const int32_t* rows, const int32_t* cols, int64_t* m_out, int64_t* n_out,
                  int32_t* row_ptr /* ≥ max(rows)+1 */, int32_t* col_ind /* nnz */) {
  int32_t m = 0, n = 0;
  for (int64_t i = 0; i < nnz; i++) { if (rows[i] > m) m = rows[i]; if (cols[i] > n) n = cols[i]; }
  *m_out = m; *n_out = n;
  /* stable counting sort by row = sortperm(rows) */
  int64_t* cnt = (int64_t*)calloc((size_t)m + 2, sizeof(int64_t));
  for (int64_t i = 0; i < nnz; i++) cnt[rows[i] + 1]++;
  for (int32_t r = 1; r <= m + 1; r++) cnt[r] += cnt[r - 1];
  int32_t* rows2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  for (int64_t i = 0; i < nnz; i++) { int64_t d = cnt[rows[i]]++; rows2[d] = rows[i]; col_ind[d] = cols[i]; }
  for (int32_t r = 0; r <= m; r++) row_ptr[r] = (int32_t)(nnz + 1);          /* :27-28 */
  int32_t prev = 0;
  for (int64_t i = 1; i <= nnz; i++)                                            /* :31-36 */
    while (rows2[i - 1] > prev) { prev += 1; row_ptr[prev - 1] = (int32_t)i; }
  free(cnt); free(rows2);
  return 0;
}

/* A_mul_B!(y, ::SparseBinMatrixCSR, x) — src/sparsebin_csr.jl:49-63 (sum in stored order). */
void orc_csr_mul(int64_t m, const int32_t* row_ptr, const int32_t* col_ind, const double* x, double* y) {
  for (int64_t row = 0; row < m; row++) {
    double tmp = 0.0;
    for (int32_t i = row_ptr[row]; i <= row_ptr[row + 1] - 1; i++) tmp += x[col_ind[i - 1] - 1];
    y[row] = tmp;
  }
}

/* A11. AtA_mul_B!(y, F, x, λ): y = Fᵀ(F·x) + λx — src/parallel_cg.jl:7-14 for F = SparseBinMatrix. */
void orc_sbm_ata_mul(int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const double* x,
                     double lambda, double* y, double* tmp /* m */) {
  orc_sbm_mul(m, n, nnz, rows, cols, x, tmp);
  orc_sbm_tmul(m, n, nnz, rows, cols, tmp, y);
  for (int64_t i = 0; i < n; i++) y[i] += lambda * x[i];
}
/* same for a dense column-major F (m×n) — generic method, test/solver.jl:14-20 */
void orc_dense_ata_mul(int64_t m, int64_t n, const double* F, const double* x, double lambda, double* y, double* tmp) {
  for (int64_t i = 0; i < m; i++) tmp[i] = 0.0;
  for (int64_t j = 0; j < n; j++) { double xj = x[j]; for (int64_t i = 0; i < m; i++) tmp[i] += F[IDX(i, j, m)] * xj; }
  for (int64_t j = 0; j < n; j++) { double s = 0.0; for (int64_t i = 0; i < m; i++) s += F[IDX(i, j, m)] * tmp[i]; y[j] = s + lambda * x[j]; }
}

/* A10. cg_AtA — src/parallel_cg.jl:63-94 (operator = SparseBinMatrix COO when F==NULL, else dense F).
 *   tol = tol*norm(b); x=0; r=b; p=r; bkden=0
 *   for iter=1:maxiter: bknum=normsq(r); err=sqrt(bknum); err<tol && return x
 *     if iter>1: bk=bknum/bkden; p = bk*p + r;  bkden=bknum
 *     z=(AᵀA+λI)p; ak=bknum/dot(z,p); x+=ak*p; r-=ak*z
 * Returns the number of operator applications performed. */
int orc_cg_ata(int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const double* F,
               const double* b, double lambda, double tol, int64_t maxiter, double* x) {
  double* r = (double*)malloc(sizeof(double) * (size_t)n * 3 + sizeof(double) * (size_t)m);
  double* p = r + n; double* z = p + n; double* tmp = z + n;
  double nb = 0.0;
  for (int64_t i = 0; i < n; i++) nb += b[i] * b[i];
  tol = tol * sqrt(nb);
  for (int64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = 0.0 + b[i]; p[i] = r[i]; }
  double bkden = 0.0;
  int its = 0;
  for (int64_t iter = 1; iter <= maxiter; iter++) {
    double bknum = 0.0;
    for (int64_t i = 0; i < n; i++) bknum += r[i] * r[i];
    double err = sqrt(bknum);
    if (err < tol) break;
    if (iter > 1) { double bk = bknum / bkden; for (int64_t i = 0; i < n; i++) p[i] = bk * p[i] + r[i]; }
    bkden = bknum;
    if (F) orc_dense_ata_mul(m, n, F, p, lambda, z, tmp);
    else orc_sbm_ata_mul(m, n, nnz, rows, cols, p, lambda, z, tmp);
    its++;
    double zp = 0.0;
    for (int64_t i = 0; i < n; i++) zp += z[i] * p[i];
    double ak = bknum / zp;
    for (int64_t i = 0; i < n; i++) x[i] += ak * p[i];
    for (int64_t i = 0; i < n; i++) r[i] -= ak * z[i];
  }
  free(r);
  return its;
}

/* solve_cg2 — src/parallel_matrix.jl:488-507: one independent cg_AtA per column of rhs, maxiter=size(rhs,1). */
void orc_solve_cg2(int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const double* F,
                   const double* rhs /* n×ncol */, int ncol, double lambda, double tol, int64_t maxiter,
                   double* beta /* n×ncol */, int* iters /* ncol or NULL */, int nthreads) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int c = 0; c < ncol; c++) {
    int it = orc_cg_ata(m, n, nnz, rows, cols, F, rhs + (size_t)c * n, lambda, tol, maxiter, beta + (size_t)c * n);
    if (iters) iters[c] = it;
  }
}

/* A9. solve_full — src/sampling.jl:314-320: (FF + λI) \ rhs (dense LU). */
int orc_solve_full(int64_t n, const double* FF, const double* rhs, int ncol, double lambda, double* out) {
  double* A = (double*)malloc(sizeof(double) * (size_t)n * n);
  memcpy(A, FF, sizeof(double) * (size_t)n * n);
  for (int64_t i = 0; i < n; i++) A[IDX(i, i, n)] += lambda;
  memcpy(out, rhs, sizeof(double) * (size_t)n * ncol);
  int info = orc_lu_solve((int)n, ncol, A, out);
  free(A);
  return info;
}

/* A8 (colouring step). mv = MvNormal(0, inv(PDMat(Symmetric(Λ)))); rand(mv, N) — src/sampling.jl:298-300.
 * PDMats [ext]: PDMat(Λ) holds chol(Λ); inv(::PDMat) = PDMat(inv(chol)) (potri: Σ = inv from the factor),
 * re-factored; rand = chol_lower(Σ)·E. Restated: Σ = inv(Λ) (LU inverse of the symmetrised Λ),
 * C = chol_lower(Σ), out[:,c] = C·E[:,c]. E is D×ncols injected standard normals. */
int orc_color_noise(int D, const double* Lambda, int64_t ncols, const double* E, double* out) {
  size_t dd = (size_t)D * D;
  double* S = (double*)malloc(sizeof(double) * dd * 2);
  double* R = S + dd;
  for (int j = 0; j < D; j++) for (int i = 0; i <= j; i++) { S[IDX(i, j, D)] = Lambda[IDX(i, j, D)]; S[IDX(j, i, D)] = Lambda[IDX(i, j, D)]; }
  int info = orc_inv(D, S);
  if (info) { free(S); return info; }
  info = orc_chol_upper(D, S, R);
  if (info) { free(S); return 1000 + info; }
  for (int64_t c = 0; c < ncols; c++)
    for (int i = 0; i < D; i++) {
      double s = 0.0;
      for (int k = 0; k <= i; k++) s += R[IDX(k, i, D)] * E[(size_t)c * D + k];
      out[(size_t)c * D + i] = s;
    }
  free(S);
  return 0;
}

/* A8. rhs of sample_beta — src/sampling.jl:300:
 *   Ft_y = Ft_mul_B(entity, sample_u_c' + rand(mv,N)') + sqrt(lambda_beta) * rand(mv,numF)'
 * Uc = U .- μ (D×N), N1 = coloured noise D×N, N2 = coloured noise D×numF. Output numF×D col-major.
 * F given as COO SparseBinMatrix (rows ∈ 1..N, cols ∈ 1..numF). */
void orc_beta_rhs_sbm(int D, int64_t N, int64_t numF, int64_t nnz, const int32_t* rows, const int32_t* cols,
                      const double* U, const double* mu, const double* N1, const double* N2, double lambda_beta,
                      double* Ft_y) {
  double sq = sqrt(lambda_beta);
  double* col = (double*)malloc(sizeof(double) * (size_t)N);
  for (int d = 0; d < D; d++) {
    for (int64_t i = 0; i < N; i++) col[i] = (U[(size_t)i * D + d] - mu[d]) + N1[(size_t)i * D + d];
    double* y = Ft_y + (size_t)d * numF;
    orc_sbm_tmul(N, numF, nnz, rows, cols, col, y);
    for (int64_t f = 0; f < numF; f++) y[f] = y[f] + sq * N2[(size_t)f * D + d];
  }
  free(col);
}

/* A14. F_mul_beta — src/RelationData.jl:314-320: uhat = (F·β)ᵀ, F COO binary N×numF, β numF×D → uhat D×N. */
void orc_f_mul_beta_sbm(int D, int64_t N, int64_t numF, int64_t nnz, const int32_t* rows, const int32_t* cols,
                        const double* beta, double* uhat) {
  double* y = (double*)malloc(sizeof(double) * (size_t)N);
  for (int d = 0; d < D; d++) {
    orc_sbm_mul(N, numF, nnz, rows, cols, beta + (size_t)d * numF, y);
    for (int64_t i = 0; i < N; i++) uhat[(size_t)i * D + d] = y[i];
  }
  free(y);
}

/* βᵀβ (D×D) — used by sample_lambda_beta (src/sampling.jl:136-142) and full_lambda_u (src/macau.jl:126-129). */
void orc_btb(int D, int64_t numF, const double* beta, double* BtB) {
  for (int j = 0; j < D; j++)
    for (int i = 0; i < D; i++) {
      double s = 0.0;
      for (int64_t f = 0; f < numF; f++) s += beta[(size_t)i * numF + f] * beta[(size_t)j * numF + f];
      BtB[IDX(i, j, D)] = s;
    }
}

/* A15. sample_lambda_beta — src/sampling.jl:136-142, with the Gamma(b,1) variate g injected:
 *   νx = ν + numF·D; μx = μ·νx/(ν + μ·tr((βᵀβ)Λ)); b = νx/2; c = 2μx/νx; λβ = c·g, g ~ Gamma(b, 1). */
double orc_lambda_beta(int D, int64_t numF, const double* BtB, const double* Lambda, double nu, double mu, double g,
                       double* shape_out) {
  double nux = nu + (double)numF * D;
  double tr = 0.0;
  for (int i = 0; i < D; i++) for (int k = 0; k < D; k++) tr += BtB[IDX(i, k, D)] * Lambda[IDX(k, i, D)];
  double mux = mu * nux / (nu + mu * tr);
  double b = nux / 2.0, c = 2.0 * mux / nux;
  if (shape_out) *shape_out = b;
  return c * g;
}

/* N1. udot / pred — src/sampling.jl:9-51: ŷ_t = Σ_k ∏_m U_m[k, id_m(t)] + mean_value. ids ntest×K col-major 1-based. */
void orc_pred(int D, int K, int64_t ntest, const int64_t* ids, double* const* U, double mean_value, double* out) {
  for (int64_t t = 0; t < ntest; t++) {
    double s = 0.0;
    for (int k = 0; k < D; k++) {
      double pr = 1.0;
      for (int m = 0; m < K; m++) pr *= U[m][(size_t)(ids[(size_t)m * ntest + t] - 1) * D + k];
      s += pr;
    }
    out[t] = s + mean_value;
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
