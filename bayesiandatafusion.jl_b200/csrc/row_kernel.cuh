// The per-entity conditional draw (reference: sample_user_basic, src/sampling.jl:200-234, driven by
// sample_latent_all2!/sample_latent_range, :149-198) as one fused sm_100a kernel.
//
// One CTA (NW warps) owns one work item = (row, chunk of that row's observations):
//   1. gather the partner factor rows of the chunk (Hadamard product of two partners for 3-mode tensors) into a
//      double-buffered shared-memory tile, KS observations per stage, software-prefetched through registers;
//   2. accumulate the lower triangle of G = Σ v vᵀ with FP64 tensor-core MMAs (DMMA.8x8x4, mma.sync m8n8k4 f64),
//      8×8 tiles dealt to the warps at compile time (tiles.cuh); the rhs Σ v·r rides along as one more column
//      of the tile when D is not a multiple of 8, else it is a DFMA side-sum;
//   3. rows split over several CTAs park their partial in a workspace; the last CTA to arrive adds the partials
//      in chunk order (deterministic);
//   4. Λ* = Λ + αG is written index-REVERSED into shared memory, factored in place (LDLᵀ-style Cholesky with the
//      rhs as an extra row = forward substitution for free), back-substituted by one warp, and the draw
//      x = W⁻ᵀ(z + W⁻¹·rhs), Λ* = W·Wᵀ (W upper) is stored — algebraically identical, for the same z, to the
//      reference's chol(inv(Λ*))ᵀ·z + inv(Λ*)·rhs (DESIGN.md §"draw formula").
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <utility>

#include "philox.cuh"
#include "tiles.cuh"

namespace bdf {

struct RowParams {
  // work list (cost-descending)
  const int32_t* item_row;    // local row index (0-based within this rank's shard)
  const int64_t* item_beg;    // first observation of the chunk (index into col/val)
  const int32_t* item_len;    // observations in the chunk
  const int32_t* item_split;  // -1 = row handled by this CTA alone, else split-row id
  const int32_t* item_chunk;  // chunk number within the split row
  const int32_t* split_nchunks;
  const int64_t* split_wsoff;  // first partial slot of the split row
  int* split_counter;          // arrival counters (self-resetting)
  double* ws;                  // partial workspace
  // observations of this mode (CSR payload); col* hold SLOT indices of the partner rows
  const int32_t* col0;
  const int32_t* col1;
  const double* val;
  const double* P0;  // partner factor buffers (slot-major, ld doubles per row)
  const double* P1;
  int ld;        // row pitch (doubles) of every factor buffer
  double* Uout;  // factor buffer being sampled
  int64_t slot_base;  // slot of local row 0 (= rank * Nper)
  const double* Lambda;  // D×D column-major
  const double* mu;      // D (mu_ld == 0) or slot-major matrix (mu_ld == ld)
  int64_t mu_ld;
  const double* Z;  // injected normals, slot-major (ld pitch), or nullptr → Philox
  double alpha, mean;
  int D;
  int rank, world;
  uint64_t seed, sweep;
  int entity;
  int* err_flag;
};

template <int N, class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl<N>(f, std::make_integer_sequence<int, N>{});
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__host__ __device__ constexpr int tri(int i) { return i * (i + 1) / 2; }

template <int DP_, int NW_, bool TENSOR_>
struct RowKernel {
  static constexpr int DP = DP_;
  static constexpr int NW = NW_;
  static constexpr bool TENSOR = TENSOR_;
  using C = TileCfg<DP, NW>;
  static constexpr int NTHR = NW * 32;
  static constexpr int S = DP + 4;  // smem row pitch: ≡ 4 or 12 (mod 16) doubles → conflict-free DMMA fragment loads
  static constexpr int OPP = NTHR / 16;                  // observations fetched per pass (16 threads per row)
  static constexpr int PASSES = TENSOR ? (NW == 8 ? 1 : (NW == 4 ? 2 : 4)) : (NW == 8 ? 2 : (NW == 4 ? 4 : 8));
  static constexpr int KS = OPP * PASSES;                // observations per stage
  static constexpr int JP = (DP / 2 + 15) / 16;          // 16-byte pieces per thread per observation
  static constexpr int TPW = C::TPW;
  static constexpr int PST = NW * TPW * 64 + DP;         // doubles per parked partial
  static constexpr int MSZ = tri(DP + 1) + DP + 1;       // packed lower triangle incl. rhs row
  static constexpr int BUFSZ = 2 * KS * S + 2 * KS;
  static constexpr int SMEM_DOUBLES = (MSZ > BUFSZ ? MSZ : BUFSZ) + 2 * DP;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;

  struct Pre {
    double2 a[PASSES][JP];
    double r[PASSES];
  };

  // ---- gather one stage into registers -------------------------------------------------------------------
  static __device__ __forceinline__ void prefetch(const RowParams& p, int64_t o0, int64_t oend, int tr, int tq, Pre& pre) {
    const int ppr = p.ld >> 1;
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++) {
      const int64_t o = o0 + tr + ps * OPP;
      const bool ok = o < oend;
      int c0 = 0, c1 = 0;
      double v = 0.0;
      if (ok) {
        c0 = __ldg(p.col0 + o);
        if (TENSOR) c1 = __ldg(p.col1 + o);
        v = __ldg(p.val + o);
      }
      pre.r[ps] = ok ? v - p.mean : 0.0;
      const double2* r0 = reinterpret_cast<const double2*>(p.P0 + (size_t)c0 * p.ld);
      const double2* r1 = TENSOR ? reinterpret_cast<const double2*>(p.P1 + (size_t)c1 * p.ld) : nullptr;
#pragma unroll
      for (int j = 0; j < JP; j++) {
        const int pc = tq + 16 * j;
        double2 x = make_double2(0.0, 0.0);
        if (ok && pc < ppr) {
          x = __ldg(r0 + pc);
          if (TENSOR) {
            const double2 y = __ldg(r1 + pc);
            x.x *= y.x;
            x.y *= y.y;
          }
        }
        pre.a[ps][j] = x;
      }
    }
  }

  static __device__ __forceinline__ void store_stage(const int D, double* buf, double* rs, int tr, int tq, const Pre& pre, bool aug) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++) {
      const int k = tr + ps * OPP;
#pragma unroll
      for (int j = 0; j < JP; j++) {
        const int pc = tq + 16 * j;
        if (pc < DP / 2) {
          double2 x = pre.a[ps][j];
          if (aug) {
            if (2 * pc == D) x.x = pre.r[ps];
            if (2 * pc + 1 == D) x.y = pre.r[ps];
          }
          *reinterpret_cast<double2*>(buf + k * S + 2 * pc) = x;
        }
      }
      if (tq == 0) rs[k] = pre.r[ps];
    }
  }

  // ---- DMMA accumulate of one stage, specialised per warp -------------------------------------------------
  template <int W>
  static __device__ __forceinline__ void compute(double (&acc)[TPW][2], const double* buf, int nk4, int lane) {
    constexpr int NF = C::nfrag(W);
    constexpr int NTW = C::ntiles(W);
    if constexpr (NTW > 0) {
      const double* base = buf + (lane & 3) * S + (lane >> 2);
      for (int k4 = 0; k4 < nk4; k4++) {
        double f[NF];
        static_for<NF>([&](auto r) {
          constexpr int R = decltype(r)::value;
          f[R] = base[k4 * 4 * S + 8 * FI<C, W, R>::blk];
        });
        static_for<NTW>([&](auto t) {
          constexpr int T = decltype(t)::value;
          using ti = TI<C, W, T>;
          dmma884(acc[T], f[ti::fa], f[ti::fb]);
        });
      }
    }
  }

  // visit every accumulator element of warp W: f(T, i, j, value&) with (i, j) the Gram-matrix coordinates
  template <int W, class F>
  static __device__ __forceinline__ void for_acc(double (&acc)[TPW][2], int lane, F&& fn) {
    constexpr int NTW = C::ntiles(W);
    static_for<NTW>([&](auto t) {
      constexpr int T = decltype(t)::value;
      using ti = TI<C, W, T>;
      const int i = 8 * ti::I + (lane >> 2);
      const int j = 8 * ti::J + 2 * (lane & 3);
      fn(T, i, j, acc[T][0]);
      fn(T, i, j + 1, acc[T][1]);
    });
  }

  template <class F>
  static __device__ __forceinline__ void warp_dispatch(int warp, F&& fn) {
    static_for<NW>([&](auto w) {
      constexpr int W = decltype(w)::value;
      if (warp == W) fn(w);
    });
  }

  // ---- the kernel body -------------------------------------------------------------------------------------
  static __device__ void run(const RowParams& p, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = tid >> 4, tq = tid & 15;
    const int D = p.D;
    const bool aug = D < DP;
    const int item = blockIdx.x;
    const int lrow = p.item_row[item];
    const int64_t obeg = p.item_beg[item];
    const int len = p.item_len[item];
    const int64_t oend = obeg + len;
    const int split = p.item_split[item];
    const int64_t slot = p.slot_base + lrow;

    double* bufs = smem;                    // [2][KS*S]
    double* rss = smem + 2 * KS * S;        // [2][KS]
    double* M = smem;                       // packed lower triangle, aliases the stage buffers after the main loop
    double* lmu = smem + (MSZ > BUFSZ ? MSZ : BUFSZ);  // Λ·μ  [DP]
    double* xs = lmu + DP;                                // solution scratch [DP]

    double acc[TPW][2];
#pragma unroll
    for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
    double bsum = 0.0;  // Σ v_tid · r  (used when !aug)

    // Λ·μ for this row (thread j < D), independent of the gather
    {
      const double* mu = p.mu + (p.mu_ld ? slot * p.mu_ld : 0);
      if (tid < DP) {
        double s = 0.0;
        if (tid < D)
          for (int i = 0; i < D; i++) s = fma(__ldg(p.Lambda + tid + (size_t)i * D), __ldg(mu + i), s);
        lmu[tid] = s;
      }
    }

    const int nst = (len + KS - 1) / KS;
    Pre pre;
    if (nst > 0) {
      prefetch(p, obeg, oend, tr, tq, pre);
      store_stage(D, bufs, rss, tr, tq, pre, aug);
    }
    __syncthreads();
    for (int s = 0; s < nst; s++) {
      const double* buf = bufs + (s & 1) * KS * S;
      const double* rs = rss + (s & 1) * KS;
      const bool more = s + 1 < nst;
      if (more) prefetch(p, obeg + (int64_t)(s + 1) * KS, oend, tr, tq, pre);
      int rem = len - s * KS;
      if (rem > KS) rem = KS;
      const int nk4 = (rem + 3) >> 2;
      warp_dispatch(warp, [&](auto w) { compute<decltype(w)::value>(acc, buf, nk4, lane); });
      if (!aug && tid < DP) {
        for (int k = 0; k < nk4 * 4; k++) bsum = fma(buf[k * S + tid], rs[k], bsum);
      }
      if (more) store_stage(D, bufs + ((s + 1) & 1) * KS * S, rss + ((s + 1) & 1) * KS, tr, tq, pre, aug);
      __syncthreads();
    }

    // ---- split rows: park the partial, last arriver reduces in chunk order -------------------------------
    if (split >= 0) {
      const int nch = p.split_nchunks[split];
      double* part = p.ws + (size_t)(p.split_wsoff[split] + p.item_chunk[item]) * PST;
#pragma unroll
      for (int t = 0; t < TPW; t++)
        *reinterpret_cast<double2*>(part + ((size_t)(warp * TPW + t) * 32 + lane) * 2) = make_double2(acc[t][0], acc[t][1]);
      if (tid < DP) part[NW * TPW * 64 + tid] = bsum;
      __threadfence();
      __syncthreads();
      __shared__ int s_last;
      if (tid == 0) {
        const int old = atomicAdd(p.split_counter + split, 1);
        s_last = (old == nch - 1);
        if (s_last) p.split_counter[split] = 0;
      }
      __syncthreads();
      if (!s_last) return;
      __threadfence();
#pragma unroll
      for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
      bsum = 0.0;
      const double* base = p.ws + (size_t)p.split_wsoff[split] * PST;
      for (int c = 0; c < nch; c++) {
        const double* pc = base + (size_t)c * PST;
#pragma unroll
        for (int t = 0; t < TPW; t++) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(pc + ((size_t)(warp * TPW + t) * 32 + lane) * 2));
          acc[t][0] += v.x;
          acc[t][1] += v.y;
        }
        if (tid < DP) bsum += __ldcg(pc + NW * TPW * 64 + tid);
      }
    }

    // ---- Λ* (index-reversed, packed lower) and rhs row into shared memory ---------------------------------
    const double alpha = p.alpha;
    warp_dispatch(warp, [&](auto w) {
      for_acc<decltype(w)::value>(acc, lane, [&](int, int i, int j, double& v) {
        if (j <= i) {
          if (i < D) {
            const int a = D - 1 - j, b = D - 1 - i;
            M[tri(a) + b] = fma(alpha, v, __ldg(p.Lambda + i + (size_t)j * D));
          } else if (i == D && j < D) {
            M[tri(D) + (D - 1 - j)] = fma(alpha, v, lmu[j]);
          }
        }
      });
    });
    if (!aug && tid < D) M[tri(D) + (D - 1 - tid)] = fma(alpha, bsum, lmu[tid]);

    // ---- in-place factorisation: unscaled columns l~_ij = L_ij·sqrt(d_j), pivots d_j on the diagonal ----------
    constexpr int TY = NTHR / 16;
    bool bad = false;
    for (int j = 0; j < D; j++) {
      __syncthreads();
      const double d = M[tri(j) + j];
      if (!(d > 0.0)) bad = true;
      const double id = 1.0 / d;
      for (int i = j + 1 + tr; i <= D; i += TY) {
        const double lij = M[tri(i) + j] * id;
        const int kmax = i < D ? i : D - 1;
        double* Mi = M + tri(i);
        for (int k = j + 1 + tq; k <= kmax; k += 16) Mi[k] = fma(-lij, M[tri(k) + j], Mi[k]);
      }
    }
    __syncthreads();
    if (bad && tid == 0) atomicOr(p.err_flag, 1);

    // ---- back substitution by warp 0:  x'_j = (g_j − Σ_{i>j} l~_ij x'_i) / d_j,  g_j = l~_Dj + z'_j·sqrt(d_j) -----
    if (warp == 0) {
      constexpr int NS = (DP + 31) / 32;
      double g[NS], rd[NS];
      const int64_t grow = (int64_t)lrow * p.world + p.rank;  // global 0-based row id
#pragma unroll
      for (int s = 0; s < NS; s++) {
        const int j = lane + 32 * s;
        g[s] = 0.0;
        rd[s] = 0.0;
        if (j < D) {
          const double d = M[tri(j) + j];
          const int jo = D - 1 - j;  // original latent index
          const double z = p.Z ? __ldg(p.Z + (size_t)slot * p.ld + jo) : philox_normal(p.seed, p.sweep, p.entity, grow, jo);
          g[s] = fma(z, sqrt(d), M[tri(D) + j]);
          rd[s] = 1.0 / d;
        }
      }
      for (int i = D - 1; i >= 0; i--) {
        const int ol = i & 31, os = i >> 5;
        double xi = 0.0;
#pragma unroll
        for (int s = 0; s < NS; s++)
          if (s == os) xi = g[s] * rd[s];
        xi = __shfl_sync(0xffffffffu, xi, ol);
        const double* Mi = M + tri(i);
#pragma unroll
        for (int s = 0; s < NS; s++) {
          const int j = lane + 32 * s;
          if (j < i) g[s] = fma(-Mi[j], xi, g[s]);
        }
        if (lane == ol) xs[i] = xi;
      }
      __syncwarp();
      double* out = p.Uout + (size_t)slot * p.ld;
      for (int j = lane; j < p.ld; j += 32) out[j] = j < D ? xs[D - 1 - j] : 0.0;
    }
  }
};

template <class K>
__global__ void __launch_bounds__(K::NTHR, (K::NW == 1 ? 16 : (K::NW == 4 ? 4 : 2))) row_kernel(const RowParams p) {
  extern __shared__ __align__(16) double smem_dyn[];
  K::run(p, smem_dyn);
}

}  // namespace bdf
