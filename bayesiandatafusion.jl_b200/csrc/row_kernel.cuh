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
//   4. Λ* = Λ + αG stays in the accumulator registers and is factored there as Λ* = W·Wᵀ with W UPPER triangular
//      ("UL" Cholesky, block rows eliminated from the last to the first): per 8-row panel every warp factors the
//      8×8 diagonal block (shuffle-based, with its inverse), the panel tiles are scaled by one DMMA pair each and
//      parked in shared memory, and the trailing update is again a DMMA syrk with the sign flipped;
//   5. warp 0 runs the two blocked substitutions and stores the draw x = W⁻ᵀ(z + W⁻¹·rhs) — algebraically
//      identical, for the same z, to the reference's chol(inv(Λ*))ᵀ·z + inv(Λ*)·rhs (DESIGN.md §"draw formula").
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
  static constexpr int NB = C::NB;
  static constexpr int PSZ = 32 * NB * NB;               // scaled panels: block row p = 8 × 8p doubles at pitch 8p+4, offset 32p²
  static constexpr int BUFSZ = 2 * KS * S + 2 * KS;
  static constexpr int REGSZ = PSZ > BUFSZ ? PSZ : BUFSZ;  // panels alias the (dead) stage buffers
  static constexpr int SMEM_DOUBLES = REGSZ + 64 + NB * 64 + 4 * DP + 8;
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
    double* Pn = smem;                      // scaled panels, alias the stage buffers after the main loop
    double* Dg = smem + REGSZ;              // current 8×8 diagonal block
    double* Wv = Dg + 64;                   // inverse diagonal blocks W_pp⁻¹ [NB][8][8]
    double* rhs = Wv + NB * 64;             // [DP]
    double* lmu = rhs + DP;                 // Λ·μ [DP]
    double* ys = lmu + DP;                  // W⁻¹·rhs (+ z) [DP]
    double* xs = ys + DP;                   // the draw [DP]
    double* ts = xs + DP;                   // [8]

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

    // ---- Λ* = Λ + αG in the accumulators; rhs = Λμ + α·Σv·r in shared memory; identity on the padding ----------
    const double alpha = p.alpha;
    warp_dispatch(warp, [&](auto w) {
      for_acc<decltype(w)::value>(acc, lane, [&](int, int i, int j, double& v) {
        if (i < D && j < D) {
          v = fma(alpha, v, __ldg(p.Lambda + (i > j ? i + (size_t)j * D : j + (size_t)i * D)));
        } else {
          if (i == D && j < D) rhs[j] = fma(alpha, v, lmu[j]);  // augmented row = Σ v·r
          v = (i == j) ? 1.0 : 0.0;
        }
      });
    });
    if (!aug && tid < D) rhs[tid] = fma(alpha, bsum, lmu[tid]);
    if (tid >= D && tid < DP) rhs[tid] = 0.0;

    // ---- blocked UL factorisation: block rows p = NB-1 … 0 --------------------------------------------------------
    bool bad = false;
    for (int pb = NB - 1; pb >= 0; pb--) {
      double* Pp = Pn + 32 * pb * pb;
      const int SP = 8 * pb + 4;
      // (a) the warp that owns the diagonal tile (pb,pb) factors it in registers, in the DMMA accumulator layout
      //     (lane = 4·row + q holds columns 2q, 2q+1): A_pp = W·Wᵀ with W upper, by elimination from the last column
      //     to the first, carrying an identity block that ends up as W⁻¹. Owners of the panel tiles (pb,J<pb) park
      //     their raw tiles in shared memory meanwhile.
      double a0 = 0.0, a1 = 0.0;
      bool own = false;
      warp_dispatch(warp, [&](auto w) {
        constexpr int W = decltype(w)::value;
        static_for<C::ntiles(W)>([&](auto t) {
          constexpr int T = decltype(t)::value;
          using ti = TI<C, W, T>;
          if (ti::I == pb) {
            if (ti::J == ti::I) {
              a0 = acc[T][0];
              a1 = acc[T][1];
              own = true;
            } else {
              *reinterpret_cast<double2*>(Pp + (lane >> 2) * SP + 8 * ti::J + 2 * (lane & 3)) = make_double2(acc[T][0], acc[T][1]);
            }
          }
        });
      });
      if (own) {
        const int r = lane >> 2, q = lane & 3;
        double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 7; j >= 0; j--) {
          const int jq = j >> 1;
          const double sel = (j & 1) ? a1 : a0;
          const double piv = __shfl_sync(0xffffffffu, sel, 4 * j + jq);
          if (!(piv > 0.0)) bad = true;
          const double sc = rsqrt(piv);
          const double wij = __shfl_sync(0xffffffffu, sel, (lane & ~3) | jq) * sc;  // w_ij of this lane's row
          const double rj0 = __shfl_sync(0xffffffffu, a0, 4 * j + q) * sc;          // w_kj for this lane's two columns
          const double rj1 = __shfl_sync(0xffffffffu, a1, 4 * j + q) * sc;
          const double ej0 = __shfl_sync(0xffffffffu, e0, 4 * j + q) * sc;          // row j of W⁻¹
          const double ej1 = __shfl_sync(0xffffffffu, e1, 4 * j + q) * sc;
          if (r < j) {
            a0 = fma(-wij, rj0, a0);
            a1 = fma(-wij, rj1, a1);
            e0 = fma(-wij, ej0, e0);
            e1 = fma(-wij, ej1, e1);
          } else if (r == j) {
            e0 = ej0;
            e1 = ej1;
          }
        }
        *reinterpret_cast<double2*>(Wv + pb * 64 + r * 8 + 2 * q) = make_double2(e0, e1);
      }
      __syncthreads();
      // (b)
      // scale this warp's panel tiles: R_pJ = W_pp⁻¹·A_pJ (= W_Jpᵀ), one DMMA pair per tile, written back in place
      {
        const double wa0 = Wv[pb * 64 + (lane >> 2) * 8 + (lane & 3)];
        const double wa1 = Wv[pb * 64 + (lane >> 2) * 8 + 4 + (lane & 3)];
        warp_dispatch(warp, [&](auto w) {
          constexpr int W = decltype(w)::value;
          static_for<C::ntiles(W)>([&](auto t) {
            constexpr int T = decltype(t)::value;
            using ti = TI<C, W, T>;
            if (ti::I == pb && ti::J < ti::I) {
              const double b0 = Pp[(lane & 3) * SP + 8 * ti::J + (lane >> 2)];
              const double b1 = Pp[(4 + (lane & 3)) * SP + 8 * ti::J + (lane >> 2)];
              double c2[2] = {0.0, 0.0};
              dmma884(c2, wa0, b0);
              dmma884(c2, wa1, b1);
              *reinterpret_cast<double2*>(Pp + (lane >> 2) * SP + 8 * ti::J + 2 * (lane & 3)) = make_double2(c2[0], c2[1]);
            }
          });
        });
      }
      __syncthreads();
      // (c) trailing update of the tiles above the panel: A_IJ −= R_pIᵀ·R_pJ  (I, J < pb)
      if (pb > 0) {
        warp_dispatch(warp, [&](auto w) {
          constexpr int W = decltype(w)::value;
          constexpr int NF = C::nfrag(W);
          constexpr int NTW = C::ntiles(W);
          if constexpr (NTW > 0) {
            const double* base = Pp + (lane & 3) * SP + (lane >> 2);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              double f[NF];
              static_for<NF>([&](auto r) {
                constexpr int R = decltype(r)::value;
                constexpr int blk = FI<C, W, R>::blk;
                f[R] = (blk < pb) ? base[h * 4 * SP + 8 * blk] : 0.0;
              });
              static_for<NTW>([&](auto t) {
                constexpr int T = decltype(t)::value;
                using ti = TI<C, W, T>;
                if (ti::I < pb) dmma884(acc[T], -f[ti::fa], f[ti::fb]);
              });
            }
          }
        });
      }
    }
    if (bad && lane == 0) atomicOr(p.err_flag, 1);
    __syncthreads();

    // ---- substitutions by warp 0: y = W⁻¹·rhs (last block first), x = W⁻ᵀ(y + z) (first block first) --------------
    if (warp == 0) {
      const int r8 = lane & 7;
      const int64_t grow = (int64_t)lrow * p.world + p.rank;  // global 0-based row id
      for (int J = NB - 1; J >= 0; J--) {
        const double* Pj = Pn + 32 * J * J;
        const int SP = 8 * J + 4;
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          y0 = fma(Wv[J * 64 + r8 * 8 + k], rhs[8 * J + k], y0);
          y1 = fma(Wv[J * 64 + r8 * 8 + k + 1], rhs[8 * J + k + 1], y1);
        }
        if (lane < 8) ys[8 * J + r8] = y0 + y1;
        __syncwarp();
        for (int c = lane; c < 8 * J; c += 32) {
          double s0 = rhs[c], s1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            s0 = fma(-Pj[k * SP + c], ys[8 * J + k], s0);
            s1 = fma(-Pj[(k + 1) * SP + c], ys[8 * J + k + 1], s1);
          }
          rhs[c] = s0 + s1;
        }
        __syncwarp();
      }
      for (int c = lane; c < DP; c += 32) {
        double z = 0.0;
        if (c < D) z = p.Z ? __ldg(p.Z + (size_t)slot * p.ld + c) : philox_normal(p.seed, p.sweep, p.entity, grow, c);
        ys[c] += z;
      }
      __syncwarp();
      for (int I = 0; I < NB; I++) {
        const double* Pi = Pn + 32 * I * I;
        const int SP = 8 * I + 4;
        const int k = lane >> 2, q = lane & 3;
        double s0 = 0.0, s1 = 0.0;
        for (int c = q; c < 8 * I; c += 8) {
          s0 = fma(Pi[k * SP + c], xs[c], s0);
          s1 = fma(Pi[k * SP + c + 4], xs[c + 4], s1);
        }
        double sacc = s0 + s1;
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (q == 0) ts[k] = ys[8 * I + k] - sacc;
        __syncwarp();
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 8; kk += 2) {
          x0 = fma(Wv[I * 64 + kk * 8 + r8], ts[kk], x0);
          x1 = fma(Wv[I * 64 + (kk + 1) * 8 + r8], ts[kk + 1], x1);
        }
        if (lane < 8) xs[8 * I + r8] = x0 + x1;
        __syncwarp();
      }
      double* out = p.Uout + (size_t)slot * p.ld;
      for (int j = lane; j < p.ld; j += 32) out[j] = j < D ? xs[j] : 0.0;
    }
  }
};

template <class K>
__global__ void __launch_bounds__(K::NTHR, (K::NW == 1 ? 16 : (K::NW == 4 ? 4 : 2))) row_kernel(const RowParams p) {
  extern __shared__ __align__(16) double smem_dyn[];
  K::run(p, smem_dyn);
}

}  // namespace bdf
