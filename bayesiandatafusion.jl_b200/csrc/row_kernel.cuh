// The per-entity conditional draw (reference: sample_user_basic, src/sampling.jl:200-234, driven by
// sample_latent_all2!/sample_latent_range, :149-198) as one fused sm_100a kernel.
//
// One CTA (NW warps) owns one work item = (row, chunk of that row's observations):
//   1. gather the partner factor rows of the chunk into a ring of shared-memory stages (2 × 28 observations for the 4-warp CTAs, 2 × 24 for
//      the one-warp CTAs) with TMA 1-D bulk copies — one cp.async.bulk (UBLKCP) per partner row, completion counted in bytes on the
//      stage's mbarrier; the partner slots / values of a stage are fetched one stage ahead with cp.async into a small shared-memory
//      table; one barrier per stage; for 3-mode tensors both partners are staged and multiplied while forming the fragments;
//   2. accumulate the lower triangle of G = Σ v vᵀ with FP64 tensor-core MMAs (DMMA.8x8x4, mma.sync m8n8k4 f64),
//      8×8 tiles dealt to the warps at compile time (tiles.cuh); the rhs Σ v·r rides along as one more column
//      of the tile when D is even and not a multiple of 8 ("aug"), else it is a DFMA side-sum;
//   3. rows split over several CTAs park their partial in a workspace; the last CTA to arrive adds the partials
//      in chunk order, in two levels for rows with many chunks (deterministic);
//   4. the Gram tiles are parked in shared memory (8×8 blocks in a swizzled layout that makes both the accumulator-layout writes and
//      the transposed fragment reads conflict-free), Λ* = Λ + αG is formed there and factored
//      as Λ* = W·Wᵀ with W UPPER triangular ("UL" Cholesky, block rows eliminated from the last to the first): per
//      8-row panel one warp factors the 8×8 diagonal block with shuffles (its inverse then replaces the diagonal tile), the panel
//      tiles are scaled by one DMMA pair each, the trailing update is a DMMA syrk with the sign flipped, and the diagonal warp factors
//      the next diagonal block while the other warps finish the update (look-ahead); y = W⁻¹·rhs rides along;
//   5. one warp runs the forward substitution and stores the draw x = W⁻ᵀ(z + W⁻¹·rhs) — algebraically identical,
//      for the same z, to the reference's chol(inv(Λ*))ᵀ·z + inv(Λ*)·rhs (DESIGN.md §"draw formula") — into its own replica and,
//      with several GPUs, into every peer's (fused all-gather).
// The body is four pieces (syrk_item, split_reduce, park_tiles, factor_and_draw) parameterised on the group-wide barrier; the opt-in
// persistent warp-specialised kernel of row_kernel_ws.cuh is built from the same pieces.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <utility>

#include "philox.cuh"
#include "tiles.cuh"

// build-time knobs (tools/ab.sh builds variants with -D...)
#ifndef BDF_K4_UNROLL
#define BDF_K4_UNROLL 1
#endif
#ifndef BDF_KS1
#define BDF_KS1 24   // observations per gather stage of the one-warp CTAs (D <= 32): two 24-row stages, same shared memory as three of 16 (+3 % at D=32)
#endif
#ifndef BDF_KS4
#define BDF_KS4 28   // observations per gather stage of the 4-warp CTAs: two 28-row stages beat three 16-row ones (fewer barrier / issue phases per observation; +2.5 % at D=100, +5 % at D=64, +10 % at D=48); 28 is the most that keeps four D=100 CTAs per SM
#endif
#ifndef BDF_KS4S
#define BDF_KS4S 28  // the same for 32 < D <= 64 (six or more CTAs per SM)
#endif
#ifndef BDF_MINB4S
#define BDF_MINB4S 6 // CTAs per SM the 4-warp kernel is compiled for at 32 < D <= 64
#endif
#ifndef BDF_NBUF4
#define BDF_NBUF4 2  // stages in their ring
#endif
#ifndef BDF_GATHER_LDGSTS
#define BDF_GATHER_LDGSTS 0  // 1: the 4-warp CTAs gather with 16-byte cp.async (LDGSTS) instead of one TMA bulk copy per row (A/B knob)
#endif
#ifndef BDF_GATHER_LDGSTS1
#define BDF_GATHER_LDGSTS1 0  // 1: the one-warp CTAs (D <= 32) gather with cp.async, 16 lanes per row, two rows per instruction
#endif
#ifndef BDF_K4_UNROLL4
#define BDF_K4_UNROLL4 2  // 4-warp CTAs (D > 32): two k-steps per loop trip — the fragment loads of the second overlap the DMMAs of the first (+1 % on C2)
#endif
#ifndef BDF_FD_UNROLL
#define BDF_FD_UNROLL 8
#endif
#ifndef BDF_NBUF1
#define BDF_NBUF1 2
#endif
#ifndef BDF_MINB1
#define BDF_MINB1 16
#endif
#ifndef BDF_VWARP
#define BDF_VWARP 1
#endif
#ifndef BDF_HALLEY
#define BDF_HALLEY 1
#endif
#ifndef BDF_BS_SWITCH
#define BDF_BS_SWITCH 1000
#endif
#ifndef BDF_UR_UNROLL
#define BDF_UR_UNROLL 2
#endif

namespace bdf {

#define BDF_MAX_USES 6
struct RelTab {
  // CSR payload of the entity's mode in one relation; col* hold SLOT indices of the partner rows (col1 == nullptr in a launch
  // of the 3-mode kernel: a 2-mode relation, the second partner is the all-ones row P1)
  const int32_t* col0;
  const int32_t* col1;
  const double* val;
  const double* P0;  // partner factor buffers (slot-major, ld doubles per row)
  const double* P1;
  double alpha, mean;
};

struct RowParams {
  // work list (cost-descending)
  const int32_t* item_row;    // local row index (0-based within this rank's shard)
  const int64_t* item_beg;    // first observation of the chunk (index into col/val)
  const int32_t* item_len;    // observations in the chunk
  const int32_t* item_split;  // -1 = row handled by this CTA alone, else split-row id
  const int32_t* item_chunk;  // chunk number within the split row
  const int32_t* split_nchunks;
  const int64_t* split_wsoff;  // first partial slot of the split row
  const int32_t* split_gsize;  // chunks per reduction group of the split row
  const int64_t* split_gcoff;  // first group counter / group-partial slot of the split row
  int64_t gslot_base;          // workspace slot of group partial 0 (the group partials follow all chunk partials)
  int* group_counter;          // arrival counters of the groups (self-resetting)
  int* split_counter;          // arrival counters of the rows' groups (self-resetting)
  double* ws;                  // partial workspace
  // observations of this mode, one table entry per relation the entity takes part in (src/sampling.jl:266-283 sums the
  // relations' contributions); item_rel[item] picks the entry, nullptr = entry 0 for every item
  RelTab rt[BDF_MAX_USES];
  const int32_t* item_rel;
  int ld;        // row pitch (doubles) of every factor buffer
  double* Uout;  // factor buffer being sampled
  double* peer_out[8];  // fused all-gather: the same buffer on every OTHER rank (IPC-mapped peer memory, NVLink stores), nullptr-terminated
  int64_t slot_base;  // slot of local row 0 (= rank * Nper)
  const int32_t* row_of_slot;  // global 0-based row of each slot (nnz-balanced shards) or nullptr = cyclic: row = local·world + rank
  const double* Lambda;  // D×D column-major
  const double* LT;      // Λ in tile order (64 doubles per tile t = tri(I)+J, identity on the padding), see prep_lambda_kernel
  const double* lmu;     // Λ·μ (DP doubles) when μ is shared by all rows, else nullptr
  const double* mu;      // D (mu_ld == 0) or slot-major matrix (mu_ld == ld)
  int64_t mu_ld;
  const double* Z;  // injected normals, slot-major (ld pitch), or nullptr → Philox
  int D;
  int rank, world;
  uint64_t seed, sweep;
  int entity;
  int* err_flag;
  int n_items;
  int* work_counter;  // work queue head of the persistent warp-specialised kernel (row_kernel_ws.cuh), zeroed before each launch
#ifdef BDF_DEBUG  // profiling builds only (tools/phase_probe*.py): the release library carries neither the fields nor the code
  int flags;       // bit 0 = return after the syrk, bit 1 = no DMMAs, bit 2 = no gather (timing experiments only; results invalid)
  long long* dbg;  // optional per-item phase clocks [n_items][8] (bdf_debug_phase_clocks), else nullptr
#endif
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// makes the mbarrier track the completion of all prior cp.async of this thread; counts as one of the barrier's expected arrivals
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(a), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
// TMA 1-D bulk copy global → shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// the rhs Σ v·r is carried as tile column D when that column is free and 16-byte pieces never straddle it
__host__ __device__ constexpr bool use_aug(int D) { return (D & 1) == 0 && (D & 7) != 0; }

// 1/sqrt(x) for x in the normal range: single-precision seed + two Newton steps in double (≈1 ulp); the pivots of a
// positive-definite Λ* are far from the denormal/overflow range the library routine's slow path guards against.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RSQ64H seed (≈ 2^-20 relative), no conversions
#if BDF_HALLEY
  // one cubically convergent step (2^-20 → 2^-60): e = 1 − x·y², y ← y + y·e·(1/2 + 3e/8); dependent depth 4 instead of 6
  const double e = fma(-x, y * y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
#else
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
#endif
}

__host__ __device__ constexpr int tri(int i) { return i * (i + 1) / 2; }

// ---- shared-memory layout of an 8×8 tile (64 doubles): element (k, m) lives at 8k + (m ^ 4·bit1(k)), i.e. rows 2, 3, 6, 7 have
// their two 4-column halves swapped. A tile is written in the DMMA accumulator layout (lane = 4·row + q holds columns 2q, 2q+1: one
// double2 per lane, 16 contiguous doubles per quarter-warp) and read back TRANSPOSED as a DMMA operand fragment (lane = 4·m + k
// holds element (k, m)). Without the swap the fragment read of a half-warp touches rows k = 0…3 × columns m = 0…3 = words
// {0-3, 8-11, 16-19, 24-27}, two rows per bank group: a 2-way conflict on every LDS.64 of the factorisation (ncu: 25 % of all
// shared-memory wavefronts of the kernel). With the swap rows 2, 3 move to words {20-23, 28-31} and every half-warp is conflict-free.
__device__ __forceinline__ int tile_acc_off(int lane) { return (2 * lane) ^ (((lane >> 3) & 1) << 2); }                 // (row lane/4, cols 2q, 2q+1)
__device__ __forceinline__ int tile_frag_off(int lane) { return 8 * (lane & 3) + ((lane >> 2) ^ ((lane & 2) << 1)); }   // element (k = lane%4, m = lane/4); k+4: +32
__device__ __forceinline__ int tile_el(int k, int m) { return 8 * k + (m ^ ((k & 2) << 1)); }
// inverse of t = tri(I) + J, 0 <= J <= I (t < 2^20)
__device__ __forceinline__ void tri_coords(int t, int& I, int& J) {
  int i = (int)((sqrtf((float)(8 * t + 1)) - 1.0f) * 0.5f);
  if (tri(i + 1) <= t) i++;
  if (tri(i) > t) i--;
  I = i;
  J = t - tri(i);
}


// group-wide synchronisation of the code below: the whole CTA (one row per CTA) or one 4-warp group of the warp-specialised
// persistent kernel (row_kernel_ws.cuh), which uses a named barrier per group
struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct GroupSync {
  int id, nthreads;
  __device__ __forceinline__ void operator()() const { named_bar(id, nthreads); }
};

// one work item, as every thread of its group sees it
struct RowCtx {
  int item, lrow, len, split;
  int64_t obeg, slot;
  const RelTab* rt;
  double alpha_f;  // a row fed by one work item scales its Gram matrix by α when parking it; the partials of a split row are parked already scaled
};

template <int DP_, int NW_, bool TENSOR_, int KS_ = 0, int NBUF_ = 0>
struct RowKernel {
  static constexpr int DP = DP_;
  static constexpr int NW = NW_;
  static constexpr bool TENSOR = TENSOR_;
  using C = TileCfg<DP, NW>;
  static constexpr int NTHR = NW * 32;
  static constexpr int S = DP + 4;  // smem row pitch: ≡ 4 or 12 (mod 16) doubles → conflict-free DMMA fragment loads
  static constexpr int OPP = NTHR / 16;          // observations fetched per pass (16 threads per factor row)
  static constexpr int JP = (DP / 2 + 15) / 16;  // 16-byte pieces per thread per observation
  static constexpr int TPW = C::TPW;
  static constexpr int NB = C::NB;
  static constexpr int PST = NW * TPW * 64 + DP;  // doubles per parked partial
  // gather ring of the row kernel
  static constexpr int GP = NW == 8 ? 1 : (NW == 4 ? 2 : 8);  // passes per stage
  static constexpr int KS = KS_ > 0 ? KS_ : (NW == 4 ? (DP <= 64 ? BDF_KS4S : BDF_KS4) : (NW == 1 ? BDF_KS1 : OPP * GP));  // observations per stage (16)
  static constexpr int NBUF = NBUF_ > 0 ? NBUF_ : (NW == 1 ? BDF_NBUF1 : (NW == 4 ? BDF_NBUF4 : 3));
  static constexpr int PF = NBUF - 1;                          // stages in flight ahead of the one being consumed
  // gather by cp.async instead of TMA: TPR threads per partner row, NTHR/TPR rows per pass
  static constexpr bool LDGSTS = (BDF_GATHER_LDGSTS && NW == 4 && KS == 16) || (BDF_GATHER_LDGSTS1 && NW == 1);
  static constexpr int TPR = NW == 1 ? 16 : 8;
  static constexpr int RPP = NTHR / TPR;
  static_assert(!LDGSTS || KS % RPP == 0, "cp.async gather: a stage is a whole number of passes");
  static constexpr int STG = KS * S * (TENSOR ? 2 : 1) + KS;   // doubles per stage: tile(s) + residuals
  static constexpr int PSZ = 64 * C::NT;  // lower-triangle tiles, 64 doubles each, tile (I,J) at 64·(tri(I)+J)
  static constexpr int REGSZ = PSZ > NBUF * STG ? PSZ : NBUF * STG;  // the tiles alias the (dead) stage ring
  static constexpr int META_D = 4 * KS;  // stage metadata, double-buffered: partner slots (2 × KS × 2 ints) and values (2 × KS doubles)
  static constexpr int SMEM_DOUBLES = REGSZ + 4 * DP + 8 + 4 + META_D;  // … rhs, Λμ, y, x, ts[8], ring mbarriers, stage metadata
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;
  // register-staged loader of the statistics kernel (stats_kernel.cuh)
  static constexpr int SPASSES = NW == 8 ? 2 : (NW == 4 ? 4 : 8);
  static constexpr int SKS = OPP * SPASSES;
  static constexpr int SBUFSZ = 2 * SKS * S + 2 * SKS;
  static_assert(KS % NW == 0 && KS % 4 == 0 && KS / NW <= 32, "a stage is issued KS/NW rows per warp and consumed in k-steps of 4");

  struct Pre {
    double2 a[SPASSES][JP];
    double r[SPASSES];
  };

  static __device__ __forceinline__ void store_stage(const int D, double* buf, double* rs, int tr, int tq, const Pre& pre, bool aug) {
#pragma unroll
    for (int ps = 0; ps < SPASSES; ps++) {
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
  template <int W, bool PROD>
  static __device__ __forceinline__ void compute(double (&acc)[TPW][2], const double* buf, int nk4, int lane) {
    constexpr int NF = C::nfrag(W);
    constexpr int NTW = C::ntiles(W);
    if constexpr (NTW > 0) {
      const double* base = buf + (lane & 3) * S + (lane >> 2);
      constexpr int kUnrollK4 = NW == 4 ? BDF_K4_UNROLL4 : BDF_K4_UNROLL;
#pragma unroll kUnrollK4
      for (int k4 = 0; k4 < nk4; k4++) {
        double f[NF];
        static_for<NF>([&](auto r) {
          constexpr int R = decltype(r)::value;
          f[R] = base[k4 * 4 * S + 8 * FI<C, W, R>::blk];
          if (PROD) f[R] *= base[KS * S + k4 * 4 * S + 8 * FI<C, W, R>::blk];  // second partner (3-mode tensor)
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

  static __device__ __forceinline__ RowCtx make_ctx(const RowParams& p, int item) {
    RowCtx c;
    c.item = item;
    c.lrow = p.item_row[item];
    c.obeg = p.item_beg[item];
    c.len = p.item_len[item];
    c.split = p.item_split[item];
    c.slot = p.slot_base + c.lrow;
    c.rt = &p.rt[p.item_rel ? p.item_rel[item] : 0];
    c.alpha_f = c.split < 0 ? c.rt->alpha : 1.0;
    return c;
  }

#ifdef BDF_DEBUG
#define BDF_STAMP(k)                                                   \
  if (p.dbg && tid == 0) p.dbg[(size_t)c.item * 8 + (k)] = clock64()
#else
#define BDF_STAMP(k)
#endif

  // the stage ring's padding columns (≥ 2·ceil(D/2)) are zeroed once: the bulk copies never touch them
  static __device__ __forceinline__ void ring_init(const RowParams& p, double* ring, uint64_t* fullb, int tid) {
    const int D = p.D;
    const int c0 = 2 * ((D + 1) >> 1);
    if (c0 < DP)
      for (int e = tid; e < NBUF * (TENSOR ? 2 : 1) * KS * (DP - c0); e += NTHR) {
        const int row = e / (DP - c0), col = c0 + e % (DP - c0);
        const int b = row / ((TENSOR ? 2 : 1) * KS), rr = row % ((TENSOR ? 2 : 1) * KS);
        ring[b * STG + rr * S + col] = (TENSOR && rr >= KS && col == D) ? 1.0 : 0.0;  // second partner's aug column = 1
      }
    if (tid == 0) {
#pragma unroll
      for (int b = 0; b < NBUF; b++) mbar_init(fullb + b, LDGSTS ? NTHR : 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }

  // ---- (1) gather + syrk of one work item ------------------------------------------------------------------------------------
  // `ring` / `fullb`: the group's stage ring and its mbarriers (ring_init + a group sync done by the caller); `gs`: stages consumed on this
  // ring so far (the barriers' phases run on across the items of a persistent group); lmu[DP] ← Λ·μ of the row, zs[DP] ← its standard
  // normals. Ends with a group sync: the ring is dead (and free for the next item) when it returns.
  template <class Sync>
  static __device__ __forceinline__ void syrk_item(const RowParams& p, const RowCtx& c, double* ring, uint64_t* fullb, double* metab, uint32_t& gs, double* lmu,
                                                   double* zs, double (&acc)[TPW][2], double& bsum, int tid, Sync sync, long long* tprof = nullptr) {
    const int lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const bool aug = use_aug(D);
    const RelTab& rt = *c.rt;
#ifdef BDF_DEBUG
    long long tp0 = tprof ? clock64() : 0;
#define BDF_TP(k) if (tprof) { const long long t1 = clock64(); tprof[k] += t1 - tp0; tp0 = t1; }
#else
#define BDF_TP(k)
#endif
    const int len = c.len;
    const int64_t obeg = c.obeg, oend = obeg + len;
#pragma unroll
    for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
    bsum = 0.0;  // Σ v_tid · r  (used when !aug)

    const int nst = (len + KS - 1) / KS;
    const int npc = (D + 1) >> 1;  // 16-byte pieces that carry latent columns
    // Gather by TMA: one 1-D bulk copy (cp.async.bulk → UBLKCP) moves one whole partner row (npc·16 bytes) and signals the stage's
    // mbarrier by byte count. The partner slot and the value of an observation are fetched one stage ahead of their use with cp.async
    // (LDGSTS) into a small double-buffered shared-memory table — not into registers: next to the 92 accumulator registers the
    // compiler spilled them, and the spill store made the warp wait out the HBM latency of the CSR stream on every stage.
    const uint32_t rowbytes = (uint32_t)npc * 16u;
    // each warp issues its share of a stage (KS/NW rows): UBLKCP is a per-thread (uniform-datapath) instruction, so 16 copies
    // from one warp serialise for ~1000 cycles per stage; spread over the warps they go out in parallel
    constexpr int KQ = KS / NW;
    const bool gl = lane < KQ;
    const int gk = warp * KQ + lane;  // this lane's observation within a stage (valid when gl)
    int* mcol = reinterpret_cast<int*>(metab);  // [2][2][KS]
    double* mval = metab + 2 * KS;              // [2][KS]
    auto load_meta = [&](int s) {
      if (gl) {
        int64_t o = obeg + (int64_t)s * KS + gk;
        if (o >= oend) o = oend - 1;
        const int mb = (int)((gs + (uint32_t)s) & 1u);
        cp_async4(mcol + (mb * 2) * KS + gk, rt.col0 + o);
        if (TENSOR && rt.col1) cp_async4(mcol + (mb * 2 + 1) * KS + gk, rt.col1 + o);
        cp_async8(mval + mb * KS + gk, rt.val + o);
      }
      cp_async_commit();
    };
    auto issue = [&](int s) {
      if constexpr (LDGSTS) {
        // TPR threads per partner row, 16-byte pieces (t % TPR) + TPR·j; rows past the end of the item are zero-filled (src-size 0)
        const int b = (int)((gs + (uint32_t)s) % NBUF);
        double* st = ring + b * STG;
        int nvalid = len - s * KS;
        if (nvalid > KS) nvalid = KS;
        const int q = tid % TPR;
        const int mb = (int)((gs + (uint32_t)s) & 1u);
#pragma unroll 1
        for (int ps = 0; ps < KS / RPP; ps++) {
          const int row = ps * RPP + tid / TPR;
          const bool ok = row < nvalid;
          const int c0 = mcol[(mb * 2) * KS + row];
          const double* src0 = rt.P0 + (size_t)c0 * p.ld;
#pragma unroll 1
          for (int pc = q; pc < npc; pc += TPR) cp_async16(st + row * S + 2 * pc, src0 + 2 * pc, ok ? 16 : 0);
          if (TENSOR) {
            const int c1 = rt.col1 ? mcol[(mb * 2 + 1) * KS + row] : 0;
            const double* src1 = rt.P1 + (size_t)c1 * p.ld;
#pragma unroll 1
            for (int pc = q; pc < npc; pc += TPR) cp_async16(st + (KS + row) * S + 2 * pc, src1 + 2 * pc, ok ? 16 : 0);
          }
          if (q == 0) {
            const double r = ok ? mval[mb * KS + row] - rt.mean : 0.0;
            st[(TENSOR ? 2 : 1) * KS * S + row] = r;
            if (aug) st[row * S + D] = r;  // the aug column (index D, D even) lies outside the copied pieces
          }
        }
        cp_async_mbar_arrive_noinc(fullb + b);
        cp_async_commit();  // the gather of a stage is one cp.async group of its own (see the wait in the stage loop)
        return;
      }
      if (gl) {
        const int b = (int)((gs + (uint32_t)s) % NBUF);
        double* st = ring + b * STG;
        int nvalid = len - s * KS;
        if (nvalid > KS) nvalid = KS;
        const bool ok = gk < nvalid;
        if (tid == 0) mbar_arrive_expect_tx(fullb + b, (uint32_t)nvalid * rowbytes * (TENSOR ? 2u : 1u));
        const int mb = (int)((gs + (uint32_t)s) & 1u);
        const int c0n = mcol[(mb * 2) * KS + gk];
        const int c1n = (TENSOR && rt.col1) ? mcol[(mb * 2 + 1) * KS + gk] : 0;
        const double rvn = mval[mb * KS + gk];
        if (ok) {
          bulk_copy_g2s(st + gk * S, rt.P0 + (size_t)c0n * p.ld, rowbytes, fullb + b);
          if (TENSOR) bulk_copy_g2s(st + (KS + gk) * S, rt.P1 + (size_t)c1n * p.ld, rowbytes, fullb + b);
        } else if (gk < ((nvalid + 3) & ~3)) {
          // rows of the last, partly filled k-step: zero them (only the final stage of an item ever takes this path)
          for (int cc = 0; cc < 2 * npc; cc += 2) {
            *reinterpret_cast<double2*>(st + gk * S + cc) = make_double2(0.0, 0.0);
            if (TENSOR) *reinterpret_cast<double2*>(st + (KS + gk) * S + cc) = make_double2(0.0, 0.0);
          }
        }
        const double r = ok ? rvn - rt.mean : 0.0;
        st[(TENSOR ? 2 : 1) * KS * S + gk] = r;
        if (aug) st[gk * S + D] = r;
      }
    };
#ifdef BDF_DEBUG
    const bool dbg_nogather = p.flags & 4, dbg_nocompute = p.flags & 2;  // timing experiments only (results invalid)
#else
    constexpr bool dbg_nogather = false, dbg_nocompute = false;
#endif
    static_assert(PF <= 2, "the metadata table is double-buffered: at most one stage's metadata may be in flight beside the one in use");
    if (nst > 0) {  // the only exposed latency of an item: the metadata of its first stage(s)
      load_meta(0);
      if (PF > 1 && nst > 1) load_meta(1);
      cp_async_wait<0>();
      if (LDGSTS) sync();  // every thread reads the metadata the issuing lanes fetched
      if (!dbg_nogather) {
        issue(0);
        if (PF > 1 && nst > 1) issue(1);
      }
      if (PF < nst) {
        if (LDGSTS) sync();  // … and is done reading table buffer 0 before it is refilled
        load_meta(PF);
      }
    }
    // ---- everything below runs under the shadow of the first gathers ---------------------------------------------------
    // Λ rides in the accumulators from the start (as Λ/α, on the row's first chunk), so that parking the tiles after the
    // syrk yields Λ* = α·(Λ/α + G) in one pass; the loads are L2 hits hidden under the first gather.
    if (c.split < 0 || p.item_chunk[c.item] == 0) {
      const double ia = 1.0 / rt.alpha;
      warp_dispatch(warp, [&](auto w) {
        constexpr int W = decltype(w)::value;
        static_for<C::ntiles(W)>([&](auto t) {
          constexpr int T = decltype(t)::value;
          using ti = TI<C, W, T>;
          const int i = 8 * ti::I + (lane >> 2), j = 8 * ti::J + 2 * (lane & 3);
          const double2 l = __ldg(reinterpret_cast<const double2*>(p.LT + 64 * (tri(ti::I) + ti::J) + 2 * lane));
          if (i < D) {
            if (j < D) acc[T][0] = l.x * ia;
            if (j + 1 < D) acc[T][1] = l.y * ia;
          }
        });
      });
    }
    // Λ·μ: precomputed when μ is shared; per-row μ (side features, src/macau.jl:102-107) is multiplied here
    if (tid < DP) {
      double s = 0.0;
      if (p.lmu) {
        s = __ldg(p.lmu + tid);
      } else if (tid < D) {
        const double* mu = p.mu + c.slot * p.mu_ld;
        for (int i = 0; i < D; i++) s = fma(__ldg(p.Lambda + tid + (size_t)i * D), __ldg(mu + i), s);
      }
      lmu[tid] = s;
    }
    // The row's standard normals (injected, or Philox + Box–Muller in double: a few hundred instructions per PAIR) are
    // produced here in parallel, under the shadow of the first gather, and parked until the substitution needs them — not serially by
    // one warp at the end of the row. Even threads compute a Box–Muller pair and hand its second member to their odd neighbour.
    {
      double z = 0.0, zn = 0.0;
      const int64_t grow0 = p.row_of_slot ? (int64_t)p.row_of_slot[c.slot] : (int64_t)c.lrow * p.world + p.rank;
      if (p.Z) {
        if (tid < D) z = __ldg(p.Z + (size_t)c.slot * p.ld + tid);
      } else if (tid < D && !(tid & 1)) {
        philox_normal_pair(p.seed, p.sweep, philox_stream(PHILOX_ROW, (uint32_t)p.entity), grow0, tid >> 1, z, zn);
      }
      if (!p.Z) {  // uniform branch: every lane of the warp takes part in the shuffle
        const double up = __shfl_up_sync(0xffffffffu, zn, 1);
        if ((tid & 1) && tid < D) z = up;
      }
      if (tid < DP) zs[tid] = z;
    }

    BDF_STAMP(1);
    BDF_TP(0)
    for (int s = 0; s < nst; s++) {
      const uint32_t g = gs + (uint32_t)s;
      if (!dbg_nogather) mbar_wait(fullb + (g % NBUF), (g / NBUF) & 1);  // the rows of stage s have landed
      BDF_TP(1)
      if (s + PF < nst) {
        // the metadata of stage s+PF, requested a stage ago. With the cp.async gather the groups alternate (metadata, gather, metadata, …):
        // all but the newest group — the gather issued in the previous trip — must have landed
        if (LDGSTS && s > 0) cp_async_wait<1>(); else cp_async_wait<0>();
      }
      sync();                                          // everyone is done with stage s-1 (its buffer may be refilled) and sees the metadata
      BDF_TP(2)
      if (s + PF + 1 < nst) load_meta(s + PF + 1);
      if (s + PF < nst && !dbg_nogather) issue(s + PF);
      BDF_TP(3)
      const double* buf = ring + (g % NBUF) * STG;
      const double* rs = buf + (TENSOR ? 2 : 1) * KS * S;
      int rem = len - s * KS;
      if (rem > KS) rem = KS;
      const int nk4 = (rem + 3) >> 2;
      if (!dbg_nocompute) warp_dispatch(warp, [&](auto w) { compute<decltype(w)::value, TENSOR>(acc, buf, nk4, lane); });
      if (!aug && tid < DP) {
        for (int k = 0; k < nk4 * 4; k++) {
          double v = buf[k * S + tid];
          if (TENSOR) v *= buf[(KS + k) * S + tid];
          bsum = fma(v, rs[k], bsum);
        }
      }
      BDF_TP(4)
    }
    gs += (uint32_t)nst;
    sync();  // the ring is dead from here on
    BDF_TP(2)
    BDF_STAMP(2);
#undef BDF_TP
  }

  // ---- (2) split rows: park the partial; the last item of a GROUP of consecutive chunks adds the group's partials in chunk order, and
  //      (rows with many chunks) parks the group sum, the last group then adds the group sums in group order. Two levels keep the
  //      serial part short — a row with 2M observations has hundreds of 46 KB partials — and the order fixed (deterministic).
  //      Returns true for the item that ends up with the complete row in its accumulators. ------------------------------------------
  template <class Sync>
  static __device__ __forceinline__ bool split_reduce(const RowParams& p, const RowCtx& c, double (&acc)[TPW][2], double& bsum, int tid,
                                                      volatile int* s_last, Sync sync) {
    const int lane = tid & 31, warp = tid >> 5;
    const int split = c.split;
    const int nch = p.split_nchunks[split];
    const int G = p.split_gsize[split];  // chunks per group
    const int ng = (nch + G - 1) / G;
    const int ch = p.item_chunk[c.item], g = ch / G;
    const int gsz = (g == ng - 1) ? nch - g * G : G;
    const int64_t gc = p.split_gcoff[split] + g;  // this group's counter / partial slot
    auto park = [&](double* part, double sc) {
#pragma unroll
      for (int t = 0; t < TPW; t++)
        *reinterpret_cast<double2*>(part + ((size_t)(warp * TPW + t) * 32 + lane) * 2) = make_double2(sc * acc[t][0], sc * acc[t][1]);
      if (tid < DP) part[NW * TPW * 64 + tid] = sc * bsum;
      __threadfence();
      sync();
    };
    auto arrive = [&](int* counter, int expect) {  // true for the last arriver (which also resets the counter)
      if (tid == 0) {
        const int old = atomicAdd(counter, 1);
        const int last = (old == expect - 1);
        *s_last = last;
        if (last) *counter = 0;
      }
      sync();
      const bool last = *s_last;
      sync();
      if (last) __threadfence();
      return last;
    };
    auto add_up = [&](const double* base, int n) {
#pragma unroll
      for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
      bsum = 0.0;
      for (int k = 0; k < n; k++) {
        const double* pc = base + (size_t)k * PST;
#pragma unroll
        for (int t = 0; t < TPW; t++) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(pc + ((size_t)(warp * TPW + t) * 32 + lane) * 2));
          acc[t][0] += v.x;
          acc[t][1] += v.y;
        }
        if (tid < DP) bsum += __ldcg(pc + NW * TPW * 64 + tid);
      }
    };
    park(p.ws + (size_t)(p.split_wsoff[split] + ch) * PST, c.rt->alpha);
    if (!arrive(p.group_counter + gc, gsz)) return false;
    add_up(p.ws + (size_t)(p.split_wsoff[split] + (int64_t)g * G) * PST, gsz);
    if (ng > 1) {
      park(p.ws + (size_t)(p.gslot_base + gc) * PST, 1.0);
      if (!arrive(p.split_counter + split, ng)) return false;
      add_up(p.ws + (size_t)(p.gslot_base + p.split_gcoff[split]) * PST, ng);
    }
    return true;
  }

  // ---- (3) park Λ* = α·acc in shared memory (tile t = tri(I)+J, 64 doubles in the tile layout above), identity on the padding;
  //      the augmented row (i == D) carries Σ v·r and becomes rhs = Λμ + α·Σv·r. No synchronisation inside. ---------------------------
  static __device__ __forceinline__ void park_tiles(const RowParams& p, const RowCtx& c, double (&acc)[TPW][2], double bsum, const double* lmu,
                                                    double* Tl, double* rhs, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const bool aug = use_aug(D);
    const double alpha = c.alpha_f;
    warp_dispatch(warp, [&](auto w) {
      constexpr int W = decltype(w)::value;
      static_for<C::ntiles(W)>([&](auto t) {
        constexpr int T = decltype(t)::value;
        using ti = TI<C, W, T>;
        double2 v = make_double2(alpha * acc[T][0], alpha * acc[T][1]);
        if constexpr (ti::I == NB - 1) {  // only the last block row / column can touch the padding
          const int i = 8 * ti::I + (lane >> 2), j = 8 * ti::J + 2 * (lane & 3);
          if (aug && i == D) {
            if (j < D) rhs[j] = fma(alpha, acc[T][0], lmu[j]);
            if (j + 1 < D) rhs[j + 1] = fma(alpha, acc[T][1], lmu[j + 1]);
          }
          if (i >= D || j >= D) v.x = (i == j) ? 1.0 : 0.0;
          if (i >= D || j + 1 >= D) v.y = (i == j + 1) ? 1.0 : 0.0;
        }
        *reinterpret_cast<double2*>(Tl + 64 * (tri(ti::I) + ti::J) + tile_acc_off(lane)) = v;
      });
    });
    if (!aug && tid < D) rhs[tid] = fma(alpha, bsum, lmu[tid]);
    if (tid >= D && tid < DP) rhs[tid] = 0.0;
  }

  // ---- (4) blocked UL factorisation Λ* = W·Wᵀ (W upper), block rows pb = NB-1 … 0, on the shared-memory tiles; y = W⁻¹·rhs rides along;
  //      then x = W⁻ᵀ(y + z) by forward substitution and the store of the draw (into every peer replica too). xs holds z on entry. -------
  // Per panel: (b) the panel tiles are scaled, R_pJ = W_pp⁻¹·A_pJ, by one DMMA pair each; (c) the tiles above the
  // panel get A_IJ −= R_pIᵀ·R_pJ, again DMMA, one block row per warp at a time; one warp takes the next diagonal tile
  // first and factors it while the other warps finish the trailing update (look-ahead). W_pp⁻¹ replaces the diagonal tile (p, p),
  // which nothing reads once it is factored.
  // The serial roles (diagonal blocks, substitutions) rotate over the warps from row to row (`rot`): warp w of every co-resident
  // group sits on the same SM sub-partition, so a fixed "warp 0" would pile every row's dependent chain onto one scheduler.
  template <class Sync>
  static __device__ __forceinline__ void factor_and_draw(const RowParams& p, const RowCtx& c, double* Tl, double* rhs, double* ys, double* xs,
                                                         double* ts, int rot, int tid, Sync sync) {
    const int lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
#if BDF_VWARP
    const int vw = (warp - rot) & (NW - 1);
#else
    const int vw = warp;
#endif
    bool bad = false;
    const int fo = tile_frag_off(lane);  // operand-fragment offset in a tile: A[m][k] = B[k][m] = tile(k, m)
    const int ao = tile_acc_off(lane);   // accumulator-layout offset (double2)
    auto Wv = [&](int pb) { return Tl + 64 * (tri(pb) + pb); };  // W_pp⁻¹, in the tile layout, in place of the diagonal tile
    auto factor_diag = [&](int pb) {
      // A_pp = W·Wᵀ by elimination from the last column to the first, in the DMMA accumulator layout (lane = 4·row+q
      // holds columns 2q, 2q+1); an identity block carried along ends up as W⁻¹.
      // The next pivot is formed from pre-update values as soon as the current scale is known, so its rsqrt overlaps
      // the rank-1 update instead of waiting for it.
      const int r = lane >> 2, q = lane & 3;
      const double2 av = *reinterpret_cast<const double2*>(Wv(pb) + ao);
      double a0 = av.x, a1 = av.y;
      double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
      double piv = __shfl_sync(0xffffffffu, a1, 4 * 7 + 3);  // a_77
      constexpr int kUnrollFD = BDF_FD_UNROLL;
#pragma unroll kUnrollFD
      for (int j = 7; j >= 0; j--) {
        const int jq = j >> 1;
        const double sel = (j & 1) ? a1 : a0;
        if (!(piv > 0.0)) bad = true;
        const double sc = fast_rsqrt(piv);
        // pre-update values needed for the next pivot: a_{j-1,j-1} and a_{j-1,j}
        double pn = 0.0, pc = 0.0;
        if (j > 0) {
          pn = __shfl_sync(0xffffffffu, ((j - 1) & 1) ? a1 : a0, 4 * (j - 1) + ((j - 1) >> 1));
          pc = __shfl_sync(0xffffffffu, sel, 4 * (j - 1) + jq);
        }
        const double wij = __shfl_sync(0xffffffffu, sel, (lane & ~3) | jq) * sc;  // w_ij of this lane's row
        const double rj0 = __shfl_sync(0xffffffffu, a0, 4 * j + q) * sc;          // w_kj for this lane's two columns
        const double rj1 = __shfl_sync(0xffffffffu, a1, 4 * j + q) * sc;
        const double ej0 = __shfl_sync(0xffffffffu, e0, 4 * j + q) * sc;          // row j of W⁻¹
        const double ej1 = __shfl_sync(0xffffffffu, e1, 4 * j + q) * sc;
        if (j > 0) {
          const double w = pc * sc;
          piv = fma(-w, w, pn);  // a_{j-1,j-1} − w_{j-1,j}²: identical to what the update below leaves in the tile
        }
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
      *reinterpret_cast<double2*>(Wv(pb) + ao) = make_double2(e0, e1);  // element (row, col) of W_pp⁻¹ at tile_el(row, col)
    };
    // trailing update of block row I above panel pb: tiles (I, J), J = j0 … I
    auto update_row = [&](int pb, int I, int j0, int j1) {
      const double* Pp = Tl + 64 * tri(pb) + fo;
      const double na0 = -Pp[64 * I], na1 = -Pp[64 * I + 32];
      double* trow = Tl + 64 * tri(I) + ao;
      // UB tiles at a time: all their operands are loaded before the first DMMA, so the shared-memory latency and the
      // two dependent DMMAs of a tile overlap across the batch (the compiler cannot hoist loads over the tile stores)
      constexpr int UB = BDF_UR_UNROLL;
      for (int J = j0; J <= j1; J += UB) {
        double b0[UB], b1[UB], c2[UB][2];
#pragma unroll
        for (int u = 0; u < UB; u++) {
          const int Ju = J + u <= j1 ? J + u : j1;
          b0[u] = Pp[64 * Ju];
          b1[u] = Pp[64 * Ju + 32];
          const double2 cv = *reinterpret_cast<const double2*>(trow + 64 * Ju);
          c2[u][0] = cv.x;
          c2[u][1] = cv.y;
        }
#pragma unroll
        for (int u = 0; u < UB; u++) dmma884(c2[u], na0, b0[u]);
#pragma unroll
        for (int u = 0; u < UB; u++) dmma884(c2[u], na1, b1[u]);
#pragma unroll
        for (int u = 0; u < UB; u++)
          if (J + u <= j1) *reinterpret_cast<double2*>(trow + 64 * (J + u)) = make_double2(c2[u][0], c2[u][1]);
      }
    };
    constexpr int BS_SWITCH = NW == 4 ? BDF_BS_SWITCH : 1000;  // panels at or above this: backsub_step rides on the chain warp (off by default)
    // one block of the substitution y = W⁻¹·rhs, runnable as soon as block row J is final: y_J = W_JJ⁻¹·rhs_J, then
    // rhs[c] −= Σ_k R_J[k][c]·y_J[k] for c < 8J. One warp; rides along with the trailing update.
    auto backsub_step = [&](int J, bool update) {
      const int r8 = lane & 7;
      const double* wv = Wv(J);
      double y0 = 0.0, y1 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        y0 = fma(wv[tile_el(r8, k)], rhs[8 * J + k], y0);
        y1 = fma(wv[tile_el(r8, k + 1)], rhs[8 * J + k + 1], y1);
      }
      if (lane < 8) ys[8 * J + r8] = y0 + y1;
      __syncwarp();
      if (update) {
        for (int cc = lane; cc < 8 * J; cc += 32) {
          const double* tp = Tl + 64 * (tri(J) + (cc >> 3));  // R_J[k][c] = tile(J, c/8)(k, c%8)
          double s0 = rhs[cc], s1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            s0 = fma(-tp[tile_el(k, cc & 7)], ys[8 * J + k], s0);
            s1 = fma(-tp[tile_el(k + 1, cc & 7)], ys[8 * J + k + 1], s1);
          }
          rhs[cc] = s0 + s1;
        }
        __syncwarp();
      }
    };

#ifdef BDF_DEBUG
    long long d_b = 0, d_c = 0, d_w1 = 0, d_w2 = 0, t_x = 0;
#define BDF_LAP(acc_) if (p.dbg) { const long long t = clock64(); acc_ += t - t_x; t_x = t; }
#define BDF_LAP0() if (p.dbg) t_x = clock64()
#else
#define BDF_LAP(acc_)
#define BDF_LAP0()
#endif
    if (vw == 0) factor_diag(NB - 1);
    sync();
    for (int pb = NB - 1; pb > 0; pb--) {
      BDF_LAP0();
      // (b) scale the panel tiles (pb, J < pb) in place
      {
        // A operand = W_pp⁻¹[m][k], m = lane/4, k = lane%4 (+4): stored untransposed, so its fragment is read along a tile row
        const int wo = tile_el(lane >> 2, lane & 3);
        const double wa0 = Wv(pb)[wo], wa1 = Wv(pb)[wo ^ 4];
        for (int J = vw; J < pb; J += NW) {
          double* tp = Tl + 64 * (tri(pb) + J);
          const double b0 = tp[fo], b1 = tp[32 + fo];
          double c2[2] = {0.0, 0.0};
          dmma884(c2, wa0, b0);
          dmma884(c2, wa1, b1);
          *reinterpret_cast<double2*>(tp + ao) = make_double2(c2[0], c2[1]);
        }
      }
      BDF_LAP(d_b)
      sync();
      BDF_LAP(d_w1)
      // (c) trailing update of block rows I < pb; rows are dealt to warps 1…NW-1 in a snake so the triangle balances
      if (NW == 1) {
        backsub_step(pb, true);
        for (int I = 0; I < pb; I++) update_row(pb, I, 0, I);
        __syncwarp();
        factor_diag(pb - 1);
      } else if (vw == 0) {
        update_row(pb, pb - 1, pb - 1, pb - 1);
        __syncwarp();
        factor_diag(pb - 1);
        if (pb >= BS_SWITCH) backsub_step(pb, true);  // large panels: the trailing warps are the critical path, the chain warp has slack
      } else {
        if (vw == 1 && pb < BS_SWITCH) backsub_step(pb, true);
        constexpr int NWC = NW > 1 ? NW - 1 : 1;
        for (int n = 0; n < pb; n++) {
          const int I = pb - 1 - n;
          const int ph = n % (2 * NWC);
          const int wo = 1 + (ph < NWC ? ph : 2 * NWC - 1 - ph);
          if (wo == vw) update_row(pb, I, 0, n == 0 ? I - 1 : I);  // (pb-1, pb-1) belongs to the chain warp
        }
      }
      BDF_LAP(d_c)
      sync();
      BDF_LAP(d_w2)
    }
#ifdef BDF_DEBUG
    if (p.dbg && lane == 0 && vw < 2) {
      long long* o = p.dbg + (size_t)gridDim.x * 8 + ((size_t)c.item * 2 + vw) * 4;
      o[0] = d_b; o[1] = d_w1; o[2] = d_c; o[3] = d_w2;
    }
#endif
    if (bad && lane == 0) atomicOr(p.err_flag, 1);
    BDF_STAMP(5);

    // ---- one warp: last block of y = W⁻¹·rhs, then x = W⁻ᵀ(y + z) by forward substitution over the block rows (a group-wide
    //      version with one barrier per block was measured slower under load: the idle warps' issue slots go to co-resident rows)
    if (vw == 0) {
      {
      const int r8 = lane & 7;
      backsub_step(0, false);
      for (int cc = lane; cc < DP; cc += 32) ys[cc] += xs[cc];  // v = y + z (z parked in xs)
      __syncwarp();
      // forward substitution R·x = v (R = Wᵀ, lower, block rows in the tiles): column-oriented — lane owns rows
      // lane, lane+32, … of v in registers; once x_J is known every lane subtracts R[i][8J..8J+7]·x_J from its rows.
      {
        constexpr int NS = (DP + 31) / 32;
        double vreg[NS];
#pragma unroll
        for (int s2 = 0; s2 < NS; s2++) vreg[s2] = (lane + 32 * s2 < DP) ? ys[lane + 32 * s2] : 0.0;
        for (int J = 0; J < NB; J++) {
          // x_J = W_JJ⁻ᵀ·v_J : v_J sits in the registers of lanes 8J%32 … of slot 8J/32 → share it through ts
          {
            const int sl = (8 * J) >> 5;
            double mine = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < NS; s2++)
              if (s2 == sl) mine = vreg[s2];
            if ((lane >> 3) == ((8 * J) & 31) >> 3) ts[lane & 7] = mine;
          }
          __syncwarp();
          const double* wv = Wv(J);
          double x0 = 0.0, x1 = 0.0;
#pragma unroll
          for (int kk = 0; kk < 8; kk += 2) {
            x0 = fma(wv[tile_el(kk, r8)], ts[kk], x0);
            x1 = fma(wv[tile_el(kk + 1, r8)], ts[kk + 1], x1);
          }
          const double xj = x0 + x1;  // lane r8 (replicated over the 4 lane groups) holds x[8J + r8]
          if (lane < 8) xs[8 * J + r8] = xj;
          // this lane's tile row is i%8 = lane%8 for every s2; rows 2, 3, 6, 7 store their column halves swapped (tile_el), so those
          // lanes take x_J with its halves swapped as well and read the row as it lies in memory
          const int hs = (lane & 2) << 1;
          double xb[8];
#pragma unroll
          for (int k = 0; k < 8; k++) xb[k] = __shfl_sync(0xffffffffu, xj, k ^ hs);
#pragma unroll
          for (int s2 = 0; s2 < NS; s2++) {
            const int i = lane + 32 * s2;
            if (i >= 8 * (J + 1) && i < DP) {
              const double* tp = Tl + 64 * (tri(i >> 3) + J) + 8 * (i & 7);  // R[i][8J .. 8J+7], physical order
              const double2 t0 = *reinterpret_cast<const double2*>(tp), t1 = *reinterpret_cast<const double2*>(tp + 2);
              const double2 t2 = *reinterpret_cast<const double2*>(tp + 4), t3 = *reinterpret_cast<const double2*>(tp + 6);
              double a = fma(t0.x, xb[0], t0.y * xb[1]), b = fma(t1.x, xb[2], t1.y * xb[3]);
              a = fma(t2.x, xb[4], fma(t2.y, xb[5], a));
              b = fma(t3.x, xb[6], fma(t3.y, xb[7], b));
              vreg[s2] -= a + b;
            }
          }
          __syncwarp();
        }
      }
      }
      __syncwarp();
      double* out = p.Uout + (size_t)c.slot * p.ld;
      for (int j = lane; j < p.ld; j += 32) out[j] = j < D ? xs[j] : 0.0;
      // fused all-gather: the drawn row goes straight into every peer's replica over NVLink (no collective afterwards)
#pragma unroll 1
      for (int r = 0; r < 8 && p.peer_out[r]; r++) {
        double* po = p.peer_out[r] + (size_t)c.slot * p.ld;
        for (int j = lane; j < p.ld; j += 32) po[j] = j < D ? xs[j] : 0.0;
      }
      __syncwarp();
#ifdef BDF_DEBUG
      if (p.dbg && lane == 0) p.dbg[(size_t)c.item * 8 + 6] = clock64();
#endif
    }
#undef BDF_LAP
#undef BDF_LAP0
  }

  // ---- the kernel body: one CTA per work item --------------------------------------------------------------------------------
  static __device__ void run(const RowParams& p, double* smem) {
    const int tid = threadIdx.x;
    const RowCtx c = make_ctx(p, blockIdx.x);
    double* ring = smem;                    // [NBUF][STG]: tile (KS×S), [second tile], residuals (KS)
    double* Tl = smem;                      // Λ* / factor tiles, alias the ring after the main loop
    double* rhs = smem + REGSZ;             // [DP]
    double* lmu = rhs + DP;                 // Λ·μ [DP]
    double* ys = lmu + DP;                  // W⁻¹·rhs (+ z) [DP]
    double* xs = ys + DP;                   // z, then the draw [DP]
    double* ts = xs + DP;                   // [8]
    uint64_t* fullb = reinterpret_cast<uint64_t*>(ts + 8);
    double* metab = ts + 8 + 4;
    __shared__ int s_last;
    BDF_STAMP(0);
    double acc[TPW][2];
    double bsum;
    uint32_t gs = 0;
    ring_init(p, ring, fullb, tid);
    __syncthreads();  // ring zero-fill and barrier init visible before any stage is issued or consumed
    syrk_item(p, c, ring, fullb, metab, gs, lmu, xs, acc, bsum, tid, CtaSync{});
    if (c.split >= 0 && !split_reduce(p, c, acc, bsum, tid, &s_last, CtaSync{})) return;
    BDF_STAMP(3);
#ifdef BDF_DEBUG
    if (p.flags & 1) {
      if (tid == 0 && acc[0][0] == 1.2345) p.Uout[0] = acc[0][1];
      if (p.dbg && tid == 0) p.dbg[(size_t)c.item * 8 + 4] = p.dbg[(size_t)c.item * 8 + 5] = p.dbg[(size_t)c.item * 8 + 6] = clock64();
      return;
    }
#endif
    park_tiles(p, c, acc, bsum, lmu, Tl, rhs, tid);
    __syncthreads();
    BDF_STAMP(4);
    factor_and_draw(p, c, Tl, rhs, ys, xs, ts, c.item, tid, CtaSync{});
  }
#undef BDF_STAMP
};

#ifndef BDF_MINB
#define BDF_MINB 3
#endif
#ifndef BDF_MINB4BIG
#define BDF_MINB4BIG 4
#endif
// 4-warp CTAs for D > 64: as many rows in flight per SM as shared memory allows (≤ BDF_MINB4BIG), registers sized to match
template <class K>
constexpr int big4_min_blocks() {
  const int fit = (int)(233472 / (K::SMEM_BYTES + 1024));
  return fit < 1 ? 1 : (fit > BDF_MINB4BIG ? BDF_MINB4BIG : fit);
}
template <class K>
__global__ void __launch_bounds__(K::NTHR, (K::NW == 1 ? BDF_MINB1 : (K::NW == 4 ? (K::DP > 64 ? big4_min_blocks<K>() : (K::TENSOR ? 4 : BDF_MINB4S)) : (K::TENSOR ? 2 : BDF_MINB)))) row_kernel(const RowParams p) {
  extern __shared__ __align__(16) double smem_dyn[];
  K::run(p, smem_dyn);
}

}  // namespace bdf
