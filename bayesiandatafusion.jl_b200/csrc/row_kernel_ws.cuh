// Warp-specialised, persistent form of the per-row draw for 64 < D ≤ 104 (the algorithm and its pieces are those of row_kernel.cuh).
//
// Why: at D ≈ 100 a short row spends half of its life in the factorisation + substitutions, a dependent chain that barely uses the FP64
// pipe, while its syrk is pure DMMA throughput. With one CTA per row the rows in flight per SM are capped at 4 by the REGISTER file
// (the Gram accumulators of a 104×104 lower triangle are 92 registers per thread of a 4-warp group) and the pipe idles 45 % of the time
// (ncu, users launch of C2). Only rows in their syrk phase need those registers; rows being factored need shared memory (the 46.6 KB of
// tiles) but few registers. So one persistent CTA per SM holds
//   * NSG = 2 "syrk" groups of 4 warps at 128 registers: each streams work items off a global queue (heaviest first) — TMA gather ring →
//     DMMA accumulate → (split rows: park/add up partials in global memory) → park Λ* into a free tile slot → signal → next item;
//   * NFG = 3 "finalise" groups of 4 warps at 72 registers (setmaxnreg), each owning one 48 KB tile slot: wait for a parked row →
//     blocked UL Cholesky → substitutions → store the draw (and into the peers) → free the slot.
// i.e. 5 rows in flight per SM, two of them feeding the pipe. Hand-off is by shared-memory mbarriers (full[slot] / empty[slot]); each
// group synchronises internally on its own named barrier. Slots are handed out round-robin by ticket; the ticket is taken and its slot
// awaited under a lock, so tickets reach the slot barriers in order and a phase parity can never be mistaken for the one two uses later.
#pragma once
#include "row_kernel.cuh"

#ifndef BDF_WS_KS
#define BDF_WS_KS 16
#endif
#ifndef BDF_WS_NBUF
#define BDF_WS_NBUF 2
#endif

namespace bdf {

template <int DP_, bool TENSOR_>
struct RowKernelWS {
  using K = RowKernel<DP_, 4, TENSOR_, BDF_WS_KS, BDF_WS_NBUF>;
  static constexpr int DP = DP_;
  static constexpr int NSG = 2, NFG = 3, GT = 128;  // syrk groups, finalise groups (= tile slots), threads per group
  static constexpr int NTHR = (NSG + NFG) * GT;
  static constexpr int TPW = K::TPW;
  static constexpr int SLOT_D = K::PSZ + 3 * DP + 8;                     // tiles, rhs, y, x/z, ts[8]
  static constexpr int GRP_D = (K::NBUF * K::STG + 2 * DP + K::NBUF + 2 + K::META_D + 1) & ~1;  // ring, Λμ, z, ring mbarriers, (s_last, hand), stage metadata; even: 16-byte aligned groups
  static constexpr int CTRL_D = 2 * NFG + 4 + 2 * NFG;                   // full[], empty[], (lock, ticket, done, pad), slot meta
  static constexpr int SMEM_DOUBLES = NFG * SLOT_D + NSG * GRP_D + CTRL_D;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;
  static constexpr int REGS_SYRK = 128, REGS_FIN = 72, REGS_LAUNCH = 96;  // setmaxnreg only moves registers WITHIN the CTA's launch allocation (NTHR × REGS_LAUNCH)
  static_assert(NSG * GT * REGS_SYRK + NFG * GT * REGS_FIN <= NTHR * REGS_LAUNCH && NTHR * REGS_LAUNCH <= 65536, "register budget: the pool is what the CTA was launched with");

  struct SlotMeta {  // 16 bytes per slot, written by the syrk group that parks a row
    int lrow;
    int rot;
    int item;
    int exit;
  };

  static __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SYRK)); }
  static __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_FIN)); }

  static __device__ void run(const RowParams& p, double* smem) {
    const int g = threadIdx.x >> 7, tid = threadIdx.x & (GT - 1);
    double* ctrl = smem + NFG * SLOT_D + NSG * GRP_D;
    uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);
    uint64_t* empty = full + NFG;
    int* ci = reinterpret_cast<int*>(empty + NFG);  // [0] lock, [1] next ticket, [2] syrk groups that ran out of work
    SlotMeta* meta = reinterpret_cast<SlotMeta*>(ctrl + 2 * NFG + 4);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NFG; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      ci[0] = ci[1] = ci[2] = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (g < NSG) {
      double* grp = smem + NFG * SLOT_D + g * GRP_D;
      K::ring_init(p, grp, reinterpret_cast<uint64_t*>(grp + K::NBUF * K::STG + 2 * DP), tid);
    }
    __syncthreads();
    if (g < NSG) {
      setmaxnreg_inc();
      syrk_role(p, smem, g, tid, full, empty, ci, meta);
    } else {
      setmaxnreg_dec();
      finalise_role(p, smem, g - NSG, tid, full, empty, meta);
    }
  }

  // take the next ticket and wait until its slot is free; returns the slot. One thread of a syrk group calls it.
  static __device__ __forceinline__ int acquire_slot(uint64_t* empty, int* ci) {
    while (atomicCAS(&ci[0], 0, 1) != 0) __nanosleep(64);
    const int ticket = ci[1]++;
    const int slot = ticket % NFG, use = ticket / NFG;
    mbar_wait(empty + slot, (use & 1) ^ 1);  // a fresh barrier passes the wait for "the phase before the first"
    __threadfence_block();
    atomicExch(&ci[0], 0);
    return slot;
  }

  static __device__ void syrk_role(const RowParams& p, double* smem, int g, int tid, uint64_t* full, uint64_t* empty, int* ci, SlotMeta* meta) {
    double* grp = smem + NFG * SLOT_D + g * GRP_D;
    double* ring = grp;
    double* lmu = grp + K::NBUF * K::STG;
    double* zs = lmu + DP;
    uint64_t* fullb = reinterpret_cast<uint64_t*>(zs + DP);
    volatile int* gi = reinterpret_cast<volatile int*>(fullb + K::NBUF);  // [0] s_last, [1] item / slot hand-over
    double* metab = reinterpret_cast<double*>(fullb + K::NBUF) + 2;
    const GroupSync sync{1 + g, GT};
    uint32_t gs = 0;
    int nparked = 0;
#ifdef BDF_DEBUG
    long long t_syrk = 0, t_slot = 0, t_park = 0, t_all = clock64(), t0 = 0;
    long long tprof[5] = {0, 0, 0, 0, 0};
#define WS_T0() if (p.dbg) t0 = clock64()
#define WS_LAP(a) if (p.dbg) { const long long t1 = clock64(); a += t1 - t0; t0 = t1; }
#else
#define WS_T0()
#define WS_LAP(a)
#endif
    for (;;) {
      if (tid == 0) gi[1] = atomicAdd(p.work_counter, 1);
      sync();
      const int item = gi[1];
      sync();
      if (item >= p.n_items) break;
      const RowCtx c = K::make_ctx(p, item);
      double acc[TPW][2];
      double bsum;
      WS_T0();
#ifdef BDF_DEBUG
      K::syrk_item(p, c, ring, fullb, metab, gs, lmu, zs, acc, bsum, tid, sync, p.dbg ? tprof : nullptr);
#else
      K::syrk_item(p, c, ring, fullb, metab, gs, lmu, zs, acc, bsum, tid, sync);
#endif
      if (c.split >= 0 && !K::split_reduce(p, c, acc, bsum, tid, gi, sync)) { WS_LAP(t_syrk) continue; }
      WS_LAP(t_syrk)
      if (tid == 0) gi[1] = acquire_slot(empty, ci);
      sync();
      WS_LAP(t_slot)
      const int slot = gi[1];
      double* sl = smem + slot * SLOT_D;
      double* Tl = sl;
      double* rhs = sl + K::PSZ;
      double* xs = rhs + 2 * DP;
      K::park_tiles(p, c, acc, bsum, lmu, Tl, rhs, tid);
      if (tid < DP) xs[tid] = zs[tid];
      if (tid == 0) {
        SlotMeta m;
        m.lrow = c.lrow; m.rot = nparked + g; m.item = c.item; m.exit = 0;
        meta[slot] = m;
      }
      nparked++;
      sync();  // every thread's tile / vector writes are done (and visible to thread 0's release below)
      if (tid == 0) mbar_arrive(full + slot);
      WS_LAP(t_park)
    }
#ifdef BDF_DEBUG
    if (p.dbg && tid == 0) {
      long long* o = p.dbg + ((size_t)blockIdx.x * (NSG + NFG) + g) * 8;
      o[0] = clock64() - t_all; o[1] = t_syrk; o[2] = t_slot; o[3] = t_park; o[4] = nparked;
      long long* o2 = p.dbg + ((size_t)gridDim.x * (NSG + NFG) + (size_t)blockIdx.x * NSG + g) * 8;
      for (int k = 0; k < 5; k++) o2[k] = tprof[k];
    }
#endif
    // out of work: the last syrk group to get here tells every finalise group to leave (the next NFG tickets cover each slot once)
    if (tid == 0) {
      if (atomicAdd(&ci[2], 1) == NSG - 1) {
        for (int s = 0; s < NFG; s++) {
          const int slot = acquire_slot(empty, ci);
          SlotMeta m;
          m.lrow = 0; m.rot = 0; m.item = 0; m.exit = 1;
          meta[slot] = m;
          mbar_arrive(full + slot);
        }
      }
    }
  }

  static __device__ void finalise_role(const RowParams& p, double* smem, int f, int tid, uint64_t* full, uint64_t* empty, SlotMeta* meta) {
    double* sl = smem + f * SLOT_D;
    double* Tl = sl;
    double* rhs = sl + K::PSZ;
    double* ys = rhs + DP;
    double* xs = ys + DP;
    double* ts = xs + DP;
    const GroupSync sync{1 + NSG + f, GT};
#ifdef BDF_DEBUG
    long long t_wait = 0, t_fac = 0, t_all = clock64(), t0 = 0;
    int nrows = 0;
#endif
    for (uint32_t use = 0;; use++) {
      WS_T0();
      mbar_wait(full + f, use & 1);
      WS_LAP(t_wait)
      const SlotMeta m = meta[f];
      if (m.exit) break;
      RowCtx c;
      c.item = m.item; c.lrow = m.lrow; c.len = 0; c.split = -1; c.obeg = 0;
      c.slot = p.slot_base + m.lrow;
      c.rt = &p.rt[0];
      c.alpha_f = 1.0;
      K::factor_and_draw(p, c, Tl, rhs, ys, xs, ts, m.rot, tid, sync);
      sync();  // all warps are done with the slot (the drawing warp has stored the row)
      if (tid == 0) mbar_arrive(empty + f);
#ifdef BDF_DEBUG
      WS_LAP(t_fac)
      nrows++;
#endif
    }
#ifdef BDF_DEBUG
    if (p.dbg && tid == 0) {
      long long* o = p.dbg + ((size_t)blockIdx.x * (NSG + NFG) + NSG + f) * 8;
      o[0] = clock64() - t_all; o[1] = t_wait; o[2] = t_fac; o[3] = 0; o[4] = nrows;
    }
#endif
  }
};

template <class W>
__global__ void __launch_bounds__(W::NTHR, 1) row_kernel_ws(const RowParams p) {
  extern __shared__ __align__(16) double smem_dyn[];
  W::run(p, smem_dyn);
}

}  // namespace bdf
