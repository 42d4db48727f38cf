// Warp-specialised, persistent form of the per-row draw for D > 64 (see row_kernel.cuh for the algorithm).
//
// With D ≈ 100 a row's factorisation + substitutions are a ~50k-cycle dependent chain that uses the FP64 pipe lightly,
// while its syrk is pure DMMA throughput. Running both in one CTA leaves the pipe idle during the chain; here they are
// decoupled: one persistent CTA per SM holds
//   * 8 "syrk" warps that stream work items (dynamic queue, heaviest first): cp.async gather ring → DMMA accumulate →
//     park the Gram tiles in a free tile buffer → signal, and go straight on to the next item;
//   * G "finalise" groups of 4 warps, each owning one tile buffer: wait for a parked row → Λ* = Λ + αG → blocked UL
//     Cholesky (DMMA panels, shuffle-factored diagonal blocks, look-ahead) → substitutions → store the draw → free the
//     buffer.
// Hand-off is by shared-memory mbarriers (full[g] / empty[g]); each role synchronises internally with named barriers.
#pragma once
#include "row_kernel.cuh"

namespace bdf {

template <int DP_, bool TENSOR_>
struct RowKernelWS {
  using K = RowKernel<DP_, 8, TENSOR_>;
  using C = typename K::C;
  static constexpr int DP = DP_;
  static constexpr bool TENSOR = TENSOR_;
  static constexpr int NB = C::NB, NT = C::NT, TPW = K::TPW;
  static constexpr int SW = 8, FW = 4;
  static constexpr int G = DP <= 104 ? 3 : 2;
  static constexpr int NSY = SW * 32, NFI = FW * 32;
  static constexpr int NTHR = NSY + G * NFI;
  static constexpr int KS = K::KS, S = K::S, STG = K::STG, NBUF = K::NBUF, OPP = K::OPP, GP = K::GP, JP = K::JP;
  static constexpr int PST = 64 * NT + DP;                         // parked partial of a split row: tiles + Σv·r side-sum
  static constexpr int GSZ = 64 * NT + NB * 64 + 5 * DP + 8 + 8;   // per group: tiles, WvT, rhs, lmu, ys, xs, bs, ts, meta
  static constexpr int CTRL = 16;                                  // mbarriers + item hand-over words
  static constexpr int SMEM_DOUBLES = NBUF * STG + G * GSZ + CTRL;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;

  struct Meta {  // lives in the group's shared memory (8 doubles)
    int lrow, split, nch, exit;
    long long item;
  };

  // ============================================================ syrk role ============================================
  static __device__ void syrk_role(const RowParams& p, double* smem, int tid) {
    const int lane = tid & 31, warp = tid >> 5, tr = tid >> 4, tq = tid & 15;
    const int D = p.D;
    const bool aug = use_aug(D);
    double* ring = smem;
    double* groups = smem + NBUF * STG;
    uint64_t* full = reinterpret_cast<uint64_t*>(groups + G * GSZ);
    uint64_t* empty = full + G;
    int* hand = reinterpret_cast<int*>(empty + G);  // [0] = next item index
    const int npc = (D + 1) >> 1;
    {
      const int c0 = 2 * npc;
      if (c0 < DP)
        for (int e = tid; e < NBUF * (TENSOR ? 2 : 1) * KS * (DP - c0); e += NSY) {
          const int row = e / (DP - c0), col = c0 + e % (DP - c0);
          const int b = row / ((TENSOR ? 2 : 1) * KS), rr = row % ((TENSOR ? 2 : 1) * KS);
          ring[b * STG + rr * S + col] = (TENSOR && rr >= KS && col == D) ? 1.0 : 0.0;
        }
    }
    if (tid == 0) hand[0] = atomicAdd(p.work_counter, 1);
    named_bar(1, NSY);
    int fseq = 0;  // rows handed to the finalise groups so far
    for (;;) {
      const int item = hand[0];
      named_bar(1, NSY);  // everyone has read hand[0]
      if (item >= p.n_items) break;
      if (tid == 0) hand[0] = atomicAdd(p.work_counter, 1);  // next item, fetched while this one is processed
      const int lrow = __ldg(p.item_row + item);
      const int64_t obeg = __ldg(p.item_beg + item);
      const int len = __ldg(p.item_len + item);
      const int split = __ldg(p.item_split + item);
      const int64_t oend = obeg + len;

      double acc[TPW][2];
#pragma unroll
      for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
      double bsum = 0.0;
      const int nst = (len + KS - 1) / KS;
      int cn0[GP], cn1[GP];
      double rn[GP];
      auto load_meta = [&](int s) {
#pragma unroll
        for (int ps = 0; ps < GP; ps++) {
          int64_t o = obeg + (int64_t)s * KS + tr + ps * OPP;
          if (o >= oend) o = oend - 1;
          if (o < obeg) o = obeg;
          cn0[ps] = __ldg(p.col0 + o);
          if (TENSOR) cn1[ps] = __ldg(p.col1 + o);
          rn[ps] = __ldg(p.val + o);
        }
      };
      auto issue = [&](int s) {
        double* st = ring + (s % NBUF) * STG;
#pragma unroll
        for (int ps = 0; ps < GP; ps++) {
          const int k = tr + ps * OPP;
          const bool ok = s * KS + k < len;
          const double* src0 = p.P0 + (size_t)cn0[ps] * p.ld;
          const double* src1 = TENSOR ? p.P1 + (size_t)cn1[ps] * p.ld : nullptr;
#pragma unroll
          for (int j = 0; j < JP; j++) {
            const int pc = tq + 16 * j;
            if (pc < npc) {
              cp_async16(st + k * S + 2 * pc, src0 + 2 * pc, ok ? 16 : 0);
              if (TENSOR) cp_async16(st + (KS + k) * S + 2 * pc, src1 + 2 * pc, ok ? 16 : 0);
            }
          }
          if (tq == 0) {
            const double r = ok ? rn[ps] - p.mean : 0.0;
            st[(TENSOR ? 2 : 1) * KS * S + k] = r;
            if (aug) st[k * S + D] = r;
          }
        }
        cp_async_commit();
      };
      if (nst > 0) { load_meta(0); issue(0); }
      if (nst > 1) { load_meta(1); issue(1); } else cp_async_commit();
      if (nst > 2) load_meta(2);
      for (int s = 0; s < nst; s++) {
        cp_async_wait<1>();
        named_bar(1, NSY);
        if (s + 2 < nst) issue(s + 2); else cp_async_commit();
        if (s + 3 < nst) load_meta(s + 3);
        const double* buf = ring + (s % NBUF) * STG;
        const double* rs = buf + (TENSOR ? 2 : 1) * KS * S;
        int rem = len - s * KS;
        if (rem > KS) rem = KS;
        const int nk4 = (rem + 3) >> 2;
        K::warp_dispatch(warp, [&](auto w) { K::template compute<decltype(w)::value, TENSOR>(acc, buf, nk4, lane); });
        if (!aug && tid < DP) {
          for (int k = 0; k < nk4 * 4; k++) {
            double v = buf[k * S + tid];
            if (TENSOR) v *= buf[(KS + k) * S + tid];
            bsum = fma(v, rs[k], bsum);
          }
        }
      }
      cp_async_wait<0>();
      named_bar(1, NSY);  // ring free for the next item

      int nch = 0;
      if (split >= 0) {
        // park the partial in global memory in tile layout; only the last chunk of the row goes on to a finalise group
        nch = __ldg(p.split_nchunks + split);
        double* part = p.ws + (size_t)(__ldg(p.split_wsoff + split) + __ldg(p.item_chunk + item)) * PST;
        K::warp_dispatch(warp, [&](auto w) {
          constexpr int W = decltype(w)::value;
          static_for<C::ntiles(W)>([&](auto t) {
            constexpr int T = decltype(t)::value;
            using ti = TI<C, W, T>;
            *reinterpret_cast<double2*>(part + 64 * (tri(ti::I) + ti::J) + 2 * lane) = make_double2(acc[T][0], acc[T][1]);
          });
        });
        if (tid < DP) part[64 * NT + tid] = bsum;
        __threadfence();
        named_bar(1, NSY);
        if (tid == 0) {
          const int old = atomicAdd(p.split_counter + split, 1);
          const int last = (old == nch - 1);
          if (last) p.split_counter[split] = 0;
          hand[1] = last;
        }
        named_bar(1, NSY);
        const int last = hand[1];
        named_bar(1, NSY);  // hand[1] read by all before it can be rewritten
        if (!last) continue;
      }
      // hand the row to finalise group g
      const int g = fseq % G;
      double* gs = groups + (size_t)g * GSZ;
      mbar_wait(empty + g, ((fseq / G) & 1) ^ 1);
      if (split < 0) {
        K::warp_dispatch(warp, [&](auto w) {
          constexpr int W = decltype(w)::value;
          static_for<C::ntiles(W)>([&](auto t) {
            constexpr int T = decltype(t)::value;
            using ti = TI<C, W, T>;
            *reinterpret_cast<double2*>(gs + 64 * (tri(ti::I) + ti::J) + 2 * lane) = make_double2(acc[T][0], acc[T][1]);
          });
        });
        if (tid < DP) gs[64 * NT + NB * 64 + 4 * DP + tid] = bsum;  // bs[]
      }
      if (tid == 0) {
        Meta* m = reinterpret_cast<Meta*>(gs + 64 * NT + NB * 64 + 5 * DP + 8);
        m->lrow = lrow; m->split = split; m->nch = nch; m->exit = 0; m->item = item;
      }
      named_bar(1, NSY);  // all parking stores are done (bar.sync orders them before the arrive below)
      if (tid == 0) mbar_arrive(full + g);
      fseq++;
    }
    // drain: tell every group to exit once its buffer is free
    for (int g = 0; g < G; g++) {
      const int uses = (fseq + G - 1 - g) / G;  // rows this group received
      mbar_wait(empty + g, ((uses & 1) ^ 1));
      if (tid == 0) {
        Meta* m = reinterpret_cast<Meta*>(groups + (size_t)g * GSZ + 64 * NT + NB * 64 + 5 * DP + 8);
        m->exit = 1;
        mbar_arrive(full + g);
      }
    }
  }

  // ========================================================== finalise role ==========================================
  static __device__ void finalize_role(const RowParams& p, double* smem, int g, int ft) {
    const int lane = ft & 31, warp = ft >> 5;  // warp 0..FW-1 within the group
    const int D = p.D;
    const bool aug = use_aug(D);
    double* groups = smem + NBUF * STG;
    uint64_t* full = reinterpret_cast<uint64_t*>(groups + G * GSZ);
    uint64_t* empty = full + G;
    double* Tl = groups + (size_t)g * GSZ;
    double* WvT = Tl + 64 * NT;
    double* rhs = WvT + NB * 64;
    double* lmu = rhs + DP;
    double* ys = lmu + DP;
    double* xs = ys + DP;
    double* bs = xs + DP;
    double* ts = bs + DP;
    Meta* meta = reinterpret_cast<Meta*>(ts + 8);
    const int bid = 2 + g;  // named barrier of this group
    const int fo = 8 * (lane & 3) + (lane >> 2);
    bool bad = false;

    auto factor_diag = [&](int pb) {
      const int r = lane >> 2, q = lane & 3;
      const double2 av = *reinterpret_cast<const double2*>(Tl + 64 * (tri(pb) + pb) + 2 * lane);
      double a0 = av.x, a1 = av.y;
      double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
#pragma unroll 1
      for (int j = 7; j >= 0; j--) {
        const int jq = j >> 1;
        const double sel = (j & 1) ? a1 : a0;
        const double piv = __shfl_sync(0xffffffffu, sel, 4 * j + jq);
        if (!(piv > 0.0)) bad = true;
        const double sc = fast_rsqrt(piv);
        const double wij = __shfl_sync(0xffffffffu, sel, (lane & ~3) | jq) * sc;
        const double rj0 = __shfl_sync(0xffffffffu, a0, 4 * j + q) * sc;
        const double rj1 = __shfl_sync(0xffffffffu, a1, 4 * j + q) * sc;
        const double ej0 = __shfl_sync(0xffffffffu, e0, 4 * j + q) * sc;
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
      WvT[pb * 64 + (2 * q) * 8 + r] = e0;
      WvT[pb * 64 + (2 * q + 1) * 8 + r] = e1;
    };
    auto update_row = [&](int pb, int I, int j0, int j1) {
      const double* Pp = Tl + 64 * tri(pb) + fo;
      const double na0 = -Pp[64 * I], na1 = -Pp[64 * I + 32];
      double* trow = Tl + 64 * tri(I) + 2 * lane;
#pragma unroll 2
      for (int J = j0; J <= j1; J++) {
        const double b0 = Pp[64 * J], b1 = Pp[64 * J + 32];
        const double2 cv = *reinterpret_cast<const double2*>(trow + 64 * J);
        double c2[2] = {cv.x, cv.y};
        dmma884(c2, na0, b0);
        dmma884(c2, na1, b1);
        *reinterpret_cast<double2*>(trow + 64 * J) = make_double2(c2[0], c2[1]);
      }
    };
    auto backsub_step = [&](int J, bool update) {
      const int r8 = lane & 7;
      double y0 = 0.0, y1 = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        y0 = fma(WvT[J * 64 + k * 8 + r8], rhs[8 * J + k], y0);
        y1 = fma(WvT[J * 64 + (k + 1) * 8 + r8], rhs[8 * J + k + 1], y1);
      }
      if (lane < 8) ys[8 * J + r8] = y0 + y1;
      __syncwarp();
      if (update) {
        for (int c = lane; c < 8 * J; c += 32) {
          const double* tp = Tl + 64 * (tri(J) + (c >> 3)) + (c & 7);
          double s0 = rhs[c], s1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            s0 = fma(-tp[8 * k], ys[8 * J + k], s0);
            s1 = fma(-tp[8 * k + 8], ys[8 * J + k + 1], s1);
          }
          rhs[c] = s0 + s1;
        }
        __syncwarp();
      }
    };

    for (int n = 0;; n++) {
      mbar_wait(full + g, n & 1);
      if (meta->exit) break;
      const int lrow = meta->lrow, split = meta->split, nch = meta->nch;
      const int64_t slot = p.slot_base + lrow;
      // Λ·μ
      if (ft < DP) {
        double s = 0.0;
        if (p.lmu) {
          s = __ldg(p.lmu + ft);
        } else if (ft < D) {
          const double* mu = p.mu + slot * p.mu_ld;
          for (int i = 0; i < D; i++) s = fma(__ldg(p.Lambda + ft + (size_t)i * D), __ldg(mu + i), s);
        }
        lmu[ft] = s;
      }
      if (split >= 0) {
        // sum the parked partials of the row in chunk order (deterministic), straight into the tile buffer
        const double* base = p.ws + (size_t)__ldg(p.split_wsoff + split) * PST;
        for (int e = ft; e < (64 * NT + DP) / 2; e += NFI) {
          double2 s = make_double2(0.0, 0.0);
          for (int c = 0; c < nch; c++) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (size_t)c * PST) + e);
            s.x += v.x;
            s.y += v.y;
          }
          if (2 * e < 64 * NT) *reinterpret_cast<double2*>(Tl + 2 * e) = s;
          else *reinterpret_cast<double2*>(bs + (2 * e - 64 * NT)) = s;
        }
      }
      named_bar(bid, NFI);
      if (!aug && ft < D) rhs[ft] = fma(p.alpha, bs[ft], lmu[ft]);
      if (ft >= D && ft < DP) rhs[ft] = 0.0;
      // Λ* = Λ + αG, rhs from the augmented row, identity on the padding
      {
        const double alpha = p.alpha;
        const int r = lane >> 2, q = lane & 3;
#pragma unroll 4
        for (int t = warp; t < NT; t += FW) {
          int I, J;
          tri_coords(t, I, J);
          const int i = 8 * I + r, j = 8 * J + 2 * q;
          const double2 gv = *reinterpret_cast<const double2*>(Tl + 64 * t + 2 * lane);
          double2 v = __ldg(reinterpret_cast<const double2*>(p.LT + 64 * t + 2 * lane));
          if (i < D) {
            if (j < D) v.x = fma(alpha, gv.x, v.x);
            if (j + 1 < D) v.y = fma(alpha, gv.y, v.y);
          } else if (aug && i == D) {
            if (j < D) rhs[j] = fma(alpha, gv.x, lmu[j]);
            if (j + 1 < D) rhs[j + 1] = fma(alpha, gv.y, lmu[j + 1]);
          }
          *reinterpret_cast<double2*>(Tl + 64 * t + 2 * lane) = v;
        }
      }
      named_bar(bid, NFI);
      // blocked UL factorisation with look-ahead (see row_kernel.cuh)
      if (warp == 0) factor_diag(NB - 1);
      named_bar(bid, NFI);
      for (int pb = NB - 1; pb > 0; pb--) {
        {
          const double wa0 = WvT[pb * 64 + fo], wa1 = WvT[pb * 64 + 32 + fo];
          for (int J = warp; J < pb; J += FW) {
            double* tp = Tl + 64 * (tri(pb) + J);
            const double b0 = tp[fo], b1 = tp[32 + fo];
            double c2[2] = {0.0, 0.0};
            dmma884(c2, wa0, b0);
            dmma884(c2, wa1, b1);
            *reinterpret_cast<double2*>(tp + 2 * lane) = make_double2(c2[0], c2[1]);
          }
        }
        named_bar(bid, NFI);
        if (warp == 0) {
          update_row(pb, pb - 1, pb - 1, pb - 1);
          __syncwarp();
          factor_diag(pb - 1);
        } else {
          if (warp == 1) backsub_step(pb, true);
          constexpr int NWC = FW - 1;
          for (int nn = 0; nn < pb; nn++) {
            const int I = pb - 1 - nn;
            const int ph = nn % (2 * NWC);
            const int wo = 1 + (ph < NWC ? ph : 2 * NWC - 1 - ph);
            if (wo == warp) update_row(pb, I, 0, nn == 0 ? I - 1 : I);
          }
        }
        named_bar(bid, NFI);
      }
      if (bad && lane == 0) atomicOr(p.err_flag, 1);
      // forward substitution and store, warp 0 of the group
      if (warp == 0) {
        const int r8 = lane & 7;
        const int64_t grow = (int64_t)lrow * p.world + p.rank;
        backsub_step(0, false);
        for (int c = lane; c < DP; c += 32) {
          double z = 0.0;
          if (c < D) z = p.Z ? __ldg(p.Z + (size_t)slot * p.ld + c) : philox_normal(p.seed, p.sweep, p.entity, grow, c);
          ys[c] += z;
        }
        __syncwarp();
        for (int I = 0; I < NB; I++) {
          const int k = lane >> 2, q = lane & 3;
          double s0 = 0.0, s1 = 0.0;
          for (int c = q; c < 8 * I; c += 8) {
            const double* tp = Tl + 64 * (tri(I) + (c >> 3)) + 8 * k + (c & 7);
            s0 = fma(tp[0], xs[c], s0);
            s1 = fma(tp[4], xs[c + 4], s1);
          }
          double sacc = s0 + s1;
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          if (q == 0) ts[k] = ys[8 * I + k] - sacc;
          __syncwarp();
          double x0 = 0.0, x1 = 0.0;
#pragma unroll
          for (int kk = 0; kk < 8; kk += 2) {
            x0 = fma(WvT[I * 64 + r8 * 8 + kk], ts[kk], x0);
            x1 = fma(WvT[I * 64 + r8 * 8 + kk + 1], ts[kk + 1], x1);
          }
          if (lane < 8) xs[8 * I + r8] = x0 + x1;
          __syncwarp();
        }
        double* out = p.Uout + (size_t)slot * p.ld;
        for (int j = lane; j < p.ld; j += 32) out[j] = j < D ? xs[j] : 0.0;
#pragma unroll 1
        for (int r = 0; r < 8 && p.peer_out[r]; r++) {
          double* po = p.peer_out[r] + (size_t)slot * p.ld;
          for (int j = lane; j < p.ld; j += 32) po[j] = j < D ? xs[j] : 0.0;
        }
      }
      named_bar(bid, NFI);  // the buffer (tiles, meta) is free again
      if (ft == 0) mbar_arrive(empty + g);
    }
  }

  static __device__ void run(const RowParams& p, double* smem) {
    const int tid = threadIdx.x;
    double* groups = smem + NBUF * STG;
    uint64_t* full = reinterpret_cast<uint64_t*>(groups + G * GSZ);
    uint64_t* empty = full + G;
    if (tid == 0) {
      for (int g = 0; g < G; g++) {
        mbar_init(full + g, 1);
        mbar_init(empty + g, 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < NSY) syrk_role(p, smem, tid);
    else finalize_role(p, smem, (tid - NSY) / NFI, (tid - NSY) % NFI);
  }
};

template <class KW>
__global__ void __launch_bounds__(KW::NTHR, 1) row_kernel_ws(const RowParams p) {
  extern __shared__ __align__(16) double smem_dyn[];
  KW::run(p, smem_dyn);
}

}  // namespace bdf
