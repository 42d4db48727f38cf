// Streaming DMMA syrk over a rank's factor rows: the N, Σu, Σuuᵀ reductions of ConditionalNormalWishart
// (src/sampling.jl:117-119). Template on the same tile configuration as the row draw.
#pragma once
#include "row_kernel.cuh"

namespace bdf {

// DMMA accumulate over a stage laid out with the statistics kernel's own stage height
template <class K, int W>
__device__ __forceinline__ void stats_compute(double (&acc)[K::TPW][2], const double* buf, int nk4, int lane) {
  using C = typename K::C;
  constexpr int NF = C::nfrag(W);
  constexpr int NTW = C::ntiles(W);
  if constexpr (NTW > 0) {
    const double* base = buf + (lane & 3) * K::S + (lane >> 2);
    for (int k4 = 0; k4 < nk4; k4++) {
      double f[NF];
      static_for<NF>([&](auto r) {
        constexpr int R = decltype(r)::value;
        f[R] = base[k4 * 4 * K::S + 8 * FI<C, W, R>::blk];
      });
      static_for<NTW>([&](auto t) {
        constexpr int T = decltype(t)::value;
        using ti = TI<C, W, T>;
        dmma884(acc[T], f[ti::fa], f[ti::fb]);
      });
    }
  }
}

// ---- statistics: each CTA reduces a contiguous range of rows into a dense packed partial ---------------------
template <class K>
__global__ void __launch_bounds__(K::NTHR) stats_kernel(const double* __restrict__ U, const double* __restrict__ uhat, int ld, int D,
                                                        int64_t slot0, int64_t nrows, int64_t rows_per_blk, double* __restrict__ ws) {
  extern __shared__ __align__(16) double smem_dyn[];
  constexpr int KS = K::SKS, S = K::S, DP = K::DP, OPP = K::OPP, PASSES = K::SPASSES, JP = K::JP, TPW = K::TPW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tr = tid >> 4, tq = tid & 15;
  const bool aug = use_aug(D);
  double* bufs = smem_dyn;
  double* rss = smem_dyn + 2 * KS * S;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_blk;
  int64_t r1 = r0 + rows_per_blk;
  if (r1 > nrows) r1 = nrows;
  const int64_t len = r1 > r0 ? r1 - r0 : 0;
  const int ppr = ld >> 1;

  double acc[TPW][2];
#pragma unroll
  for (int t = 0; t < TPW; t++) acc[t][0] = acc[t][1] = 0.0;
  double bsum = 0.0;

  auto fetch = [&](int64_t o0, typename K::Pre& pre) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++) {
      const int64_t o = o0 + tr + ps * OPP;
      const bool ok = o < r1;
      pre.r[ps] = ok ? 1.0 : 0.0;
      const double2* src = reinterpret_cast<const double2*>(U + (size_t)(slot0 + o) * ld);
      const double2* sub = uhat ? reinterpret_cast<const double2*>(uhat + (size_t)(slot0 + o) * ld) : nullptr;
#pragma unroll
      for (int j = 0; j < JP; j++) {
        const int pc = tq + 16 * j;
        double2 x = make_double2(0.0, 0.0);
        if (ok && pc < ppr) {
          x = src[pc];
          if (sub) {
            const double2 y = sub[pc];
            x.x -= y.x;
            x.y -= y.y;
          }
        }
        pre.a[ps][j] = x;
      }
    }
  };

  const int nst = (int)((len + KS - 1) / KS);
  typename K::Pre pre;
  if (nst > 0) {
    fetch(r0, pre);
    K::store_stage(D, bufs, rss, tr, tq, pre, aug);
  }
  __syncthreads();
  for (int s = 0; s < nst; s++) {
    const double* buf = bufs + (s & 1) * KS * S;
    const double* rs = rss + (s & 1) * KS;
    const bool more = s + 1 < nst;
    if (more) fetch(r0 + (int64_t)(s + 1) * KS, pre);
    int64_t rem = len - (int64_t)s * KS;
    if (rem > KS) rem = KS;
    const int nk4 = (int)((rem + 3) >> 2);
    K::warp_dispatch(warp, [&](auto w) { stats_compute<K, decltype(w)::value>(acc, buf, nk4, lane); });
    if (!aug && tid < DP)
      for (int k = 0; k < nk4 * 4; k++) bsum = fma(buf[k * S + tid], rs[k], bsum);
    if (more) K::store_stage(D, bufs + ((s + 1) & 1) * KS * S, rss + ((s + 1) & 1) * KS, tr, tq, pre, aug);
    __syncthreads();
  }
  double* part = ws + (size_t)blockIdx.x * tri(D + 1);
  K::warp_dispatch(warp, [&](auto w) {
    K::template for_acc<decltype(w)::value>(acc, lane, [&](int, int i, int j, double& v) {
      if (j <= i && (i < D || (i == D && j < D))) part[tri(i) + j] = v;
    });
  });
  if (!aug && tid < D) part[tri(D) + tid] = bsum;
}

}  // namespace bdf
