// One padded latent dimension (-DBDF_DP=...) of the row-draw and statistics kernels; build.py compiles this file
// once per DP so the 16 instances build in parallel.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "../../include/bdf_b200.h"
#include "engine.cuh"
#include "row_kernel.cuh"
#include "row_kernel_ws.cuh"
#include "stats_kernel.cuh"

#ifndef BDF_DP
#error "compile with -DBDF_DP=<padded latent dimension>"
#endif

using namespace bdf;

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
static constexpr int kDP = BDF_DP;
#ifndef BDF_NW_BIG
#define BDF_NW_BIG 4
#endif
#ifndef BDF_NW1_MAXDP
#define BDF_NW1_MAXDP 32  // largest padded D that runs one warp per row
#endif
static constexpr int kNW = kDP <= BDF_NW1_MAXDP ? 1 : (kDP <= 64 ? 4 : BDF_NW_BIG);

// 64 < D <= 104, 2-mode relations: the persistent warp-specialised kernel (5 rows in flight per SM instead of 4, row_kernel_ws.cuh) is
// compiled in and parity-tested, but opt-in (BDF_ROWS_WS=1 at bdf_create): on C2 it measured slower than one CTA per row
#ifndef BDF_USE_WS
#define BDF_USE_WS 1
#endif
static constexpr bool kWS = BDF_USE_WS && kDP > 64 && kDP <= 104;  // three 8·tri(DP/8)·64-byte tile slots + two gather rings must fit 227 KB

static int launch_rows_ws(bdf_t* h, const RowParams& p0, int n_items) {
  if constexpr (kWS) {
    using W = RowKernelWS<kDP, false>;
    static_assert(!kWS || W::SMEM_BYTES <= 227 * 1024, "shared memory of the persistent kernel");
    if (!(h->smem_optin & BDF_OPTIN_ROWS_WS)) {
      CU(cudaFuncSetAttribute(row_kernel_ws<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::SMEM_BYTES));
      h->smem_optin |= BDF_OPTIN_ROWS_WS;
    }
    if (n_items > 0) {
      RowParams p = p0;
      p.work_counter = h->work_counter;
      p.n_items = n_items;
      CU(cudaMemsetAsync(h->work_counter, 0, sizeof(int), h->stream));
      const int grid = std::min(h->num_sms, (n_items + W::NSG - 1) / W::NSG);
      row_kernel_ws<W><<<grid, W::NTHR, W::SMEM_BYTES, h->stream>>>(p);
      h->launches++;
      CU(cudaGetLastError());
    }
  }
  return BDF_OK;
}

// warp-per-row kernel on 3-mode tensors: both partners are staged, so a 16-observation stage ring costs 28 KB per one-warp CTA and only 7
// of them fit an SM (ncu: 11 % of the warp slots active); 8-observation stages double the rows in flight
#ifndef BDF_TKS1
#define BDF_TKS1 8
#endif
template <bool TENSOR>
static int launch_rows_t(bdf_t* h, const RowParams& p, int n_items) {
  if constexpr (kWS && !TENSOR) {
    if (h->use_ws) return launch_rows_ws(h, p, n_items);
  }
  using K = RowKernel<kDP, kNW, TENSOR, (kNW == 1 && TENSOR) ? BDF_TKS1 : 0, (kNW == 1 && TENSOR) ? 3 : 0>;
  const uint32_t bit = TENSOR ? BDF_OPTIN_ROWS_TENSOR : BDF_OPTIN_ROWS;
  if (!(h->smem_optin & bit)) {
    CU(cudaFuncSetAttribute(row_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
    h->smem_optin |= bit;
  }
  if (n_items > 0) {
    row_kernel<K><<<n_items, K::NTHR, K::SMEM_BYTES, h->stream>>>(p);
    h->launches++;
    CU(cudaGetLastError());
  }
  return BDF_OK;
}

int CAT(bdf_launch_rows_, BDF_DP)(bdf_t* h, const RowParams& p, int n_items, bool tensor) {
  return tensor ? launch_rows_t<true>(h, p, n_items) : launch_rows_t<false>(h, p, n_items);
}

// returns the number of partials written to h->ws (each tri(D+1) doubles), or a negative error
int CAT(bdf_launch_stats_, BDF_DP)(bdf_t* h, const double* U, const double* uhat, int64_t slot0, int64_t nrows) {
  using K = RowKernel<kDP, kNW, false>;
  const size_t smem = sizeof(double) * K::SBUFSZ;
  if (!(h->smem_optin & BDF_OPTIN_STATS)) {
    CU(cudaFuncSetAttribute(stats_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->smem_optin |= BDF_OPTIN_STATS;
  }
  int64_t nblk = (nrows + 4 * K::SKS - 1) / (4 * K::SKS);
  if (nblk > 296) nblk = 296;
  if (nblk < 1) nblk = 1;
  const int64_t rpb = (nrows + nblk - 1) / nblk;
  const size_t need = sizeof(double) * (size_t)nblk * tri(h->D + 1);
  if (need > h->ws_bytes) {
    h->err = "workspace too small for statistics";
    return BDF_ERR_STATE;
  }
  stats_kernel<K><<<(int)nblk, K::NTHR, smem, h->stream>>>(U, uhat, h->ld, h->D, slot0, nrows, rpb, h->ws);
  CU(cudaGetLastError());
  h->launches += 1;
  return (int)nblk;
}

int64_t CAT(bdf_pst_, BDF_DP)() { return RowKernel<kDP, kNW, false>::PST; }
