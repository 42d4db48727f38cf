// libbdf_b200.so — C ABI (include/bdf_b200.h) over the sm_100a kernels. No CPU fallback: every numeric entry
// launches device kernels; host code only validates arguments, moves buffers and builds launch metadata.
#include "../../include/bdf_b200.h"
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <numeric>

#include "../../include/bdf_b200.h"
#include "nw_kernels.cuh"

using namespace bdf;

static thread_local std::string g_create_err;

// ---- kernel dispatch over the padded latent dimension (instances live in row_inst.cu, one object per DP) ----------
#define DP_CASES(X) X(8) X(16) X(24) X(32) X(40) X(48) X(56) X(64) X(72) X(80) X(88) X(96) X(104) X(112) X(120) X(128)
#define X(dp)                                                                                           \
  int bdf_launch_rows_##dp(bdf_t* h, const RowParams& p, int n_items, bool tensor);                     \
  int bdf_launch_stats_##dp(bdf_t* h, const double* U, const double* uhat, int64_t slot0, int64_t nrows);     \
  int64_t bdf_pst_##dp();
DP_CASES(X)
#undef X


namespace {

template <class T>
int dev_alloc(bdf_t* h, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CU(cudaMalloc((void**)p, n * sizeof(T)));
  return BDF_OK;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

int launch_rows(bdf_t* h, const RowParams& p, int n_items, bool tensor) {
  switch (h->DP) {
#define X(dp) \
  case dp:    \
    return bdf_launch_rows_##dp(h, p, n_items, tensor);
    DP_CASES(X)
#undef X
  }
  FAIL(BDF_ERR_INVALID, "unsupported num_latent");
}

int launch_stats_partials(bdf_t* h, const double* U, const double* uhat, int64_t slot0, int64_t nrows) {
  switch (h->DP) {
#define X(dp) \
  case dp:    \
    return bdf_launch_stats_##dp(h, U, uhat, slot0, nrows);
    DP_CASES(X)
#undef X
  }
  FAIL(BDF_ERR_INVALID, "unsupported num_latent");
}

int64_t pst_of(int DP) {
  switch (DP) {
#define X(dp) \
  case dp:    \
    return bdf_pst_##dp();
    DP_CASES(X)
#undef X
  }
  return 0;
}

// ---- ingestion kernels ---------------------------------------------------------------------------------------
// slot of 0-based global row i: rows are dealt cyclically to ranks (src/sampling.jl:154), each rank's rows contiguous
__host__ __device__ inline int64_t slot_of(int64_t i, int world, int64_t nper) { return (i % world) * nper + i / world; }

__global__ void make_keys_kernel(const int64_t* ids, int64_t nnz, int64_t N, int world, int64_t nper, const int32_t* __restrict__ tab,
                                 uint32_t* keys, uint32_t* idx, int* bad) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nnz; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[o];
    if (id < 1 || id > N) {
      *bad = 1;
      keys[o] = 0;
    } else {
      keys[o] = tab ? (uint32_t)tab[id - 1] : (uint32_t)slot_of(id - 1, world, nper);
    }
    idx[o] = (uint32_t)o;
  }
}

__global__ void count_rows_kernel(const uint32_t* keys, int64_t nnz, int64_t slot0, int64_t nrows, unsigned long long* counts) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nnz; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = (int64_t)keys[o] - slot0;
    if (r >= 0 && r < nrows) atomicAdd(counts + r, 1ULL);
  }
}

__global__ void lower_bound_kernel(const uint32_t* keys, int64_t nnz, uint32_t target, int64_t* out) {
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

__global__ void gather_obs_kernel(const uint32_t* perm, int64_t base, int64_t n, const int64_t* ids_other, int world, int64_t nper_other,
                                  const int32_t* __restrict__ tab, int32_t* col) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = ids_other[perm[base + o]] - 1;
    col[o] = tab ? tab[i] : (int32_t)slot_of(i, world, nper_other);
  }
}

__global__ void gather_val_kernel(const uint32_t* perm, int64_t base, int64_t n, const double* vals, double* out) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) out[o] = vals[perm[base + o]];
}

__global__ void ids_to_slots_kernel(const int64_t* ids, int64_t n, int64_t N, int world, int64_t nper, const int32_t* __restrict__ tab, int32_t* out,
                                    int* bad) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[o];
    if (id < 1 || id > N) { *bad = 1; out[o] = 0; } else out[o] = tab ? tab[id - 1] : (int32_t)slot_of(id - 1, world, nper);
  }
}

// rows of a D×N column-major host-order staging matrix → slot-major ld-pitched device buffer (and back) through an explicit row → slot map
__global__ void scatter_rows_kernel(const double* __restrict__ stage, const int32_t* __restrict__ slot_of_row, int64_t N, int D, int ld, double* __restrict__ dev) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * D; e += (int64_t)gridDim.x * blockDim.x)
    dev[(size_t)slot_of_row[e / D] * ld + e % D] = stage[e];
}
__global__ void gather_rows_kernel(const double* __restrict__ dev, const int32_t* __restrict__ slot_of_row, int64_t N, int D, int ld, double* __restrict__ stage) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * D; e += (int64_t)gridDim.x * blockDim.x)
    stage[e] = dev[(size_t)slot_of_row[e / D] * ld + e % D];
}

__global__ void set_identity_kernel(double* A, int D, double v) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D * D; e += gridDim.x * blockDim.x) A[e] = (e % D == e / D) ? v : 0.0;
}

inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// host↔device copies between Julia's D×N column-major matrix and the slot-major, ld-pitched device buffer
int copy_rows_h2d(bdf_t* h, const EntityS& e, const double* host, double* dev) {
  const int D = h->D, W = h->world;
  if (e.slot_of_row) {
    int rc = bdf_ensure_arena(h, sizeof(double) * (size_t)e.N * D);
    if (rc) return rc;
    double* stage = reinterpret_cast<double*>(h->arena);
    CU(cudaMemcpyAsync(stage, host, sizeof(double) * (size_t)e.N * D, cudaMemcpyHostToDevice, h->stream));
    scatter_rows_kernel<<<grid_for(e.N * D), 256, 0, h->stream>>>(stage, e.slot_of_row, e.N, D, h->ld, dev);
    CU(cudaGetLastError());
    return BDF_OK;
  }
  for (int r = 0; r < W; r++) {
    const int64_t cnt = (e.N - r + W - 1) / W;  // rows r, r+W, ...
    if (cnt <= 0) continue;
    CU(cudaMemcpy2DAsync(dev + (size_t)r * e.Nper * h->ld, sizeof(double) * h->ld, host + (size_t)r * D, sizeof(double) * D * W,
                         sizeof(double) * D, (size_t)cnt, cudaMemcpyHostToDevice, h->stream));
  }
  return BDF_OK;
}
int copy_rows_d2h(bdf_t* h, const EntityS& e, const double* dev, double* host) {
  const int D = h->D, W = h->world;
  if (e.slot_of_row) {
    int rc = bdf_ensure_arena(h, sizeof(double) * (size_t)e.N * D);
    if (rc) return rc;
    double* stage = reinterpret_cast<double*>(h->arena);
    gather_rows_kernel<<<grid_for(e.N * D), 256, 0, h->stream>>>(dev, e.slot_of_row, e.N, D, h->ld, stage);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, stage, sizeof(double) * (size_t)e.N * D, cudaMemcpyDeviceToHost, h->stream));
    return BDF_OK;
  }
  for (int r = 0; r < W; r++) {
    const int64_t cnt = (e.N - r + W - 1) / W;
    if (cnt <= 0) continue;
    CU(cudaMemcpy2DAsync(host + (size_t)r * D, sizeof(double) * D * W, dev + (size_t)r * e.Nper * h->ld, sizeof(double) * h->ld,
                         sizeof(double) * D, (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
  }
  return BDF_OK;
}

int check_err_flag(bdf_t* h) {
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (flag) {
    CU(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
    FAIL(BDF_ERR_NUMERIC, flag & 1 ? "row draw: precision matrix not positive definite" : "Normal-Wishart draw: matrix not positive definite");
  }
  return BDF_OK;
}

// a draw started by bdf_nw_sample_async writes e.mu / e.Lambda and uses h->scratch on the side stream: order it before main-stream work
int join_draw(bdf_t* h, EntityS& e) {
  if (e.draw_pending) { CU(cudaStreamWaitEvent(h->stream, e.ev_done, 0)); e.draw_pending = false; }
  return BDF_OK;
}

int ensure_ws(bdf_t* h, size_t bytes) {
  if (bytes <= h->ws_bytes) return BDF_OK;
  if (h->ws) {
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaFree(h->ws));
    h->ws = nullptr;
    h->ws_bytes = 0;
  }
  CU(cudaMalloc((void**)&h->ws, bytes));
  h->ws_bytes = bytes;
  return BDF_OK;
}

// Split heavy rows into chunks, order work items by cost (heaviest first) and upload the list. `rps` holds one host row_ptr per
// relation the entity takes part in (src/sampling.jl:266-283 sums the relations' contributions to a row): a row whose
// observations come as several items — chunks of a long row and/or one relation each — is a "split" row whose partials are
// added in item order (relation order, then chunk order) by the last item to finish.
int build_work_list(bdf_t* h, ModeIndex& mi, const std::vector<const std::vector<int64_t>*>& rps, int64_t real_rows) {
  // observations per chunk of a split row (a multiple of every KS): 8192, less when this rank's share of the table is small (more
  // GPUs), so that the heavy rows still break into enough items to fill the 148 SMs several times over
  int64_t total = 0;
  for (auto* rp : rps) total += (*rp)[real_rows] - (*rp)[0];
  const int64_t CH = std::min<int64_t>(8192, std::max<int64_t>(1024, (total / 4096 + 15) / 16 * 16));
  const bool multi = rps.size() > 1;
  std::vector<int32_t> irow, ilen, isplit, ichunk, irel, snch, sgsize;
  std::vector<int64_t> ibeg, swoff, sgcoff;
  int64_t slots = 0, groups = 0;
  struct Piece { int64_t beg, len; int rel; };
  std::vector<Piece> pieces;
  for (int64_t r = 0; r < real_rows; r++) {
    pieces.clear();
    for (size_t s = 0; s < rps.size(); s++) {
      const int64_t b = (*rps[s])[r], n = (*rps[s])[r + 1] - b;
      if (n == 0) continue;
      if (n <= CH + CH / 2) {
        pieces.push_back({b, n, (int)s});
      } else {
        const int64_t nch = (n + CH - 1) / CH;
        for (int64_t c = 0; c < nch; c++) pieces.push_back({b + c * CH, std::min(n - c * CH, CH), (int)s});
      }
    }
    if (pieces.empty()) pieces.push_back({(*rps[0])[r], 0, 0});  // no observations anywhere: the row is drawn from the prior
    const int sid = pieces.size() > 1 ? (int)snch.size() : -1;
    if (sid >= 0) {
      const int64_t nch = (int64_t)pieces.size();
      // partials are added up in two levels once a row has many of them: groups of ~sqrt(nch) consecutive chunks
      const int64_t G = nch <= 32 ? nch : (int64_t)std::ceil(std::sqrt((double)nch));
      snch.push_back((int32_t)nch);
      swoff.push_back(slots);
      sgsize.push_back((int32_t)G);
      sgcoff.push_back(groups);
      slots += nch;
      groups += (nch + G - 1) / G;
    }
    for (size_t c = 0; c < pieces.size(); c++) {
      irow.push_back((int32_t)r); ibeg.push_back(pieces[c].beg); ilen.push_back((int32_t)pieces[c].len);
      isplit.push_back(sid); ichunk.push_back((int32_t)c); irel.push_back(pieces[c].rel);
    }
  }
  const size_t ni = irow.size();
  std::vector<int32_t> order(ni);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ilen[a] > ilen[b]; });
  auto permute32 = [&](std::vector<int32_t>& v) { std::vector<int32_t> t(ni); for (size_t i = 0; i < ni; i++) t[i] = v[order[i]]; v.swap(t); };
  auto permute64 = [&](std::vector<int64_t>& v) { std::vector<int64_t> t(ni); for (size_t i = 0; i < ni; i++) t[i] = v[order[i]]; v.swap(t); };
  permute32(irow); permute32(ilen); permute32(isplit); permute32(ichunk); permute32(irel); permute64(ibeg);
  mi.n_items = (int)ni;
  mi.n_split = (int)snch.size();
  mi.chunk_slots = slots;
  mi.ws_slots = slots + groups;
  int rc;
  if ((rc = dev_alloc(h, &mi.item_row, ni))) return rc;
  if ((rc = dev_alloc(h, &mi.item_beg, ni))) return rc;
  if ((rc = dev_alloc(h, &mi.item_len, ni))) return rc;
  if ((rc = dev_alloc(h, &mi.item_split, ni))) return rc;
  if ((rc = dev_alloc(h, &mi.item_chunk, ni))) return rc;
  if (multi && (rc = dev_alloc(h, &mi.item_rel, ni))) return rc;
  if ((rc = dev_alloc(h, &mi.split_nchunks, snch.size()))) return rc;
  if ((rc = dev_alloc(h, &mi.split_wsoff, snch.size()))) return rc;
  if ((rc = dev_alloc(h, &mi.split_counter, snch.size()))) return rc;
  if ((rc = dev_alloc(h, &mi.split_gsize, snch.size()))) return rc;
  if ((rc = dev_alloc(h, &mi.split_gcoff, snch.size()))) return rc;
  if ((rc = dev_alloc(h, &mi.group_counter, (size_t)groups))) return rc;
  if (ni) {
    CU(cudaMemcpyAsync(mi.item_row, irow.data(), ni * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.item_beg, ibeg.data(), ni * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.item_len, ilen.data(), ni * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.item_split, isplit.data(), ni * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.item_chunk, ichunk.data(), ni * 4, cudaMemcpyHostToDevice, h->stream));
    if (multi) CU(cudaMemcpyAsync(mi.item_rel, irel.data(), ni * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (!snch.empty()) {
    CU(cudaMemcpyAsync(mi.split_nchunks, snch.data(), snch.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.split_wsoff, swoff.data(), swoff.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.split_gsize, sgsize.data(), sgsize.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(mi.split_gcoff, sgcoff.data(), sgcoff.size() * 8, cudaMemcpyHostToDevice, h->stream));
  }
  CU(cudaMemsetAsync(mi.group_counter, 0, std::max<size_t>(1, (size_t)groups) * sizeof(int), h->stream));
  CU(cudaMemsetAsync(mi.split_counter, 0, std::max<size_t>(1, snch.size()) * sizeof(int), h->stream));
  CU(cudaStreamSynchronize(h->stream));  // host vectors die here
  return ensure_ws(h, std::max<size_t>(sizeof(double) * (size_t)(slots + groups) * h->pst, sizeof(double) * 296 * (size_t)tri(h->D + 1)));
}

void free_work_list(ModeIndex& mi) {
  cudaFree(mi.item_row); cudaFree(mi.item_beg); cudaFree(mi.item_len); cudaFree(mi.item_split); cudaFree(mi.item_chunk); cudaFree(mi.item_rel);
  cudaFree(mi.split_nchunks); cudaFree(mi.split_wsoff); cudaFree(mi.split_counter); cudaFree(mi.split_gsize); cudaFree(mi.split_gcoff); cudaFree(mi.group_counter);
  mi.split_gsize = nullptr; mi.split_gcoff = nullptr; mi.group_counter = nullptr;
  mi.item_row = mi.item_len = mi.item_split = mi.item_chunk = mi.item_rel = mi.split_nchunks = nullptr;
  mi.item_beg = mi.split_wsoff = nullptr;
  mi.split_counter = nullptr;
  mi.n_items = mi.n_split = 0;
}

void prep_lambda(bdf_t* h, const double* Lambda, const double* mu, int D, int DP, double* LT, double* lmu);

int sample_entity(bdf_t* h, int entity, const double* mu_dev, int64_t mu_ld, const double* Lambda_dev, const double* Z_dev,
                  long long* dbg = nullptr) {
  EntityS& e = h->ents[entity];
  if (e.uses.empty()) FAIL(BDF_ERR_STATE, "entity takes part in no relation");
  if (e.uses.size() > BDF_MAX_USES) FAIL(BDF_ERR_INVALID, "an entity may take part in at most 6 relations");
  { int rcj = join_draw(h, e); if (rcj) return rcj; }
  // an entity in several relations (src/sampling.jl:251-289) runs off a merged work list, rebuilt when a relation was added
  const ModeIndex* wl = &h->rels[e.uses[0].first].modes[e.uses[0].second];
  if (e.uses.size() > 1) {
    if (e.merged_uses != e.uses.size()) {
      free_work_list(e.merged);
      std::vector<const std::vector<int64_t>*> rps;
      for (auto& u : e.uses) rps.push_back(&h->rels[u.first].modes[u.second].h_row_ptr);
      int rc = build_work_list(h, e.merged, rps, e.nlocal);
      if (rc) return rc;
      e.merged_uses = e.uses.size();
    }
    wl = &e.merged;
  }
  RowParams p{};
  p.item_row = wl->item_row; p.item_beg = wl->item_beg; p.item_len = wl->item_len; p.item_split = wl->item_split; p.item_chunk = wl->item_chunk;
  p.item_rel = wl->item_rel;
  p.split_nchunks = wl->split_nchunks; p.split_wsoff = wl->split_wsoff; p.split_counter = wl->split_counter; p.ws = h->ws;
  p.split_gsize = wl->split_gsize; p.split_gcoff = wl->split_gcoff; p.group_counter = wl->group_counter; p.gslot_base = wl->chunk_slots;
  bool tensor = false;
  for (auto& u : e.uses) tensor = tensor || h->rels[u.first].K > 2;
  for (size_t i = 0; i < e.uses.size(); i++) {
    RelationS& rel = h->rels[e.uses[i].first];
    ModeIndex& mi = rel.modes[e.uses[i].second];
    RelTab& t = p.rt[i];
    t.col0 = mi.col[0]; t.col1 = rel.K > 2 ? mi.col[1] : nullptr;
    t.val = rel.F ? mi.val_adj : mi.val;  // with relation features the offset is per observation: linear_values[getI(...)], src/sampling.jl:273
    t.P0 = h->ents[mi.other_entity[0]].U;
    t.P1 = rel.K > 2 ? h->ents[mi.other_entity[1]].U : (tensor ? h->ones : nullptr);
    t.alpha = rel.alpha; t.mean = rel.F ? 0.0 : rel.mean;
  }
  p.ld = h->ld; p.Uout = e.U; p.slot_base = (int64_t)h->rank * e.Nper; p.row_of_slot = e.row_of_slot;
  { int np = 0; for (int r = 0; r < 8; r++) if (e.peerU[r]) p.peer_out[np++] = e.peerU[r]; }
  p.Lambda = Lambda_dev; p.mu = mu_dev; p.mu_ld = mu_ld; p.Z = Z_dev;
  p.LT = h->lt; p.lmu = mu_ld ? nullptr : h->lt + 64 * (h->DP / 8) * (h->DP / 8 + 1) / 2;
  prep_lambda(h, Lambda_dev, mu_ld ? nullptr : mu_dev, h->D, h->DP, h->lt, h->lt + 64 * (h->DP / 8) * (h->DP / 8 + 1) / 2);
  h->launches++;
  p.D = h->D; p.rank = h->rank; p.world = h->world;
  p.seed = h->seed; p.sweep = h->sweep; p.entity = entity; p.err_flag = h->err_flag;
#ifdef BDF_DEBUG
  p.dbg = dbg;
  { const char* f = getenv("BDF_DEBUG_FLAGS"); p.flags = f ? atoi(f) : 0; }
#else
  (void)dbg;
#endif
  return launch_rows(h, p, wl->n_items, tensor);
}

void prep_lambda(bdf_t* h, const double* Lambda, const double* mu, int D, int DP, double* LT, double* lmu) {
  prep_lambda_kernel<<<8, 256, 0, h->stream>>>(Lambda, mu, D, DP, LT, lmu);
}


// Training residuals for sample_alpha (src/macau.jl:85-87): one warp per work item of the relation's first mode (rows, long rows in
// chunks), err = Σ_k Π_m U_m[k] + mean − value per observation, squared and summed per item; the per-item sums are then added in
// item order by one block, so the result does not depend on scheduling.
__global__ void sse_items_kernel(const int32_t* __restrict__ item_row, const int64_t* __restrict__ item_beg, const int32_t* __restrict__ item_len,
                                 int n_items, const int32_t* __restrict__ col0, const int32_t* __restrict__ col1, const double* __restrict__ val,
                                 const double* __restrict__ U, const double* __restrict__ P0, const double* __restrict__ P1, int64_t slot_base, int ld,
                                 int D, double mean, double* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int item = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (item >= n_items) return;
  const double* u = U + (size_t)(slot_base + item_row[item]) * ld;
  const int64_t b = item_beg[item];
  const int n = item_len[item];
  double acc = 0.0;
  for (int o = 0; o < n; o++) {
    const double* p0 = P0 + (size_t)__ldg(col0 + b + o) * ld;
    const double* p1 = P1 ? P1 + (size_t)__ldg(col1 + b + o) * ld : nullptr;
    double s = 0.0;
    for (int k = lane; k < D; k += 32) {
      double t = u[k] * __ldg(p0 + k);
      if (p1) t *= __ldg(p1 + k);
      s += t;
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    const double err = s + mean - __ldg(val + b + o);
    acc = fma(err, err, acc);
  }
  if (lane == 0) partial[item] = acc;
}

// res[perm[o]] = value − udot − mean for every training observation (table order) — the residual of sample_beta_rel, src/sampling.jl:327
__global__ void residual_items_kernel(const int32_t* __restrict__ item_row, const int64_t* __restrict__ item_beg, const int32_t* __restrict__ item_len,
                                      int n_items, const int32_t* __restrict__ col0, const int32_t* __restrict__ col1, const double* __restrict__ val,
                                      const uint32_t* __restrict__ perm, const double* __restrict__ U, const double* __restrict__ P0,
                                      const double* __restrict__ P1, int64_t slot_base, int ld, int D, double mean, double* __restrict__ res) {
  const int lane = threadIdx.x & 31;
  const int item = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (item >= n_items) return;
  const double* u = U + (size_t)(slot_base + item_row[item]) * ld;
  const int64_t b = item_beg[item];
  const int n = item_len[item];
  for (int o = 0; o < n; o++) {
    const double* p0 = P0 + (size_t)__ldg(col0 + b + o) * ld;
    const double* p1 = P1 ? P1 + (size_t)__ldg(col1 + b + o) * ld : nullptr;
    double s = 0.0;
    for (int k = lane; k < D; k += 32) {
      double t = u[k] * __ldg(p0 + k);
      if (p1) t *= __ldg(p1 + k);
      s += t;
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    if (lane == 0) res[perm[b + o]] = __ldg(val + b + o) - s - mean;
  }
}

// out[0] = Σ partial (fixed order: thread-strided sums, then a tree over the block), out[1] = optional alpha draw
__global__ void sse_reduce_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// sample_alpha (src/sampling.jl:129-134): SW = inv(inv(lambda0) + err'err); alpha = rand(Wishart(nu0 + n, SW))[1] = SW·chi2(nu0 + n)
__global__ void alpha_draw_kernel(double sse, double n, double lambda0, double nu0, double chi2_inj, uint64_t seed, uint64_t sweep, uint32_t stream,
                                  double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double SW = 1.0 / (1.0 / lambda0 + sse);
    const double c2 = chi2_inj == chi2_inj ? chi2_inj : 2.0 * gamma_mt(0.5 * (nu0 + n), seed, sweep, stream, 0);
    out[0] = SW * c2;
  }
}

}  // namespace

// residuals of relation `rel` in table order into r.res (world == 1)
int bdf_relation_residuals(bdf_t* h, int rel) {
  RelationS& r = h->rels[rel];
  ModeIndex& mi = r.modes[0];
  EntityS& e = h->ents[r.entity_of_mode[0]];
  if (mi.n_items > 0) {
    const int wpb = 8;
    residual_items_kernel<<<(mi.n_items + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(
        mi.item_row, mi.item_beg, mi.item_len, mi.n_items, mi.col[0], mi.col[1], mi.val, mi.perm, e.U, h->ents[mi.other_entity[0]].U,
        r.K > 2 ? h->ents[mi.other_entity[1]].U : nullptr, (int64_t)h->rank * e.Nper, h->ld, h->D, r.mean, r.res);
    CU(cudaGetLastError());
    h->launches++;
  }
  return BDF_OK;
}

// statistics of an arbitrary row-major (rows × ld) buffer into a [count, colsum(D), Gram(D×D)] stats block (features.cu uses it for betaᵀbeta)
int bdf_stats_of(bdf_t* h, const double* X, const double* sub, int64_t slot0, int64_t nrows, double* stats) {
  const int nblk = launch_stats_partials(h, X, sub, slot0, nrows);
  if (nblk < 0) return nblk;
  stats_reduce_kernel<<<(tri(h->D + 1) + 127) / 128, 128, 0, h->stream>>>(h->ws, nblk, h->D, (double)nrows, stats);
  h->launches++;
  CU(cudaGetLastError());
  return BDF_OK;
}
int bdf_check_err_flag(bdf_t* h) { return check_err_flag(h); }
int bdf_join_side(bdf_t* h) {
  for (auto& e : h->ents) { int rc = join_draw(h, e); if (rc) return rc; }
  return BDF_OK;
}
int bdf_copy_rows_h2d_impl(bdf_t* h, int entity, const double* host, double* dev) { return copy_rows_h2d(h, h->ents[entity], host, dev); }
int bdf_ensure_arena(bdf_t* h, size_t bytes) {
  if (bytes <= h->arena_bytes) return BDF_OK;
  if (h->arena) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->arena)); h->arena = nullptr; h->arena_bytes = 0; }
  bytes = bytes + bytes / 4;
  CU(cudaMalloc((void**)&h->arena, bytes));
  h->arena_bytes = bytes;
  return BDF_OK;
}
int bdf_ensure_arena2(bdf_t* h, size_t bytes) {
  if (bytes <= h->arena2_bytes) return BDF_OK;
  if (h->arena2) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->arena2)); h->arena2 = nullptr; h->arena2_bytes = 0; }
  bytes = bytes + bytes / 4;
  CU(cudaMalloc((void**)&h->arena2, bytes));
  h->arena2_bytes = bytes;
  return BDF_OK;
}
int bdf_sample_entity_impl(bdf_t* h, int entity, const double* mu_dev, int64_t mu_ld, const double* Lambda_dev, const double* Z_dev);
void bdf_dense_teardown(bdf_t* h);
int bdf_refresh_relation_offsets(bdf_t* h, int rel);

namespace {

int stats_entity(bdf_t* h, int entity) {
  EntityS& e = h->ents[entity];
  const int nblk = launch_stats_partials(h, e.U, nullptr, (int64_t)h->rank * e.Nper, e.nlocal);
  if (nblk < 0) return nblk;
  stats_reduce_kernel<<<(tri(h->D + 1) + 127) / 128, 128, 0, h->stream>>>(h->ws, nblk, h->D, (double)e.nlocal, e.stats);
  h->launches++;
  CU(cudaGetLastError());
  return BDF_OK;
}

}  // namespace
int bdf_sample_entity_impl(bdf_t* h, int entity, const double* mu_dev, int64_t mu_ld, const double* Lambda_dev, const double* Z_dev) {
  return sample_entity(h, entity, mu_dev, mu_ld, Lambda_dev, Z_dev);
}
namespace {

int draw_entity(bdf_t* h, int entity, const double* mu0_dev, double b0, const double* Tinv_dev, double nu, const double* A_dev,
                const double* z_dev, cudaStream_t side = nullptr) {
  EntityS& e = h->ents[entity];
  NWDrawParams p{};
  p.D = h->D; p.stats = e.stats; p.mu0 = mu0_dev; p.Tinv = Tinv_dev; p.b0 = b0; p.nu = nu; p.A_inj = A_dev; p.z_inj = z_dev;
  p.seed = h->seed; p.sweep = h->sweep; p.stream = philox_stream(PHILOX_NW, 4u * (uint32_t)entity);  // the draw uses streams +0 … +3
  p.scratch = h->scratch;
  p.mu_out = e.mu; p.Lam_out = e.Lambda; p.err_flag = h->err_flag;
#ifdef BDF_DEBUG
  p.debug = getenv("BDF_DEBUG_NW") != nullptr;
#endif
  const size_t dd8 = sizeof(double) * (size_t)h->D * h->D;
  p.nsm = 2 * dd8 <= 227 * 1024 ? 2 : (dd8 <= 227 * 1024 ? 1 : 0);
  // on a side stream the draw runs next to the following half-sweep's row kernel: the no-shared-memory, <= 64-register variant fits
  // into the slot one retiring row CTA leaves behind (the row kernel owns all shared memory and registers of an SM otherwise)
  if (side) p.nsm = 0;
  cudaStream_t st = side ? side : h->stream;
  if (!(h->smem_optin & BDF_OPTIN_NWDRAW)) {
    CU(cudaFuncSetAttribute(nw_draw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU(cudaFuncSetAttribute(nw_draw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->smem_optin |= BDF_OPTIN_NWDRAW;
  }
  if (p.nsm == 2) nw_draw_kernel<2><<<1, 256, 2 * dd8, st>>>(p);
  else if (p.nsm == 1) nw_draw_kernel<1><<<1, 256, dd8, st>>>(p);
  else nw_draw_kernel<0><<<1, 256, 0, st>>>(p);
  h->launches++;
  CU(cudaGetLastError());
  return BDF_OK;
}

}  // namespace

// =================================================================================================================
extern "C" {

int bdf_version(void) { return 100; }

const char* bdf_last_error(const bdf_t* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int bdf_create(bdf_t** out, int device, int num_latent, int rank, int world) {
  if (!out) return BDF_ERR_INVALID;
  *out = nullptr;
  if (num_latent < 1 || num_latent > 128) { g_create_err = "num_latent must be in 1..128"; return BDF_ERR_INVALID; }
  if (world < 1 || rank < 0 || rank >= world) { g_create_err = "need 0 <= rank < world"; return BDF_ERR_INVALID; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) { g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(ce); return BDF_ERR_CUDA; }
  if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return BDF_ERR_INVALID; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { g_create_err = "libbdf_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor); return BDF_ERR_CUDA; }
  bdf_t* h = new bdf_t();
  h->device = device; h->D = num_latent; h->rank = rank; h->world = world;
  h->ld = round_up(num_latent, 4);
  h->DP = round_up(num_latent, 8);
  h->NW = h->DP <= 32 ? 1 : (h->DP <= 64 ? 4 : 8);
  h->pst = pst_of(h->DP);
  auto bail = [&](cudaError_t e, const char* what) { g_create_err = std::string(what) + ": " + cudaGetErrorString(e); delete h; return BDF_ERR_CUDA; };
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail(ce, "cudaSetDevice");
  if ((ce = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "cudaStreamCreate");
  h->stream = h->own_stream;
  if ((ce = cudaMalloc((void**)&h->err_flag, sizeof(int))) != cudaSuccess) return bail(ce, "cudaMalloc");
  {
    if ((ce = cudaMalloc((void**)&h->ones, sizeof(double) * h->ld)) != cudaSuccess) return bail(ce, "cudaMalloc");
    std::vector<double> one(h->ld, 0.0);
    for (int j = 0; j < h->D; j++) one[j] = 1.0;
    cudaMemcpy(h->ones, one.data(), sizeof(double) * h->ld, cudaMemcpyHostToDevice);
  }
  h->num_sms = prop.multiProcessorCount;
  if ((ce = cudaMalloc((void**)&h->work_counter, sizeof(int))) != cudaSuccess) return bail(ce, "cudaMalloc");
  // the persistent warp-specialised row kernel (row_kernel_ws.cuh) is opt-in: measured slower than one CTA per row on C2 (profiles/README.md)
  { const char* ws = getenv("BDF_ROWS_WS"); h->use_ws = ws && ws[0] == '1'; }
  cudaMemset(h->err_flag, 0, sizeof(int));
  if ((ce = cudaMalloc((void**)&h->scratch, sizeof(double) * ((size_t)4 * num_latent * num_latent + 4 * num_latent))) != cudaSuccess) return bail(ce, "cudaMalloc");
  if ((ce = cudaMalloc((void**)&h->lt, sizeof(double) * (64 * (size_t)(h->DP / 8) * (h->DP / 8 + 1) / 2 + h->DP))) != cudaSuccess) return bail(ce, "cudaMalloc");
  h->ws_bytes = sizeof(double) * 296 * (size_t)tri(num_latent + 1);
  if ((ce = cudaMalloc((void**)&h->ws, h->ws_bytes)) != cudaSuccess) return bail(ce, "cudaMalloc");
  *out = h;
  return BDF_OK;
}

int bdf_destroy(bdf_t* h) {
  if (!h) return BDF_OK;
  cudaSetDevice(h->device);
  if (h->side) cudaStreamSynchronize(h->side);
  cudaStreamSynchronize(h->stream);
  bdf_dense_teardown(h);
  for (auto& e : h->ents) {
    for (int r = 0; r < 8; r++) if (e.peerU[r]) cudaIpcCloseMemHandle(e.peerU[r]);
    for (int r = 0; r < 8; r++) if (e.peer_beta[r]) cudaIpcCloseMemHandle(e.peer_beta[r]);
    cudaFree(e.slot_of_row); cudaFree(e.row_of_slot);
    cudaFree(e.inj); if (e.pinned) cudaFreeHost(e.pinned); if (e.ev_done) cudaEventDestroy(e.ev_done);
    cudaFree(e.U); cudaFree(e.mu); cudaFree(e.Lambda); cudaFree(e.mu_rows); cudaFree(e.Z); cudaFree(e.stats); cudaFree(e.hyper);
    cudaFree(e.f_rowptr); cudaFree(e.f_colind); cudaFree(e.f_colptr); cudaFree(e.f_rowind); cudaFree(e.beta); cudaFree(e.uhat); cudaFree(e.cgbuf); cudaFree(e.btb);
    cudaFree(e.sp_items[0]); cudaFree(e.sp_items[1]); cudaFree(e.sp_long[0]); cudaFree(e.sp_long[1]); cudaFree(e.sp_part);
    cudaFree(e.f_val_csr); cudaFree(e.f_val_csc); cudaFree(e.f_dense); cudaFree(e.FF);
    free_work_list(e.merged);
  }
  for (auto& r : h->rels) {
    cudaFree(r.F); cudaFree(r.FF); cudaFree(r.beta); cudaFree(r.linear); cudaFree(r.res);
    bdf_free_test(r);
  }
  for (auto& r : h->rels)
    for (int m = 0; m < r.K; m++) {
      ModeIndex& mi = r.modes[m];
      cudaFree(mi.row_ptr); cudaFree(mi.col[0]); cudaFree(mi.col[1]); cudaFree(mi.val);
      free_work_list(mi);
      cudaFree(mi.perm); cudaFree(mi.val_adj);
    }
  cudaFree(h->ws); cudaFree(h->scratch); cudaFree(h->err_flag); cudaFree(h->lt); cudaFree(h->ones); cudaFree(h->arena); cudaFree(h->arena2); cudaFree(h->work_counter);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return BDF_OK;
}

int bdf_set_stream(bdf_t* h, void* s) {
  CHECK_H();
  CU(cudaStreamSynchronize(h->stream));
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return BDF_OK;
}

int bdf_set_seed(bdf_t* h, uint64_t seed) { CHECK_H(); h->seed = seed; return BDF_OK; }
int64_t bdf_sweep_counter(const bdf_t* h) { return h ? (int64_t)h->sweep : -1; }
int64_t bdf_launch_count(const bdf_t* h) { return h ? h->launches : -1; }
int bdf_synchronize(bdf_t* h) { CHECK_H(); CU(cudaSetDevice(h->device)); CU(cudaStreamSynchronize(h->stream)); return BDF_OK; }

static int add_entity_impl(bdf_t* h, int64_t count, int64_t nper, int64_t nlocal, const std::vector<int32_t>* slot_of_row) {
  EntityS e;
  e.N = count;
  e.Nper = nper;
  e.nlocal = nlocal;
  const int D = h->D;
  const size_t un = (size_t)e.Nper * h->world * h->ld;
  int rc;
  if ((rc = dev_alloc(h, &e.U, un))) return rc;
  if ((rc = dev_alloc(h, &e.mu, (size_t)D))) return rc;
  if ((rc = dev_alloc(h, &e.Lambda, (size_t)D * D))) return rc;
  if ((rc = dev_alloc(h, &e.stats, (size_t)1 + D + (size_t)D * D))) return rc;
  if ((rc = dev_alloc(h, &e.hyper, (size_t)D + (size_t)D * D))) return rc;
  if (slot_of_row) {
    std::vector<int32_t> inv((size_t)e.Nper * h->world, -1);
    for (int64_t i = 0; i < count; i++) inv[(*slot_of_row)[i]] = (int32_t)i;
    if ((rc = dev_alloc(h, &e.slot_of_row, (size_t)count))) return rc;
    if ((rc = dev_alloc(h, &e.row_of_slot, inv.size()))) return rc;
    CU(cudaMemcpy(e.slot_of_row, slot_of_row->data(), sizeof(int32_t) * count, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e.row_of_slot, inv.data(), sizeof(int32_t) * inv.size(), cudaMemcpyHostToDevice));
  }
  // initModel! — src/RelationData.jl:66-90
  CU(cudaMemsetAsync(e.U, 0, un * sizeof(double), h->stream));
  CU(cudaMemsetAsync(e.mu, 0, D * sizeof(double), h->stream));
  set_identity_kernel<<<grid_for(D * D), 256, 0, h->stream>>>(e.Lambda, D, 5.0);
  CU(cudaMemsetAsync(e.hyper, 0, D * sizeof(double), h->stream));
  set_identity_kernel<<<grid_for(D * D), 256, 0, h->stream>>>(e.hyper + D, D, 1.0);
  CU(cudaGetLastError());
  e.mu0.assign(D, 0.0);
  e.WI.assign((size_t)D * D, 0.0);
  for (int i = 0; i < D; i++) e.WI[i + (size_t)i * D] = 1.0;
  e.b0 = 2.0;
  e.nu0 = D;
  h->ents.push_back(e);
  return (int)h->ents.size() - 1;
}

int bdf_add_entity(bdf_t* h, int64_t count) {
  CHECK_H();
  if (count < 1 || count > 2000000000LL) FAIL(BDF_ERR_INVALID, "entity count must be in 1..2e9");
  CU(cudaSetDevice(h->device));
  int64_t nlocal = (count - h->rank + h->world - 1) / h->world;
  if (nlocal < 0) nlocal = 0;
  return add_entity_impl(h, count, (count + h->world - 1) / h->world, nlocal, nullptr);
}

int bdf_add_entity_partitioned(bdf_t* h, int64_t count, const int32_t* rank_of_row) {
  CHECK_H();
  if (count < 1 || count > 2000000000LL) FAIL(BDF_ERR_INVALID, "entity count must be in 1..2e9");
  if (!rank_of_row) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  std::vector<int64_t> cnt(h->world, 0);
  for (int64_t i = 0; i < count; i++) {
    if (rank_of_row[i] < 0 || rank_of_row[i] >= h->world) FAIL(BDF_ERR_INVALID, "rank_of_row entries must lie in 0..world-1");
    cnt[rank_of_row[i]]++;
  }
  int64_t nper = 1;
  for (int r = 0; r < h->world; r++) nper = std::max(nper, cnt[r]);
  if (nper * h->world >= 2147483647LL) FAIL(BDF_ERR_INVALID, "too many slots");
  std::vector<int32_t> slot((size_t)count);
  std::vector<int64_t> next(h->world, 0);
  for (int64_t i = 0; i < count; i++) {  // rows keep their relative order inside a shard
    const int r = rank_of_row[i];
    slot[i] = (int32_t)(r * nper + next[r]++);
  }
  return add_entity_impl(h, count, nper, cnt[h->rank], &slot);
}

int bdf_add_relation(bdf_t* h, int K, const int* entity_of_mode, int64_t nnz, const int64_t* ids, const double* vals) {
  CHECK_H();
  if (K < 2 || K > 3) FAIL(BDF_ERR_INVALID, "relations with 2 or 3 modes are supported");
  if (nnz < 0 || nnz >= 2147483647LL) FAIL(BDF_ERR_INVALID, "nnz must be < 2^31");
  if (!entity_of_mode || (nnz > 0 && (!ids || !vals))) FAIL(BDF_ERR_INVALID, "null argument");
  for (int m = 0; m < K; m++) {
    CHECK_ENT(entity_of_mode[m]);
    for (int m2 = 0; m2 < m; m2++)
      if (entity_of_mode[m] == entity_of_mode[m2]) FAIL(BDF_ERR_INVALID, "an entity may appear once per relation");
  }
  CU(cudaSetDevice(h->device));
  RelationS rel;
  rel.K = K; rel.nnz = nnz;
  for (int m = 0; m < K; m++) rel.entity_of_mode[m] = entity_of_mode[m];
  // staging copies of the table
  int64_t* d_ids = nullptr; double* d_vals = nullptr; uint32_t *keys = nullptr, *keys2 = nullptr, *idx = nullptr, *idx2 = nullptr;
  int* d_bad = nullptr; int64_t* d_lb = nullptr; unsigned long long* d_cnt = nullptr; void* d_tmp = nullptr;
  int rc = BDF_OK;
  auto cleanup = [&]() { cudaFree(d_ids); cudaFree(d_vals); cudaFree(keys); cudaFree(keys2); cudaFree(idx); cudaFree(idx2); cudaFree(d_bad); cudaFree(d_lb); cudaFree(d_cnt); cudaFree(d_tmp); };
#define TRY(x) do { rc = (x); if (rc) { cleanup(); return rc; } } while (0)
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return BDF_ERR_CUDA; } } while (0)
  const size_t n1 = std::max<int64_t>(nnz, 1);
  TRY(dev_alloc(h, &d_ids, n1 * K)); TRY(dev_alloc(h, &d_vals, n1));
  TRY(dev_alloc(h, &keys, n1)); TRY(dev_alloc(h, &keys2, n1)); TRY(dev_alloc(h, &idx, n1)); TRY(dev_alloc(h, &idx2, n1));
  TRY(dev_alloc(h, &d_bad, 1)); TRY(dev_alloc(h, &d_lb, 2));
  CUT(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  if (nnz) {
    CUT(cudaMemcpyAsync(d_ids, ids, sizeof(int64_t) * nnz * K, cudaMemcpyHostToDevice, h->stream));
    CUT(cudaMemcpyAsync(d_vals, vals, sizeof(double) * nnz, cudaMemcpyHostToDevice, h->stream));
  }
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)n1, 0, 32, h->stream);
  size_t scan_bytes = 0;
  int64_t max_rows = 0;
  for (int m = 0; m < K; m++) max_rows = std::max(max_rows, h->ents[entity_of_mode[m]].Nper);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned long long*)nullptr, (int64_t*)nullptr, (int)(max_rows + 1), h->stream);
  tmp_bytes = std::max(tmp_bytes, scan_bytes);
  CUT(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
  TRY(dev_alloc(h, &d_cnt, (size_t)max_rows + 1));

  for (int m = 0; m < K; m++) {
    EntityS& e = h->ents[entity_of_mode[m]];
    ModeIndex& mi = rel.modes[m];
    mi.nrows = e.Nper;
    const int64_t slot0 = (int64_t)h->rank * e.Nper;
    int bits = 1;
    while ((1LL << bits) < e.Nper * h->world) bits++;
    make_keys_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(d_ids + (size_t)m * nnz, nnz, e.N, h->world, e.Nper, e.slot_of_row, keys, idx, d_bad);
    CUT(cudaGetLastError());
    if (nnz) CUT(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys, keys2, idx, idx2, (int)nnz, 0, bits, h->stream));  // stable: table order kept
    // local segment of the sorted table
    lower_bound_kernel<<<1, 1, 0, h->stream>>>(keys2, nnz, (uint32_t)slot0, d_lb);
    lower_bound_kernel<<<1, 1, 0, h->stream>>>(keys2, nnz, (uint32_t)(slot0 + e.Nper), d_lb + 1);
    CUT(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * (e.Nper + 1), h->stream));
    count_rows_kernel<<<grid_for(nnz), 256, 0, h->stream>>>(keys2, nnz, slot0, e.Nper, d_cnt);
    CUT(cudaGetLastError());
    TRY(dev_alloc(h, &mi.row_ptr, (size_t)e.Nper + 1));
    CUT(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, mi.row_ptr, (int)(e.Nper + 1), h->stream));
    int64_t lb[2] = {0, 0};
    int bad = 0;
    CUT(cudaMemcpyAsync(lb, d_lb, sizeof(lb), cudaMemcpyDeviceToHost, h->stream));
    CUT(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUT(cudaStreamSynchronize(h->stream));
    if (bad) { h->err = "relation id outside 1..count of its entity"; cleanup(); return BDF_ERR_INVALID; }
    mi.nnz = lb[1] - lb[0];
    int no = 0;
    for (int m2 = 0; m2 < K; m2++) {
      if (m2 == m) continue;
      EntityS& eo = h->ents[entity_of_mode[m2]];
      mi.other_entity[no] = entity_of_mode[m2];
      TRY(dev_alloc(h, &mi.col[no], (size_t)std::max<int64_t>(mi.nnz, 1)));
      gather_obs_kernel<<<grid_for(mi.nnz), 256, 0, h->stream>>>(idx2, lb[0], mi.nnz, d_ids + (size_t)m2 * nnz, h->world, eo.Nper, eo.slot_of_row, mi.col[no]);
      CUT(cudaGetLastError());
      no++;
    }
    TRY(dev_alloc(h, &mi.val, (size_t)std::max<int64_t>(mi.nnz, 1)));
    gather_val_kernel<<<grid_for(mi.nnz), 256, 0, h->stream>>>(idx2, lb[0], mi.nnz, d_vals, mi.val);
    TRY(dev_alloc(h, &mi.perm, (size_t)std::max<int64_t>(mi.nnz, 1)));
    if (mi.nnz) CUT(cudaMemcpyAsync(mi.perm, idx2 + lb[0], sizeof(uint32_t) * mi.nnz, cudaMemcpyDeviceToDevice, h->stream));
    CUT(cudaGetLastError());
    std::vector<int64_t> rp((size_t)e.Nper + 1);
    CUT(cudaMemcpyAsync(rp.data(), mi.row_ptr, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost, h->stream));
    CUT(cudaStreamSynchronize(h->stream));
    mi.h_row_ptr = rp;
    TRY(build_work_list(h, mi, {&mi.h_row_ptr}, e.nlocal));
  }
  CUT(cudaStreamSynchronize(h->stream));
  cleanup();
#undef TRY
#undef CUT
  h->rels.push_back(rel);
  const int rid = (int)h->rels.size() - 1;
  for (int m = 0; m < K; m++) h->ents[entity_of_mode[m]].uses.push_back({rid, m});
  return rid;
}

int bdf_set_relation_params(bdf_t* h, int rel, double alpha, double mean_value) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  if (!(alpha > 0.0)) FAIL(BDF_ERR_INVALID, "alpha must be positive");
  h->rels[rel].alpha = alpha;
  h->rels[rel].mean = mean_value;
  if (h->rels[rel].F) return bdf_refresh_relation_offsets(h, rel);  // linear_values = mean_value + F·beta moves with the mean
  return BDF_OK;
}

int bdf_set_factors(bdf_t* h, int entity, const double* U) {
  CHECK_H(); CHECK_ENT(entity);
  if (!U) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  CU(cudaMemsetAsync(e.U, 0, sizeof(double) * (size_t)e.Nper * h->world * h->ld, h->stream));
  int rc = copy_rows_h2d(h, e, U, e.U);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

int bdf_get_factors(bdf_t* h, int entity, double* U) {
  CHECK_H(); CHECK_ENT(entity);
  if (!U) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  int rc = copy_rows_d2h(h, h->ents[entity], h->ents[entity].U, U);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

int bdf_factors_dev(bdf_t* h, int entity, void** dev_ptr, int64_t* nper, int64_t* ld) {
  CHECK_H(); CHECK_ENT(entity);
  if (dev_ptr) *dev_ptr = h->ents[entity].U;
  if (nper) *nper = h->ents[entity].Nper;
  if (ld) *ld = h->ld;
  return BDF_OK;
}

int bdf_ipc_export(bdf_t* h, int entity, unsigned char* handle64) {
  CHECK_H(); CHECK_ENT(entity);
  if (!handle64) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  CU(cudaIpcGetMemHandle(&mh, h->ents[entity].U));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &mh, 64);
  return BDF_OK;
}

int bdf_ipc_import(bdf_t* h, int entity, int peer_rank, const unsigned char* handle64) {
  CHECK_H(); CHECK_ENT(entity);
  if (!handle64 || peer_rank < 0 || peer_rank >= h->world || peer_rank >= 8 || peer_rank == h->rank) FAIL(BDF_ERR_INVALID, "bad peer rank (fused all-gather supports up to 8 ranks)");
  CU(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void* ptr = nullptr;
  CU(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
  h->ents[entity].peerU[peer_rank] = (double*)ptr;
  return BDF_OK;
}

int bdf_stats_dev(bdf_t* h, int entity, void** dev_ptr, int64_t* count) {
  CHECK_H(); CHECK_ENT(entity);
  if (dev_ptr) *dev_ptr = h->ents[entity].stats;
  if (count) *count = 1 + h->D + (int64_t)h->D * h->D;
  return BDF_OK;
}

int bdf_sample_mode(bdf_t* h, int entity, const double* mu, int64_t mu_ld, const double* Lambda, const double* z) {
  CHECK_H(); CHECK_ENT(entity);
  if (!mu || !Lambda) FAIL(BDF_ERR_INVALID, "null argument");
  if (mu_ld != 0 && mu_ld != h->D) FAIL(BDF_ERR_INVALID, "mu_ld must be 0 (vector) or num_latent (D×N matrix)");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  const int D = h->D;
  const size_t un = (size_t)e.Nper * h->world * h->ld;
  int rc;
  if ((rc = join_draw(h, e))) return rc;
  CU(cudaMemcpyAsync(e.Lambda, Lambda, sizeof(double) * D * D, cudaMemcpyHostToDevice, h->stream));
  const double* mu_dev = e.mu;
  int64_t mu_pitch = 0;
  if (mu_ld == 0) {
    CU(cudaMemcpyAsync(e.mu, mu, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  } else {
    if (!e.mu_rows) { if ((rc = dev_alloc(h, &e.mu_rows, un))) return rc; CU(cudaMemsetAsync(e.mu_rows, 0, un * sizeof(double), h->stream)); }
    if ((rc = copy_rows_h2d(h, e, mu, e.mu_rows))) return rc;
    mu_dev = e.mu_rows;
    mu_pitch = h->ld;
  }
  const double* z_dev = nullptr;
  if (z) {
    if (!e.Z) { if ((rc = dev_alloc(h, &e.Z, un))) return rc; CU(cudaMemsetAsync(e.Z, 0, un * sizeof(double), h->stream)); }
    if ((rc = copy_rows_h2d(h, e, z, e.Z))) return rc;
    z_dev = e.Z;
  }
  if ((rc = sample_entity(h, entity, mu_dev, mu_pitch, e.Lambda, z_dev))) return rc;
  if (h->async_mode && !z) return BDF_OK;  // deferred: the flag is read by the next synchronising call (injected noise lives in the arena: drain)
  return check_err_flag(h);
}

int bdf_nw_stats(bdf_t* h, int entity, double* N, double* NU, double* NS) {
  CHECK_H(); CHECK_ENT(entity);
  CU(cudaSetDevice(h->device));
  int rc = stats_entity(h, entity);
  if (rc) return rc;
  const int D = h->D;
  std::vector<double> buf((size_t)1 + D + (size_t)D * D);
  CU(cudaMemcpyAsync(buf.data(), h->ents[entity].stats, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (N) *N = buf[0];
  if (NU) memcpy(NU, buf.data() + 1, sizeof(double) * D);
  if (NS) memcpy(NS, buf.data() + 1 + D, sizeof(double) * D * D);
  return BDF_OK;
}

int bdf_set_nw_stats(bdf_t* h, int entity, double N, const double* NU, const double* NS) {
  CHECK_H(); CHECK_ENT(entity);
  if (!NU || !NS) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  const int D = h->D;
  std::vector<double> buf((size_t)1 + D + (size_t)D * D);
  buf[0] = N;
  memcpy(buf.data() + 1, NU, sizeof(double) * D);
  memcpy(buf.data() + 1 + D, NS, sizeof(double) * D * D);
  CU(cudaMemcpyAsync(h->ents[entity].stats, buf.data(), sizeof(double) * buf.size(), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

int bdf_nw_sample(bdf_t* h, int entity, const double* mu0, double b0, const double* Tinv, double nu, const double* bartlettA,
                  const double* z, double* mu_out, double* Lambda_out) {
  CHECK_H(); CHECK_ENT(entity);
  if (!mu0 || !Tinv) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  const int D = h->D;
  const size_t dd = (size_t)D * D;
  { int rcj = bdf_join_side(h); if (rcj) return rcj; }
  CU(cudaMemcpyAsync(e.hyper, mu0, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(e.hyper + D, Tinv, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream));
  double* A_dev = nullptr; double* z_dev = nullptr; double* stage = nullptr;
  if (bartlettA || z) {
    int rca = bdf_ensure_arena(h, sizeof(double) * (dd + D));  // injected variates are staged in the handle's grow-only arena
    if (rca) return rca;
    stage = reinterpret_cast<double*>(h->arena);
    if (bartlettA) { A_dev = stage; CU(cudaMemcpyAsync(A_dev, bartlettA, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream)); }
    if (z) { z_dev = stage + dd; CU(cudaMemcpyAsync(z_dev, z, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream)); }
  }
  int rc = draw_entity(h, entity, e.hyper, b0, e.hyper + D, nu, A_dev, z_dev);
  if (!rc && mu_out) { cudaError_t ce = cudaMemcpyAsync(mu_out, e.mu, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream); if (ce != cudaSuccess) rc = BDF_ERR_CUDA; }
  if (!rc && Lambda_out) { cudaError_t ce = cudaMemcpyAsync(Lambda_out, e.Lambda, sizeof(double) * dd, cudaMemcpyDeviceToHost, h->stream); if (ce != cudaSuccess) rc = BDF_ERR_CUDA; }
  if (!rc) rc = check_err_flag(h); else cudaStreamSynchronize(h->stream);
  return rc;
}

int bdf_set_async(bdf_t* h, int on) { CHECK_H(); h->async_mode = on != 0; return BDF_OK; }

int bdf_nw_sample_async(bdf_t* h, int entity, const double* mu0, double b0, const double* Tinv, double nu, const double* bartlettA, const double* z) {
  CHECK_H(); CHECK_ENT(entity);
  if (!mu0 || !Tinv) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  const int D = h->D;
  const size_t dd = (size_t)D * D;
  if (!h->side) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  }
  if (!e.pinned) {
    CU(cudaMallocHost((void**)&e.pinned, sizeof(double) * (dd + D + 1)));
    CU(cudaMalloc((void**)&e.inj, sizeof(double) * (dd + D)));
    CU(cudaEventCreateWithFlags(&e.ev_done, cudaEventDisableTiming));
  }
  int rc;
  if ((rc = join_draw(h, e))) return rc;
  CU(cudaMemcpyAsync(e.hyper, mu0, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(e.hyper + D, Tinv, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream));
  double* A_dev = nullptr; double* z_dev = nullptr;
  if (bartlettA) { A_dev = e.inj; CU(cudaMemcpyAsync(A_dev, bartlettA, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream)); }
  if (z) { z_dev = e.inj + dd; CU(cudaMemcpyAsync(z_dev, z, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream)); }
  CU(cudaEventRecord(h->ev_ready, h->stream));      // statistics, hyper-priors and variates are in place
  CU(cudaStreamWaitEvent(h->side, h->ev_ready, 0));
  if ((rc = draw_entity(h, entity, e.hyper, b0, e.hyper + D, nu, A_dev, z_dev, h->side))) return rc;
  CU(cudaMemcpyAsync(e.pinned, e.mu, sizeof(double) * D, cudaMemcpyDeviceToHost, h->side));
  CU(cudaMemcpyAsync(e.pinned + D, e.Lambda, sizeof(double) * dd, cudaMemcpyDeviceToHost, h->side));
  CU(cudaMemcpyAsync(e.pinned + D + dd, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->side));
  CU(cudaEventRecord(e.ev_done, h->side));
  e.draw_pending = true;
  return BDF_OK;
}

int bdf_nw_sample_fetch(bdf_t* h, int entity, double* mu_out, double* Lambda_out) {
  CHECK_H(); CHECK_ENT(entity);
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (!e.ev_done) FAIL(BDF_ERR_STATE, "no bdf_nw_sample_async call to fetch");
  const int D = h->D;
  const size_t dd = (size_t)D * D;
  CU(cudaEventSynchronize(e.ev_done));
  int rc = join_draw(h, e);
  if (rc) return rc;
  if (mu_out) memcpy(mu_out, e.pinned, sizeof(double) * D);
  if (Lambda_out) memcpy(Lambda_out, e.pinned + D, sizeof(double) * dd);
  int flag = 0;
  memcpy(&flag, e.pinned + D + dd, sizeof(int));
  if (flag) {
    CU(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
    FAIL(BDF_ERR_NUMERIC, flag & 1 ? "row draw: precision matrix not positive definite" : "Normal-Wishart draw: matrix not positive definite");
  }
  return BDF_OK;
}

int bdf_step_sample(bdf_t* h, int entity) {
  CHECK_H(); CHECK_ENT(entity);
  EntityS& e = h->ents[entity];
  return sample_entity(h, entity, e.mu, 0, e.Lambda, nullptr);
}
int bdf_step_nw_stats(bdf_t* h, int entity) { CHECK_H(); CHECK_ENT(entity); return stats_entity(h, entity); }
int bdf_step_nw_draw(bdf_t* h, int entity) {
  CHECK_H(); CHECK_ENT(entity);
  EntityS& e = h->ents[entity];
  return draw_entity(h, entity, e.hyper, e.b0, e.hyper + h->D, e.nu0, nullptr, nullptr);
}

int bdf_step_nw_draw_on(bdf_t* h, int entity, void* cuda_stream) {
  CHECK_H(); CHECK_ENT(entity);
  if (!cuda_stream) FAIL(BDF_ERR_INVALID, "null stream");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  return draw_entity(h, entity, e.hyper, e.b0, e.hyper + h->D, e.nu0, nullptr, nullptr, (cudaStream_t)cuda_stream);
}

int bdf_sweep(bdf_t* h, int nsweeps) {
  CHECK_H();
  if (h->world != 1) FAIL(BDF_ERR_STATE, "bdf_sweep drives one GPU; with world > 1 use the bdf_step_* entries around the collectives");
  CU(cudaSetDevice(h->device));
  for (int s = 0; s < nsweeps; s++) {
    for (int e = 0; e < (int)h->ents.size(); e++) {  // Gauss-Seidel order of src/macau.jl:96-134
      int rc;
      if ((rc = bdf_step_sample(h, e))) return rc;
      if ((rc = bdf_step_nw_stats(h, e))) return rc;
      if ((rc = bdf_step_nw_draw(h, e))) return rc;
    }
    h->sweep++;
  }
  return BDF_OK;
}

int bdf_advance_sweep(bdf_t* h) { CHECK_H(); h->sweep++; return BDF_OK; }

int bdf_get_hyper(bdf_t* h, int entity, double* mu, double* Lambda) {
  CHECK_H(); CHECK_ENT(entity);
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (mu) CU(cudaMemcpyAsync(mu, e.mu, sizeof(double) * h->D, cudaMemcpyDeviceToHost, h->stream));
  if (Lambda) CU(cudaMemcpyAsync(Lambda, e.Lambda, sizeof(double) * h->D * h->D, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_err_flag(h);
}

int bdf_set_hyper(bdf_t* h, int entity, const double* mu, const double* Lambda) {
  CHECK_H(); CHECK_ENT(entity);
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (mu) CU(cudaMemcpyAsync(e.mu, mu, sizeof(double) * h->D, cudaMemcpyHostToDevice, h->stream));
  if (Lambda) CU(cudaMemcpyAsync(e.Lambda, Lambda, sizeof(double) * h->D * h->D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

int bdf_debug_row_noise(bdf_t* h, int entity, uint64_t sweep, double* z) {
  CHECK_H(); CHECK_ENT(entity);
  if (!z) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, sizeof(double) * (size_t)e.N * h->D));
  row_noise_kernel<<<grid_for(e.N * h->D), 256, 0, h->stream>>>(d, h->D, e.N, h->seed, sweep, philox_stream(PHILOX_ROW, (uint32_t)entity));
  cudaError_t ce = cudaMemcpyAsync(z, d, sizeof(double) * (size_t)e.N * h->D, cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (ce != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce));
  return BDF_OK;
}

int bdf_debug_phase_clocks(bdf_t* h, int entity, double* mean_cycles /* 7 */, int64_t* n_items) {
  CHECK_H(); CHECK_ENT(entity);
#ifndef BDF_DEBUG
  (void)mean_cycles; (void)n_items;
  FAIL(BDF_ERR_STATE, "phase clocks are compiled out of the release library (build with BDF_EXTRA_NVCC=-DBDF_DEBUG)");
#else
  CU(cudaSetDevice(h->device));
  EntityS& e = h->ents[entity];
  if (e.uses.empty()) FAIL(BDF_ERR_STATE, "entity takes part in no relation");
  const int ni = e.uses.size() > 1 && e.merged_uses == e.uses.size() ? e.merged.n_items : h->rels[e.uses[0].first].modes[e.uses[0].second].n_items;
  long long* d = nullptr;
  const size_t nd = 16 * (size_t)std::max(ni, 1000);
  CU(cudaMalloc((void**)&d, sizeof(long long) * nd));
  CU(cudaMemsetAsync(d, 0, sizeof(long long) * nd, h->stream));
  int rc = sample_entity(h, entity, e.mu, 0, e.Lambda, nullptr, d);
  std::vector<long long> hbuf(nd);
  cudaMemcpyAsync(hbuf.data(), d, sizeof(long long) * hbuf.size(), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (rc) return rc;
  if (getenv("BDF_DEBUG_WS")) {  // persistent warp-specialised kernel: per-group busy / wait cycles, averaged over the CTAs
    double a[5][5] = {{0}};
    const int nb = std::min(h->num_sms, (ni + 1) / 2);
    for (int b = 0; b < nb; b++)
      for (int g = 0; g < 5; g++)
        for (int k = 0; k < 5; k++) a[g][k] += (double)hbuf[((size_t)b * 5 + g) * 8 + k] / nb;
    for (int g = 0; g < 2; g++)
      fprintf(stderr, "  syrk group %d: total %.0f  syrk+split %.0f  wait-for-slot %.0f  park %.0f  rows parked %.0f\n", g, a[g][0], a[g][1], a[g][2], a[g][3], a[g][4]);
    {
      double t[5] = {0};
      for (int b = 0; b < nb; b++)
        for (int g = 0; g < 2; g++)
          for (int k = 0; k < 5; k++) t[k] += (double)hbuf[((size_t)nb * 5 + (size_t)b * 2 + g) * 8 + k] / (2 * nb);
      fprintf(stderr, "  syrk_item of warp 0, per group: prologue %.0f  wait-for-stage %.0f  group barrier %.0f  issue+meta %.0f  compute %.0f\n", t[0], t[1], t[2], t[3], t[4]);
    }
    for (int g = 2; g < 5; g++)
      fprintf(stderr, "  finalise group %d: total %.0f  wait-for-row %.0f  factor+draw %.0f  rows %.0f\n", g - 2, a[g][0], a[g][1], a[g][2], a[g][4]);
    for (int k = 0; k < 7; k++) mean_cycles[k] = 0.0;
    if (n_items) *n_items = ni;
    return BDF_OK;
  }
  for (int k = 0; k < 7; k++) mean_cycles[k] = 0.0;
  int64_t cnt = 0;
  for (int i = 0; i < ni; i++) {
    const long long* t = &hbuf[(size_t)i * 8];
    if (t[6] == 0) continue;  // split-row chunk that did not finalise
    for (int k = 1; k < 7; k++) mean_cycles[k - 1] += (double)(t[k] - t[k - 1]);
    mean_cycles[6] += (double)(t[6] - t[0]);
    cnt++;
  }
  for (int k = 0; k < 7; k++) mean_cycles[k] /= (double)std::max<int64_t>(cnt, 1);
  if (getenv("BDF_DEBUG_PANEL")) {
    double acc[8] = {0};
    for (int i = 0; i < ni; i++) for (int k = 0; k < 8; k++) acc[k] += (double)hbuf[(size_t)ni * 8 + (size_t)i * 8 + k];
    fprintf(stderr, "panel loop mean cycles per row: warp0 [scale %.0f, wait1 %.0f, trailing+diag %.0f, wait2 %.0f]  warp1 [scale %.0f, wait1 %.0f, backsub+trailing %.0f, wait2 %.0f]\n",
            acc[0] / ni, acc[1] / ni, acc[2] / ni, acc[3] / ni, acc[4] / ni, acc[5] / ni, acc[6] / ni, acc[7] / ni);
  }
  if (n_items) *n_items = cnt;
  return BDF_OK;
#endif
}

// yhat[t] += Σ_f F[t, f]·beta[f] — the `F * r.model.beta` term of pred(r, probe_vec, F), src/sampling.jl:13
__global__ void add_fbeta_kernel(const double* __restrict__ F, const double* __restrict__ beta, int64_t n, int64_t nF, double* __restrict__ y) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t f = 0; f < nF; f++) s = fma(F[t + f * n], beta[f], s);
    y[t] += s;
  }
}

int bdf_predict(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, double* yhat) { return bdf_predict_f(h, rel, ntest, ids, nullptr, yhat); }

int bdf_predict_f(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, const double* test_F, double* yhat) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  if (ntest < 0 || (ntest > 0 && (!ids || !yhat))) FAIL(BDF_ERR_INVALID, "null argument");
  if (ntest == 0) return BDF_OK;
  CU(cudaSetDevice(h->device));
  RelationS& r = h->rels[rel];
  if (r.F && !test_F) FAIL(BDF_ERR_INVALID, "the relation has features: pass the feature rows of the test observations (bdf_predict_f)");
  if (!r.F && test_F) FAIL(BDF_ERR_INVALID, "the relation has no features");
  // staging lives in a grow-only arena of the handle: no cudaMalloc/cudaFree (both synchronise the device) per call
  const size_t b_ids = sizeof(int64_t) * ntest * r.K, b_s = sizeof(int32_t) * ntest * r.K, b_out = sizeof(double) * ntest;
  const size_t b_f = test_F ? sizeof(double) * (size_t)ntest * r.nF : 0;
  const size_t off_s = (b_ids + 255) / 256 * 256, off_out = off_s + (b_s + 255) / 256 * 256, off_bad = off_out + (b_out + 255) / 256 * 256;
  const size_t off_f = off_bad + 256;
  int rc_arena = bdf_ensure_arena(h, off_f + b_f + 256);
  if (rc_arena) return rc_arena;
  int64_t* d_ids = reinterpret_cast<int64_t*>(h->arena);
  int32_t* d_s = reinterpret_cast<int32_t*>(h->arena + off_s);
  double* d_out = reinterpret_cast<double*>(h->arena + off_out);
  int* d_bad = reinterpret_cast<int*>(h->arena + off_bad);
  CU(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  CU(cudaMemcpyAsync(d_ids, ids, b_ids, cudaMemcpyHostToDevice, h->stream));
  const double* Us[3] = {nullptr, nullptr, nullptr};
  for (int m = 0; m < r.K; m++) {
    EntityS& e = h->ents[r.entity_of_mode[m]];
    Us[m] = e.U;
    ids_to_slots_kernel<<<grid_for(ntest), 256, 0, h->stream>>>(d_ids + (size_t)m * ntest, ntest, e.N, h->world, e.Nper, e.slot_of_row, d_s + (size_t)m * ntest, d_bad);
  }
  predict_kernel<<<grid_for(ntest), 256, 0, h->stream>>>(r.K, Us[0], Us[1], Us[2], d_s, d_s + ntest, d_s + 2 * ntest, h->ld, h->D, ntest, r.mean, d_out);
  h->launches += 1 + r.K;
  if (test_F) {
    double* d_f = reinterpret_cast<double*>(h->arena + off_f);
    CU(cudaMemcpyAsync(d_f, test_F, b_f, cudaMemcpyHostToDevice, h->stream));
    add_fbeta_kernel<<<grid_for(ntest), 256, 0, h->stream>>>(d_f, r.beta, ntest, r.nF, d_out);
    h->launches++;
  }
  int bad = 0;
  cudaError_t ce = cudaMemcpyAsync(yhat, d_out, sizeof(double) * ntest, cudaMemcpyDeviceToHost, h->stream);
  cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t ce2 = cudaStreamSynchronize(h->stream);
  if (ce != cudaSuccess || ce2 != cudaSuccess) FAIL(BDF_ERR_CUDA, cudaGetErrorString(ce != cudaSuccess ? ce : ce2));
  if (bad) FAIL(BDF_ERR_INVALID, "test id outside 1..count of its entity");
  return BDF_OK;
}

int bdf_train_sse(bdf_t* h, int rel, double* sse, int64_t* count) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  if (!sse) FAIL(BDF_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  RelationS& r = h->rels[rel];
  ModeIndex& mi = r.modes[0];
  EntityS& e = h->ents[r.entity_of_mode[0]];
  int rc = bdf_ensure_arena(h, sizeof(double) * ((size_t)std::max(mi.n_items, 1) + 2));
  if (rc) return rc;
  double* part = reinterpret_cast<double*>(h->arena);
  double* out = part + std::max(mi.n_items, 1);
  if (mi.n_items > 0) {
    const int wpb = 8;
    sse_items_kernel<<<(mi.n_items + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>(
        mi.item_row, mi.item_beg, mi.item_len, mi.n_items, mi.col[0], mi.col[1], r.F ? mi.val_adj : mi.val, e.U, h->ents[mi.other_entity[0]].U,
        r.K > 2 ? h->ents[mi.other_entity[1]].U : nullptr, (int64_t)h->rank * e.Nper, h->ld, h->D, r.F ? 0.0 : r.mean, part);
    CU(cudaGetLastError());
  }
  sse_reduce_kernel<<<1, 256, 0, h->stream>>>(part, mi.n_items, out);
  h->launches += 2;
  CU(cudaMemcpyAsync(sse, out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (count) *count = mi.nnz;
  return BDF_OK;
}

int bdf_sample_alpha(bdf_t* h, int rel, double alpha_lambda0, double alpha_nu0, double sse, double count, double chi2_variate, double* alpha_out) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  if (!alpha_out || !(alpha_lambda0 > 0.0) || !(sse >= 0.0)) FAIL(BDF_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(h->device));
  { int rcj = bdf_join_side(h); if (rcj) return rcj; }  // h->scratch is shared with an in-flight asynchronous draw
  double* out = h->scratch;
  alpha_draw_kernel<<<1, 32, 0, h->stream>>>(sse, count, alpha_lambda0, alpha_nu0, chi2_variate, h->seed, h->sweep, philox_stream(PHILOX_ALPHA, (uint32_t)rel), out);
  h->launches++;
  CU(cudaMemcpyAsync(alpha_out, out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->rels[rel].alpha = *alpha_out;
  return BDF_OK;
}

}  // extern "C"
