// Test-set prediction with the posterior accumulators on the device — SURVEY §8f N1; replaces the host-side bookkeeping of
// src/macau.jl:142-186 (probe_rat = pred(rel, test_vec, test_F); running posterior mean probe_rat_all, sum of squares probe_stdev,
// clamped RMSE, accuracy against class_cut; makeClamped src/sampling.jl:99-106).
//
// The held-out observations are registered ONCE (bdf_set_test: ids → factor slots, values, class labels, optional relation-level
// feature rows); every sweep then costs one kernel over the test set and a 40-byte device→host read of the reduced scalars,
// instead of re-uploading the ids and downloading one prediction per observation.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/bdf_b200.h"
#include "engine.cuh"

using namespace bdf;

namespace {

inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

__global__ void test_slots_kernel(const int64_t* ids, int64_t n, int64_t N, int world, int64_t nper, const int32_t* __restrict__ tab, int32_t* out, int* bad) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = ids[o];
    if (id < 1 || id > N) { *bad = 1; out[o] = 0; }
    else out[o] = tab ? tab[id - 1] : (int32_t)(((id - 1) % world) * nper + (id - 1) / world);
  }
}

struct TestParams {
  int K, ld, D;
  const double* U[3];
  const int32_t* slot[3];
  const double* vals;
  const double* F;      // ntest × nF column-major relation-level feature rows, or nullptr
  const double* beta;   // nF
  int64_t nF;
  int64_t nt;
  double mean, lo, hi, cut;
  int mode;             // 0: burn-in (probe_rat_all = probe_rat); 1: first posterior sample; 2: later posterior samples
  double counter;       // posterior samples averaged so far (mode 2), src/macau.jl:173-176
  double* last;         // probe_rat
  double* avg;          // probe_rat_all
  double* sq;           // probe_stdev (running sum of squares)
  double* part;         // [gridDim.x][4] block partials
};

__device__ __forceinline__ double clampd(double x, double lo, double hi) {
  // makeClamped (src/sampling.jl:99-106): x < lo → lo, x > hi → hi; NaN bounds = no clamping
  if (lo == lo && x < lo) x = lo;
  if (hi == hi && x > hi) x = hi;
  return x;
}

__global__ void __launch_bounds__(256) predict_accumulate_kernel(const TestParams p) {
  double s_avg = 0.0, s_cur = 0.0, s_ok = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.nt; t += (int64_t)gridDim.x * blockDim.x) {
    const double* a = p.U[0] + (size_t)p.slot[0][t] * p.ld;
    const double* b = p.U[1] + (size_t)p.slot[1][t] * p.ld;
    const double* c = p.K > 2 ? p.U[2] + (size_t)p.slot[2][t] * p.ld : nullptr;
    double s = 0.0;
    for (int k = 0; k < p.D; k++) {   // same order as predict_kernel / udot, src/sampling.jl:27-51
      double pr = a[k] * b[k];
      if (c) pr *= c[k];
      s += pr;
    }
    double y = s + p.mean;
    if (p.F) {
      double f = 0.0;
      for (int64_t j = 0; j < p.nF; j++) f = fma(p.F[t + j * p.nt], p.beta[j], f);
      y += f;
    }
    p.last[t] = y;
    double all;
    if (p.mode == 2) {
      all = (p.counter * p.avg[t] + y) / (p.counter + 1.0);   // src/macau.jl:174
      p.sq[t] += y * y;                                        // :175
    } else {
      all = y;
      if (p.mode == 1) p.sq[t] = y * y;                        // :171
    }
    p.avg[t] = all;
    const double v = p.vals[t];
    const double ea = v - clampd(all, p.lo, p.hi), ec = v - clampd(y, p.lo, p.hi);
    s_avg = fma(ea, ea, s_avg);
    s_cur = fma(ec, ec, s_cur);
    s_ok += ((v < p.cut) == (all < p.cut)) ? 1.0 : 0.0;       // rel.test_label .== (probe_rat_all .< class_cut), :193
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = s_avg; sh[1][threadIdx.x] = s_cur; sh[2][threadIdx.x] = s_ok;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 3) p.part[(size_t)blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
}

// block partials → out[0..2] in block order (deterministic)
__global__ void test_reduce_kernel(const double* __restrict__ part, int nblk, double* __restrict__ out) {
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += part[(size_t)b * 4 + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

}  // namespace

void bdf_free_test(RelationS& r) {
  for (int m = 0; m < 3; m++) { cudaFree(r.t_slot[m]); r.t_slot[m] = nullptr; }
  cudaFree(r.t_vals); cudaFree(r.t_F); cudaFree(r.t_last); cudaFree(r.t_avg); cudaFree(r.t_sq); cudaFree(r.t_part);
  r.t_vals = r.t_F = r.t_last = r.t_avg = r.t_sq = r.t_part = nullptr;
  r.ntest = 0; r.t_counter = 0;
}

extern "C" {

int bdf_set_test(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, const double* vals, const double* test_F, double class_cut) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (ntest < 0 || ntest >= 2147483647LL || (ntest > 0 && (!ids || !vals))) FAIL(BDF_ERR_INVALID, "bad test set");
  if (r.F && ntest > 0 && !test_F) FAIL(BDF_ERR_INVALID, "Relation has features, please supply features with test data");  // src/RelationData.jl:215-217
  if (!r.F && test_F) FAIL(BDF_ERR_INVALID, "the relation has no features");
  CU(cudaSetDevice(h->device));
  bdf_free_test(r);
  if (ntest == 0) return BDF_OK;
  const size_t b_ids = sizeof(int64_t) * (size_t)ntest * r.K;
  int rc = bdf_ensure_arena(h, b_ids + 256);
  if (rc) return rc;
  int64_t* d_ids = reinterpret_cast<int64_t*>(h->arena);
  int* d_bad = reinterpret_cast<int*>(h->arena + (b_ids + 255) / 256 * 256);
  CU(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  CU(cudaMemcpyAsync(d_ids, ids, b_ids, cudaMemcpyHostToDevice, h->stream));
  for (int m = 0; m < r.K; m++) {
    EntityS& e = h->ents[r.entity_of_mode[m]];
    CU(cudaMalloc((void**)&r.t_slot[m], sizeof(int32_t) * (size_t)ntest));
    test_slots_kernel<<<grid_for(ntest), 256, 0, h->stream>>>(d_ids + (size_t)m * ntest, ntest, e.N, h->world, e.Nper, e.slot_of_row, r.t_slot[m], d_bad);
  }
  CU(cudaGetLastError());
  const size_t bv = sizeof(double) * (size_t)ntest;
  CU(cudaMalloc((void**)&r.t_vals, bv)); CU(cudaMalloc((void**)&r.t_last, bv)); CU(cudaMalloc((void**)&r.t_avg, bv)); CU(cudaMalloc((void**)&r.t_sq, bv));
  CU(cudaMalloc((void**)&r.t_part, sizeof(double) * (4 * 148 * 8 + 8)));
  CU(cudaMemcpyAsync(r.t_vals, vals, bv, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(r.t_last, 0, bv, h->stream)); CU(cudaMemsetAsync(r.t_avg, 0, bv, h->stream)); CU(cudaMemsetAsync(r.t_sq, 0, bv, h->stream));
  if (test_F) {
    CU(cudaMalloc((void**)&r.t_F, bv * (size_t)r.nF));
    CU(cudaMemcpyAsync(r.t_F, test_F, bv * (size_t)r.nF, cudaMemcpyHostToDevice, h->stream));
  }
  int bad = 0;
  CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (bad) { bdf_free_test(r); FAIL(BDF_ERR_INVALID, "test id outside 1..count of its entity"); }
  r.ntest = ntest; r.t_cut = class_cut; r.t_counter = 0;
  return BDF_OK;
}

int bdf_test_reset(bdf_t* h, int rel) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  h->rels[rel].t_counter = 0;
  return BDF_OK;
}

int bdf_predict_accumulate(bdf_t* h, int rel, int posterior, double clamp_lo, double clamp_hi, double* out5) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (!out5) FAIL(BDF_ERR_INVALID, "null argument");
  if (r.ntest <= 0) FAIL(BDF_ERR_STATE, "no test set registered (bdf_set_test)");
  CU(cudaSetDevice(h->device));
  TestParams p{};
  p.K = r.K; p.ld = h->ld; p.D = h->D;
  for (int m = 0; m < r.K; m++) { p.U[m] = h->ents[r.entity_of_mode[m]].U; p.slot[m] = r.t_slot[m]; }
  p.vals = r.t_vals; p.F = r.F ? r.t_F : nullptr; p.beta = r.beta; p.nF = r.nF; p.nt = r.ntest;
  p.mean = r.mean; p.lo = clamp_lo; p.hi = clamp_hi; p.cut = r.t_cut;
  p.mode = !posterior ? 0 : (r.t_counter == 0 ? 1 : 2);
  p.counter = (double)r.t_counter;
  p.last = r.t_last; p.avg = r.t_avg; p.sq = r.t_sq; p.part = r.t_part;
  const int g = grid_for(r.ntest);
  predict_accumulate_kernel<<<g, 256, 0, h->stream>>>(p);
  double* dout = r.t_part + (size_t)4 * 148 * 8;
  test_reduce_kernel<<<1, 32, 0, h->stream>>>(r.t_part, g, dout);
  h->launches += 2;
  CU(cudaGetLastError());
  if (posterior) r.t_counter++;
  double res[3];
  CU(cudaMemcpyAsync(res, dout, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  out5[0] = res[0]; out5[1] = res[1]; out5[2] = res[2]; out5[3] = (double)r.ntest; out5[4] = (double)r.t_counter;
  return bdf_check_err_flag(h);
}

int bdf_get_test_predictions(bdf_t* h, int rel, double* avg_out, double* sumsq_out, double* last_out) {
  CHECK_H();
  if (rel < 0 || rel >= (int)h->rels.size()) FAIL(BDF_ERR_INVALID, "relation id out of range");
  RelationS& r = h->rels[rel];
  if (r.ntest <= 0) FAIL(BDF_ERR_STATE, "no test set registered (bdf_set_test)");
  CU(cudaSetDevice(h->device));
  const size_t bv = sizeof(double) * (size_t)r.ntest;
  if (avg_out) CU(cudaMemcpyAsync(avg_out, r.t_avg, bv, cudaMemcpyDeviceToHost, h->stream));
  if (sumsq_out) CU(cudaMemcpyAsync(sumsq_out, r.t_sq, bv, cudaMemcpyDeviceToHost, h->stream));
  if (last_out) CU(cudaMemcpyAsync(last_out, r.t_last, bv, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BDF_OK;
}

}  // extern "C"
