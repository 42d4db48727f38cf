// Internal state of a bdf_t handle (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "row_kernel.cuh"

namespace bdf {

struct ModeIndex {            // one mode of one relation, restricted to the rows this rank owns
  int64_t nrows = 0;          // local rows (= Nper of the entity)
  int64_t nnz = 0;            // local observations
  int64_t* row_ptr = nullptr; // [nrows+1] device
  int32_t* col[2] = {nullptr, nullptr};  // partner slot indices per other mode (device)
  double* val = nullptr;
  uint32_t* perm = nullptr;   // [nnz] position in the relation's observation table of each CSR entry (relation-level features index by it)
  double* val_adj = nullptr;  // [nnz] val − F·beta_rel (relation with features: the per-observation offset replaces mean_value, src/sampling.jl:273)
  int other_entity[2] = {-1, -1};
  // work list
  int n_items = 0, n_split = 0;
  int32_t* item_row = nullptr;
  int64_t* item_beg = nullptr;
  int32_t* item_len = nullptr;
  int32_t* item_split = nullptr;
  int32_t* item_chunk = nullptr;
  int32_t* item_rel = nullptr;  // merged list of an entity in several relations: index into EntityS::uses (else nullptr)
  int32_t* split_nchunks = nullptr;
  int64_t* split_wsoff = nullptr;
  int* split_counter = nullptr;
  int32_t* split_gsize = nullptr;   // two-level reduction of the partials: chunks per group, first group index, group counters
  int64_t* split_gcoff = nullptr;
  int* group_counter = nullptr;
  int64_t chunk_slots = 0;          // workspace slots taken by chunk partials; the group partials follow
  int64_t ws_slots = 0;
  std::vector<int64_t> h_row_ptr;  // host copy of row_ptr (merged work lists are built from it)
};

struct RelationS {
  int K = 0;
  int entity_of_mode[3] = {-1, -1, -1};
  int64_t nnz = 0;
  double alpha = 1.0, mean = 0.0;
  ModeIndex modes[3];
  // relation-level features (Relation.F, src/RelationData.jl:127-160): F is nnz × nF column-major in table order
  int64_t nF = 0;
  double* F = nullptr;
  double* FF = nullptr;       // FᵀF, nF × nF (r.temp.FF)
  double* beta = nullptr;     // nF (r.model.beta)
  double* linear = nullptr;   // nnz, F·beta without the mean (r.temp.linear_values − mean_value), table order
  double* res = nullptr;      // nnz work vector (table order)
  // held-out observations registered with bdf_set_test and their posterior accumulators (predict.cu)
  int64_t ntest = 0;
  int32_t* t_slot[3] = {nullptr, nullptr, nullptr};  // factor slots per mode
  double* t_vals = nullptr;   // test values
  double* t_F = nullptr;      // ntest × nF relation-level feature rows (column-major)
  double* t_last = nullptr;   // probe_rat: the current sample's predictions
  double* t_avg = nullptr;    // probe_rat_all: running posterior mean
  double* t_sq = nullptr;     // probe_stdev: running sum of squares
  double* t_part = nullptr;   // reduction scratch
  double t_cut = 0.0;         // class_cut
  int64_t t_counter = 0;      // posterior samples accumulated
};

struct EntityS {
  int64_t N = 0, Nper = 0;  // real rows; rows per rank (slots = world*Nper)
  int64_t nlocal = 0;       // real rows owned by this rank
  double* U = nullptr;      // world*Nper × ld, slot-major
  // explicit row → slot map of an entity created with bdf_add_entity_partitioned (nnz-balanced shards); nullptr = the cyclic map
  int32_t* slot_of_row = nullptr;  // [N]
  int32_t* row_of_slot = nullptr;  // [world*Nper], -1 on padding slots
  double* peerU[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // IPC-mapped replicas on the other ranks
  double* mu = nullptr;     // D
  double* Lambda = nullptr; // D×D col-major
  double* mu_rows = nullptr;  // optional per-row mean, slot-major (ld pitch)
  double* Z = nullptr;        // injected noise staging, slot-major
  double* stats = nullptr;    // [N, NU(D), NS(D*D)]
  double* hyper = nullptr;    // [mu0(D), WI(D*D), b0, nu0] device copy for the draw kernel
  std::vector<double> mu0, WI;
  double b0 = 2.0, nu0 = 0.0;
  // bdf_nw_sample_async / _fetch: injected-variate staging, pinned result buffer [mu(D), Lambda(D*D), err flag], completion event
  double* inj = nullptr;
  double* pinned = nullptr;
  cudaEvent_t ev_done = nullptr;
  bool draw_pending = false;
  std::vector<std::pair<int, int>> uses;  // (relation, mode) pairs this entity takes part in
  ModeIndex merged;                       // work list over ALL uses (only its item_*/split_* fields), built lazily when uses.size() > 1
  size_t merged_uses = 0;                 // number of uses the merged list was built for
  // side features (Macau): sparse binary F (N × numF) as CSR and CSC index lists, link matrix beta (numF × ld, row-major)
  int64_t numF = 0, fnnz = 0;
  int64_t* f_rowptr = nullptr;  // [N+1]   rows of F
  int32_t* f_colind = nullptr;  // [fnnz]  0-based feature ids, stored order = stable sort of the COO list by row
  int64_t* f_colptr = nullptr;  // [numF+1] rows of Fᵀ
  int32_t* f_rowind = nullptr;  // [fnnz]  0-based row ids, stored order = stable sort of the COO list by column
  double* f_val_csr = nullptr;  // [fnnz] stored values in CSR order (general sparse F) or nullptr (0/1 matrix)
  double* f_val_csc = nullptr;  // [fnnz] the same in CSC order
  double* f_dense = nullptr;    // dense feature matrix (N × numF, column-major) when F was registered with bdf_set_features_dense
  double* FF = nullptr;         // full(FᵀF), numF × numF column-major (src/RelationData.jl:337-339), when the direct solve is in use
  bool use_ff = false;
  double* beta = nullptr;       // numF × ld
  double* peer_beta[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // IPC-mapped beta replicas of the other ranks
  double* uhat = nullptr;       // slots × ld, (F·beta)
  double* cgbuf = nullptr;      // CG work vectors
  double* btb = nullptr;        // [numF, colsum(D), betaᵀbeta(D×D)] in the stats layout
  void* sp_items[2] = {nullptr, nullptr};  // SpMM work lists (rows / SPLIT-index chunks of long rows) for F and Fᵀ
  void* sp_long[2] = {nullptr, nullptr};   // the split rows and their partial slots
  int sp_nitems[2] = {0, 0}, sp_nlong[2] = {0, 0};
  void* sp_part = nullptr;                 // chunk partials (slots × ld)
};

}  // namespace bdf

// shared by the translation units that implement the C ABI
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
      return BDF_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)
#define FAIL(code, msg)  \
  do {                   \
    h->err = (msg);      \
    return (code);       \
  } while (0)
#define CHECK_H()  \
  if (!h) return BDF_ERR_INVALID
#define CHECK_ENT(e) \
  if ((e) < 0 || (e) >= (int)h->ents.size()) FAIL(BDF_ERR_INVALID, "entity id out of range")

struct bdf_handle;
int bdf_ensure_arena(bdf_handle* h, size_t bytes);
int bdf_ensure_arena2(bdf_handle* h, size_t bytes);
int bdf_check_err_flag(bdf_handle* h);
void bdf_free_test(bdf::RelationS& r);

struct bdf_handle {
  int device = 0, D = 0, ld = 0, DP = 0, NW = 1, rank = 0, world = 1;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  uint64_t seed = 0x5eedULL;
  uint64_t sweep = 0;
  int64_t launches = 0;
  std::vector<bdf::EntityS> ents;
  std::vector<bdf::RelationS> rels;
  double* ws = nullptr;  // partial workspace (split rows / stats partials)
  size_t ws_bytes = 0;
  double* scratch = nullptr;  // 4 D×D matrices for the Normal-Wishart draw
  int* err_flag = nullptr;
  double* ones = nullptr;  // a row of ones (ld doubles, zero padding): the second partner of a 2-mode relation inside a 3-mode launch
  int num_sms = 148;
  double* lt = nullptr;  // Λ in tile order + Λ·μ, rebuilt per half-sweep
  void* cublas = nullptr;    // cublasHandle_t of the dense-feature products (plain dgemm / dgemv; created on first use)
  char* arena = nullptr;  // grow-only staging for host-facing calls (predict ids/slots/output, beta sampler temporaries)
  size_t arena_bytes = 0;
  char* arena2 = nullptr;  // second grow-only region: solver work space that must coexist with arena-resident operands
  size_t arena2_bytes = 0;
  int64_t pst = 0;  // doubles per parked partial for this D
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: the opt-in is remembered per handle (a handle is bound to one
  // device), one bit per kernel family — not in a process-wide static
  uint32_t smem_optin = 0;
  int* work_counter = nullptr;      // work queue head of the persistent row kernel
  bool use_ws = false;              // BDF_ROWS_WS=1 in the environment at bdf_create selects the persistent warp-specialised kernel for 64 < D <= 104
  bool async_mode = false;          // bdf_set_async
  cudaStream_t side = nullptr;      // high-priority side stream of bdf_nw_sample_async (created on first use)
  cudaEvent_t ev_ready = nullptr;
  std::string err;
};
enum { BDF_OPTIN_ROWS = 1, BDF_OPTIN_ROWS_TENSOR = 2, BDF_OPTIN_STATS = 4, BDF_OPTIN_NWDRAW = 8, BDF_OPTIN_COLORED = 16, BDF_OPTIN_ROWS_WS = 32 };
