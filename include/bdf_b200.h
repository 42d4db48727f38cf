/*
 * bdf_b200.h — C ABI of libbdf_b200.so, the B200 (sm_100a) Gibbs-sampling engine that replaces the
 * latent-factor hot path of BayesianDataFusion.jl. Plain pointers and sizes only; every entry returns
 * an int status (0 = ok, negative = error; bdf_last_error() gives the message). No C++ exceptions
 * cross this boundary and the library never calls exit().
 *
 * Conventions (they are the reference's own, so a Julia `ccall` needs no conversion):
 *   - all matrices are column-major Float64 with explicit dimensions (pointer(A) of a Julia Matrix);
 *     a factor matrix is D×N ("sample", src/RelationData.jl:15) — one latent vector per column;
 *   - indices are 1-based: Int64 for relation ids (FastIDF.ids, src/IndexedDF.jl:46-50),
 *     Int32 for sparse-binary rows/cols (src/parallel_matrix.jl:9-17);
 *   - host arrays are copied during the call and never retained; device memory belongs to the handle;
 *   - entry points taking host pointers are synchronous (they return after the stream is drained);
 *     the *_dev / step entries only enqueue work on the handle's stream.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference checkout).
 */
#ifndef BDF_B200_H
#define BDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bdf_handle bdf_t;

#define BDF_OK 0
#define BDF_ERR_INVALID (-1)  /* bad argument / DimensionMismatch / ArgumentError */
#define BDF_ERR_CUDA (-2)     /* CUDA runtime failure (message in bdf_last_error) */
#define BDF_ERR_NUMERIC (-3)  /* matrix not positive definite etc. */
#define BDF_ERR_STATE (-4)    /* call order violated (e.g. sampling before data registration) */

int bdf_version(void);
/* Message of the last failure on `h` (or of the last failed bdf_create when h == NULL). */
const char* bdf_last_error(const bdf_t* h);

/* Handle lifecycle. One handle drives ONE device; multi-GPU = one process (one handle) per GPU with
 * rank/world set here; rows of every entity are dealt out cyclically, row i (0-based) → rank i % world,
 * exactly like the reference's worker shards `i:Nprocs:N` (src/sampling.jl:154). */
int bdf_create(bdf_t** out, int device, int num_latent, int rank, int world);
int bdf_destroy(bdf_t* h);
/* Run all work of this handle on the given cudaStream_t (NULL = the handle's own stream). */
int bdf_set_stream(bdf_t* h, void* cuda_stream);
/* Philox key for the device-side noise (used whenever an injected-noise pointer is NULL). */
int bdf_set_seed(bdf_t* h, uint64_t seed);

/* Entity(name) + initModel! — src/RelationData.jl:42-90: sample=0, mu=0, Lambda=5I, mu0=0, b0=2, WI=I, nu0=D.
 * Returns the entity id (>= 0) or a negative error. */
int bdf_add_entity(bdf_t* h, int64_t count);
/* The same with an explicit shard map: rank_of_row[i] in 0..world-1 (identical on every rank) replaces the cyclic deal, e.g. an
 * nnz-balanced assignment when a few rows carry a large share of the observations (the reference balances its workers' blocks by
 * measured work in the same spirit, balanced_parallelsbm, src/parallel_matrix.jl:167). Rows keep their relative order inside a shard; results do not
 * depend on the map (noise is keyed by global row id). Must be used for all entities before they enter a relation. */
int bdf_add_entity_partitioned(bdf_t* h, int64_t count, const int32_t* rank_of_row);

/* FastIDF(rel.data) + @spawnat (src/macau.jl:51-52, src/IndexedDF.jl:46-70): registers the observation table
 * of a K-mode relation and builds one device CSR per mode (stable in table order, duplicates kept).
 * ids: nnz×K column-major, 1-based; entity_of_mode[m] = entity id of mode m. Returns the relation id. */
int bdf_add_relation(bdf_t* h, int K, const int* entity_of_mode, int64_t nnz, const int64_t* ids, const double* vals);
/* rel.model.alpha, rel.model.mean_value (src/RelationData.jl:107-118, :348). */
int bdf_set_relation_params(bdf_t* h, int rel, double alpha, double mean_value);

/* model.sample in / out: U is D×N column-major on the host. */
int bdf_set_factors(bdf_t* h, int entity, const double* U);
int bdf_get_factors(bdf_t* h, int entity, double* U);
/* Device view for the torch.distributed plumbing: base pointer of the slot-ordered factor buffer
 * (world*Nper rows × ld doubles, row-major; rank r owns rows [r*Nper, (r+1)*Nper)). */
int bdf_factors_dev(bdf_t* h, int entity, void** dev_ptr, int64_t* nper, int64_t* ld);

/* Fused all-gather (replaces the per-half-sweep shipping of `sample_m` to every worker, src/sampling.jl:155-168): each rank exports
 * the CUDA IPC handle (64 bytes) of an entity's factor buffer, the host exchanges the handles, every rank imports its peers'. From
 * then on the row-draw kernel stores each drawn row straight into every peer replica over NVLink, and no all-gather is needed;
 * the all-reduce of the Normal-Wishart statistics that follows orders the peer stores before the next half-sweep reads them. */
int bdf_ipc_export(bdf_t* h, int entity, unsigned char* handle64);
int bdf_ipc_import(bdf_t* h, int entity, int peer_rank, const unsigned char* handle64);

/* sample_latent_all2!(rel, dataRefs, procs, mode, mu_u, Lambda_u) — src/sampling.jl:149-172 (and the general
 * sample_user2_all! path, :251-289, when the entity sits in several relations): one half-sweep over `entity`,
 * overwriting its factors on the device. mu: D vector (mu_ld == 0) or D×N matrix (mu_ld == D) as in
 * src/macau.jl:102-107; Lambda: D×D; z: D×N injected standard normals in place of randn(D) per row, or NULL for
 * device Philox noise. Only rows owned by this rank are sampled. */
int bdf_sample_mode(bdf_t* h, int entity, const double* mu, int64_t mu_ld, const double* Lambda, const double* z);

/* ConditionalNormalWishart's reductions — src/sampling.jl:117-119: N = size(U,2), NU = sum(U,2), NS = U*U'
 * over the rows owned by this rank (all-reduce across ranks is the caller's, see bdf_stats_dev). */
int bdf_nw_stats(bdf_t* h, int entity, double* N, double* NU, double* NS);
/* The inverse: overwrite the statistics the next bdf_nw_sample* of `entity` reads. For callers that sum the ranks' statistics on the host
 * (Julia's master/worker remotecalls instead of an NCCL all-reduce on bdf_stats_dev): every rank calls bdf_nw_stats, the master adds the
 * (N, NU, NS) triples up, every rank calls bdf_set_nw_stats with the sums and then draws — identically, from the shared Philox key. */
int bdf_set_nw_stats(bdf_t* h, int entity, double N, const double* NU, const double* NS);
/* Device buffer holding [N, NU(D), NS(D×D col-major)] of the last bdf_nw_stats / bdf_step_nw_stats. */
int bdf_stats_dev(bdf_t* h, int entity, void** dev_ptr, int64_t* count);

/* rand(ConditionalNormalWishart(U, mu0, b0, Tinv, nu)) — src/sampling.jl:116-127 + src/normal_wishart.jl:38-42,
 * from the statistics currently in the entity's stats buffer. bartlettA (D×D lower: A[i,i]=sqrt(chi2(nu_N-i+1)),
 * A[i>j]~N(0,1)) and z (D) are the injected variates of Distributions' Wishart / MvNormal samplers; NULL = Philox.
 * Stores (mu, Lambda) as the entity's current hyper-parameters and returns them when the pointers are non-NULL. */
int bdf_nw_sample(bdf_t* h, int entity, const double* mu0, double b0, const double* Tinv, double nu,
                  const double* bartlettA, const double* z, double* mu_out, double* Lambda_out);

/* Device-resident sweep pieces (no host buffers): the loop body of src/macau.jl:96-134 for one entity, using the
 * entity's current (mu, Lambda) and Philox noise. With world > 1 the caller all-gathers bdf_factors_dev after
 * bdf_step_sample and all-reduces bdf_stats_dev after bdf_step_nw_stats. */
int bdf_step_sample(bdf_t* h, int entity);
int bdf_step_nw_stats(bdf_t* h, int entity);
int bdf_step_nw_draw(bdf_t* h, int entity);
/* The same draw enqueued on another cudaStream_t (the caller orders it after the statistics with events): the new (mu, Lambda) of an
 * entity are first needed by ITS next half-sweep, so the draw can run beside the next entity's row kernel. Uses the small-footprint
 * variant of the kernel (global scratch, <= 64 registers) that fits next to resident row-kernel CTAs. */
int bdf_step_nw_draw_on(bdf_t* h, int entity, void* cuda_stream);
/* Whole sweeps on one GPU (world == 1): for each entity {sample, stats, draw}. */
int bdf_sweep(bdf_t* h, int nsweeps);
/* Advance the Philox sweep counter by one (drivers that sequence the bdf_step_* / host-pointer entries themselves call
 * this once per Gibbs iteration; bdf_sweep does it internally). */
int bdf_advance_sweep(bdf_t* h);
int bdf_get_hyper(bdf_t* h, int entity, double* mu, double* Lambda);
int bdf_set_hyper(bdf_t* h, int entity, const double* mu, const double* Lambda);
/* The standard normals the device Philox stream yields for (entity, sweep): D×N column-major. Lets a test feed
 * the oracle the very noise a Philox-mode half-sweep used. */
int bdf_debug_row_noise(bdf_t* h, int entity, uint64_t sweep, double* z);
/* Profiling hook: one Philox half-sweep of `entity` with per-work-item phase clocks; returns the mean SM cycles of
 * [setup, gather+syrk, split-reduce, build, factorisation, substitutions, total] over the items that finalised a row. */
int bdf_debug_phase_clocks(bdf_t* h, int entity, double* mean_cycles, int64_t* n_items);
int64_t bdf_sweep_counter(const bdf_t* h);
int bdf_synchronize(bdf_t* h);
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t bdf_launch_count(const bdf_t* h);

/* pred(rel, test_vec, test_F) without relation features — src/sampling.jl:9-51: yhat[t] = sum_k prod_m U_m[k, ids[t,m]]
 * + mean_value. ids: ntest×K column-major, 1-based. */
int bdf_predict(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, double* yhat);

/* sample_alpha — src/macau.jl:84-88, src/sampling.jl:129-134. bdf_train_sse: err'err with err = pred(r) - getValues(r.data) over the
 * training observations whose first-mode row this rank owns (count = their number; all-reduce both across ranks).
 * bdf_sample_alpha: alpha = rand(Wishart(alpha_nu0 + n, inv(inv(alpha_lambda0) + err'err)))[1] = SW·chi2(alpha_nu0 + n); chi2_variate is
 * the injected chi-square draw behind the 1×1 Wishart (NaN → Philox). Stores the draw as the relation's alpha. */
int bdf_train_sse(bdf_t* h, int rel, double* sse, int64_t* count);
int bdf_sample_alpha(bdf_t* h, int rel, double alpha_lambda0, double alpha_nu0, double sse, double count, double chi2_variate, double* alpha_out);

/* pred_all(r) = udot_all(r) + mean_value — src/sampling.jl:72-97, used by macau(full_prediction = true), src/macau.jl:145-146: every cell of a
 * relation, column-major: N1 × N2 for a matrix relation (one cuBLAS dgemm of the two factor matrices), N1 × N2 × N3 for a 3-mode tensor
 * (one thread per cell, like the reference's enumeration :78-89; test/parallel_latent_tensor.jl:33-39). With several ranks every rank holds
 * all factor rows, so any rank can call it (no collective); sharded / partitioned entities are put back into row order first. */
int bdf_predict_all(bdf_t* h, int rel, double* out);

/* ---- relation-level features (Relation.F: one feature row per training observation) — src/macau.jl:89-92, src/sampling.jl:322-337 ---- */
/* r.F (nnz × nF column-major, rows in the order of the observation table given to bdf_add_relation); computes r.temp.FF = F'F, sets
 * beta = 0 and linear_values = mean_value (reset!, src/RelationData.jl:349-353). From then on the row draws and bdf_train_sse use
 * the per-observation offset linear_values[i] in place of mean_value (src/sampling.jl:273, :17-19). One GPU. */
int bdf_set_relation_features(bdf_t* h, int rel, int64_t nnz, int64_t nF, const double* F);
/* r.model.beta = sample_beta_rel(r); r.temp.linear_values = mean_value + F*beta. z1 (nnz) and z2 (nF) are the injected standard normals
 * behind randn(N) and randn(F) (consumed in that order), NULL = Philox. K = alpha*FF + lambda*I is SPD: blocked Cholesky
 * (csrc/dense_spd.cuh) for the reference's `\`. beta_out (nF) may be NULL. */
int bdf_sample_beta_rel(bdf_t* h, int rel, double lambda_beta, const double* z1, const double* z2, double* beta_out);
int bdf_get_relation_beta(bdf_t* h, int rel, double* beta);
int bdf_set_relation_beta(bdf_t* h, int rel, const double* beta);
/* pred(r, probe_vec, F) = udot + F*beta + mean_value — src/sampling.jl:9-14; test_F is ntest × nF column-major (NULL only for a relation
 * without features, where this is bdf_predict). */
int bdf_predict_f(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, const double* test_F, double* yhat);

/* ---- test set and posterior accumulators on the device — src/macau.jl:142-203 ---------------------------------- */
/* Registers the held-out observations of a relation ONCE (setTest! / assignToTest!, src/RelationData.jl:191-233): ids ntest×K
 * column-major 1-based, vals (ntest), test_F = the relation-level feature rows (ntest × nF column-major; NULL unless the relation has
 * features), class_cut = the threshold behind test_label (src/RelationData.jl:203). Replaces any earlier test set; ntest = 0 removes it.
 * With world > 1 every rank registers its own share of the test set (all ranks hold every factor row) and adds up the sums below. */
int bdf_set_test(bdf_t* h, int rel, int64_t ntest, const int64_t* ids, const double* vals, const double* test_F, double class_cut);
/* Forget the posterior accumulators (counter_prob = 0). */
int bdf_test_reset(bdf_t* h, int rel);
/* One iteration of src/macau.jl:143-200 on the device: probe_rat = pred(rel, test_vec, test_F); posterior == 0 (burn-in):
 * probe_rat_all = probe_rat; posterior != 0: the first such call sets probe_rat_all = probe_rat, probe_stdev = probe_rat.^2, later ones
 * probe_rat_all = (counter_prob*probe_rat_all + probe_rat)/(counter_prob + 1), probe_stdev += probe_rat.^2 (:164-176). Returns
 * out[0] = sum((test_values - makeClamped(probe_rat_all, clamp)).^2)   (rmse_avg = sqrt(out[0]/out[3]), :196)
 * out[1] = the same for the current sample probe_rat
 * out[2] = sum(test_label .== (probe_rat_all .< class_cut))            (err_avg = out[2]/out[3], :193-194)
 * out[3] = ntest, out[4] = counter_prob after this call.
 * clamp_lo / clamp_hi: NaN = not clamped (clamp = Float64[], src/sampling.jl:99-106). The only device→host traffic is these 5 doubles. */
int bdf_predict_accumulate(bdf_t* h, int rel, int posterior, double clamp_lo, double clamp_hi, double* out5);
/* probe_rat_all (unclamped), probe_stdev (the running sum of squares) and the last sample's probe_rat, ntest doubles each; any may be
 * NULL. For result["predictions"] (src/macau.jl:233-241) and ROC (src/ROC.jl), which stay on the host. */
int bdf_get_test_predictions(bdf_t* h, int rel, double* avg_out, double* sumsq_out, double* last_out);

/* ---- deferred completion of the host-pointer seams ------------------------------------------------------------------------------- */
/* on != 0: bdf_sample_mode / bdf_sample_mode_uhat return as soon as their work is enqueued (the host arrays have been copied by then);
 * a numeric failure (precision matrix not positive definite) is reported by the next call that synchronises — bdf_nw_stats*,
 * bdf_nw_sample*, bdf_predict*, bdf_get_*, bdf_synchronize. Default 0: every host-pointer entry returns after the stream is drained. */
int bdf_set_async(bdf_t* h, int on);
/* rand(ConditionalNormalWishart(...)) split in two so that the draw of one entity overlaps the next entity's half-sweep — the draw is first
 * needed by the SAME entity's next half-sweep (src/macau.jl:96-134 keeps entities independent within an iteration). _async copies the
 * host arguments, runs the draw on the handle's high-priority side stream and returns; _fetch blocks until that draw is done, returns
 * (mu, Lambda) and orders it before whatever the handle does next. Same arguments and results as bdf_nw_sample. */
int bdf_nw_sample_async(bdf_t* h, int entity, const double* mu0, double b0, const double* Tinv, double nu, const double* bartlettA, const double* z);
int bdf_nw_sample_fetch(bdf_t* h, int entity, double* mu_out, double* Lambda_out);

/* ---- Macau side features: the link-matrix (beta) path ------------------------------------------------------------- */

/* Entity(F = SparseBinMatrix(m, n, rows, cols)) — src/parallel_matrix.jl:9-24: registers a sparse 0/1 feature matrix given
 * as COO index lists (Int32, 1-based, any order, duplicates counted twice). Builds the CSR of F (the SparseBinMatrixCSR
 * constructor, src/sparsebin_csr.jl:22-37: stable sort by row) and of Fᵀ on the device, allocates beta = zeros(n, D)
 * (src/RelationData.jl:76). m must equal the entity count (src/RelationData.jl:263-268). With world > 1 every rank registers the whole
 * F and keeps the whole beta: the beta draw is replicated (identical on all ranks), uhat is filed by slot, each rank samples and reduces
 * its own rows. */
int bdf_set_features_sbm(bdf_t* h, int entity, int64_t m, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols);
/* Entity(F = ::SparseMatrixCSC{Float64,Int64}) — a general sparse feature matrix given by Julia's CSC fields (colptr n+1, rowval nnz,
 * nzval nnz; 1-based), as in the reference's own tests (test/parallel_latent_basic.jl:4, test/parallel_mult.jl:4-18). Products
 * accumulate in Julia's order with un-fused multiply-add. */
int bdf_set_features_csc(bdf_t* h, int entity, int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval);
/* Entity(F = ::Matrix{Float64}) — a dense feature matrix (m × n column-major, m = entity count). The feature-operator duck type
 * (src/RelationData.jl:314-329) then runs on cuBLAS dgemm (a plain library GEMM; its summation order is the library's). */
int bdf_set_features_dense(bdf_t* h, int entity, int64_t m, int64_t n, const double* F);
/* en.FF = full(At_mul_B(en.F, en.F)); en.use_FF = true — reset!, src/RelationData.jl:337-339 (numF <= compute_ff_size). Afterwards
 * bdf_sample_beta solves with solve_full instead of CG (src/sampling.jl:303-304). FF_out (n × n column-major) may be NULL. */
int bdf_compute_ff(bdf_t* h, int entity, double* FF_out);
/* en.use_FF — switch between the direct solve and CG (the FF matrix stays resident). */
int bdf_set_use_ff(bdf_t* h, int entity, int use_ff);
/* solve_full(FF, rhs, lambda) — src/sampling.jl:314-320: (FF + lambda·I) \ rhs, rhs and x n × ncol column-major, ncol == num_latent.
 * The regularised matrix is symmetric positive definite; the device factorisation is a blocked Cholesky (csrc/dense_spd.cuh). */
int bdf_solve_full(bdf_t* h, int entity, const double* rhs, int ncol, double lambda, double* x);
/* Parity hook: the device CSR in the reference's representation — row_ptr (m+1, or n+1 for the transpose) and col_ind
 * (nnz), Int32, 1-based (fields of SparseBinMatrixCSR, src/sparsebin_csr.jl:6-11). */
int bdf_debug_features_csr(bdf_t* h, int entity, int transpose, int32_t* ptr_out, int32_t* ind_out);
/* The feature-operator duck type (src/RelationData.jl:314-329): F*X (transpose == 0, X is n×ncol → Y m×ncol) and
 * At_mul_B(F, X) (transpose != 0, X is m×ncol → Y n×ncol); A_mul_B!/At_mul_B! of src/parallel_matrix.jl:242-267 and
 * src/sparsebin_csr.jl:49-63 for ncol columns at once. Column-major host matrices, ncol <= num_latent. */
int bdf_spmm(bdf_t* h, int entity, int transpose, const double* X, int ncol, double* Y);
/* AtA_mul_B!(y, F, x, lambda): y = Fᵀ(F·x) + lambda·x — src/parallel_cg.jl:7-14. */
int bdf_ata_mul(bdf_t* h, int entity, const double* x, double lambda, double* y);
/* solve_cg2(Frefs, rhs, lambda; tol, maxiter) — src/parallel_matrix.jl:488-507 — one cg_AtA (src/parallel_cg.jl:63-94) per
 * column, all num_latent columns advanced together with per-column scalars and convergence masks. rhs, x: n×ncol
 * column-major, ncol == num_latent; tol NaN → eps()·n (src/sampling.jl:294-296); maxiter <= 0 → n; iters (ncol) may be NULL. */
int bdf_cg_solve(bdf_t* h, int entity, const double* rhs, int ncol, double lambda, double tol, int64_t maxiter, double* x, int* iters);
/* model.beta in / out (n × num_latent, column-major). */
int bdf_set_beta(bdf_t* h, int entity, const double* beta);
int bdf_get_beta(bdf_t* h, int entity, double* beta);
/* mj.uhat = F_mul_beta(en)'; mu_matrix = mj.mu .+ mj.uhat — src/macau.jl:102-104, src/RelationData.jl:314-320. The per-row mean
 * stays on the device for bdf_sample_mode_uhat; uhat_out (D×N) may be NULL. */
int bdf_update_uhat(bdf_t* h, int entity, const double* mu, double* uhat_out);
/* sample_latent_all2!(..., mu_matrix, Lambda) — src/macau.jl:105 — with the mu_matrix of the last bdf_update_uhat. */
int bdf_sample_mode_uhat(bdf_t* h, int entity, const double* Lambda, const double* z);
/* ConditionalNormalWishart reductions of U = mj.sample - mj.uhat — src/macau.jl:124. */
int bdf_nw_stats_uhat(bdf_t* h, int entity, double* N, double* NU, double* NS);
/* mj.beta' * mj.beta (D×D) — src/macau.jl:128, src/sampling.jl:138. */
int bdf_beta_gram(bdf_t* h, int entity, double* BtB);
/* sample_beta(entity, sample .- mu, Lambda_u, lambda_beta, use_ff = false, tol) — src/sampling.jl:291-312, via update_beta!
 * (:361-370): rhs = Fᵀ((U - mu)ᵀ + E1c) + sqrt(lambda_beta)·E2c with E·c = chol_lower(inv(Lambda))·E, then CG. E1 (D×N) and
 * E2 (D×n) are the injected standard normals behind rand(mv, N) and rand(mv, numF) (consumed in that order); NULL → Philox.
 * beta stays on the device; beta_out / rhs_out (n×D, column-major) and iters_out (D) may be NULL. */
int bdf_sample_beta(bdf_t* h, int entity, const double* mu, const double* Lambda, double lambda_beta, double tol, const double* E1,
                    const double* E2, double* beta_out, double* rhs_out, int* iters_out);
/* Column-split beta solve over the ranks — solve_cg2 hands the num_latent right-hand sides to its workers one column at a time
 * (src/parallel_matrix.jl:488-507). Exchange the CUDA IPC handle of every rank's beta buffer once (after bdf_set_features_*); with all
 * world-1 peers mapped, bdf_sample_beta (CG path, sparse F) solves only columns [D·rank/world, D·(rank+1)/world) on this rank — the gathers
 * of the two products are as narrow as the window — and stores the solved columns into its own and every peer's replica over NVLink. The
 * caller must synchronise the ranks (any collective on the handles' streams, e.g. the all-reduce of iters_out) before beta is used, and
 * pass beta_out = rhs_out = NULL. Without the mappings every rank solves all columns (replicated). */
int bdf_ipc_export_beta(bdf_t* h, int entity, unsigned char* handle64);
int bdf_ipc_import_beta(bdf_t* h, int entity, int peer_rank, const unsigned char* handle64);
/* Profiling hook: mean milliseconds of one device-resident AtA_mul_B! (src/parallel_cg.jl:7-14) on all num_latent columns. */
int bdf_debug_ata_time(bdf_t* h, int entity, int reps, double* ms_per_apply);
/* The same on a packed window of `ncols` < num_latent columns (the operand shape of one rank of a column-split CG). */
int bdf_debug_ata_time_window(bdf_t* h, int entity, int reps, int ncols, double* ms_per_apply);
/* sample_lambda_beta(beta, Lambda_u, nu, mu) — src/sampling.jl:136-142. gamma_variate = the injected Gamma(shape, 1) draw behind
 * rand(Gamma(b, c)), NaN → Philox. shape_out may be NULL. */
int bdf_sample_lambda_beta(bdf_t* h, int entity, const double* Lambda, double nu, double mu, double gamma_variate, double* lambda_beta_out,
                           double* shape_out);

#ifdef __cplusplus
}
#endif
#endif /* BDF_B200_H */
