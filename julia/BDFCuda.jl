# BDFCuda.jl — reference-side binding of libbdf_b200.so (include/bdf_b200.h) for BayesianDataFusion.jl.
#
# NOT runnable in the build image (no Julia there); the ABI it binds is pinned by tests/test_abi.py and exercised
# through the identical ctypes table in bayesiandatafusion.jl_b200/_lib.py. Written for Julia 0.4/0.5 syntax like the
# reference (`Ptr{Void}`; on Julia >= 0.7 replace with `Ptr{Cvoid}`).
#
# Usage inside src/macau.jl (see INTEGRATION.md): `macau(data; backend = :cuda, devices = [0])` swaps
#   sample_latent_all2!      -> BDFCuda.sample_mode!
#   ConditionalNormalWishart -> BDFCuda.nw_stats + BDFCuda.nw_sample
#   update_beta!             -> BDFCuda.sample_beta! / BDFCuda.sample_lambda_beta
#   pred                     -> BDFCuda.predict
module BDFCuda

const LIB = get(ENV, "BDF_B200_LIB", "libbdf_b200.so")

type Handle
  ptr::Ptr{Void}
end

function check(h::Handle, rc::Cint)
  if rc != 0
    msg = bytestring(ccall((:bdf_last_error, LIB), Ptr{UInt8}, (Ptr{Void},), h.ptr))
    rc == -1 && throw(ArgumentError(msg))   # BDF_ERR_INVALID ~ DimensionMismatch / ArgumentError in the reference
    error("libbdf_b200 ($rc): $msg")
  end
  nothing
end

function create(num_latent::Int; device::Int = 0, rank::Int = 0, world::Int = 1)
  out = Ref{Ptr{Void}}(C_NULL)
  rc = ccall((:bdf_create, LIB), Cint, (Ptr{Ptr{Void}}, Cint, Cint, Cint, Cint), out, device, num_latent, rank, world)
  rc == 0 || error(bytestring(ccall((:bdf_last_error, LIB), Ptr{UInt8}, (Ptr{Void},), C_NULL)))
  h = Handle(out[])
  finalizer(h, x -> ccall((:bdf_destroy, LIB), Cint, (Ptr{Void},), x.ptr))
  return h
end

function add_entity(h::Handle, count::Integer)
  e = ccall((:bdf_add_entity, LIB), Cint, (Ptr{Void}, Int64), h.ptr, count)
  e >= 0 || check(h, e)
  return e
end

## FastIDF(rel.data) — ids is nnz x K Int64 (1-based), values Float64
function add_relation(h::Handle, entities::Vector{Cint}, ids::Matrix{Int64}, values::Vector{Float64})
  r = ccall((:bdf_add_relation, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cint}, Int64, Ptr{Int64}, Ptr{Cdouble}),
            h.ptr, size(ids, 2), entities, size(ids, 1), ids, values)
  r >= 0 || check(h, r)
  return r
end

set_relation_params(h::Handle, rel, alpha, mean_value) =
  check(h, ccall((:bdf_set_relation_params, LIB), Cint, (Ptr{Void}, Cint, Cdouble, Cdouble), h.ptr, rel, alpha, mean_value))

## model.sample (D x N, column-major) in / out
set_factors(h::Handle, entity, U::Matrix{Float64}) = check(h, ccall((:bdf_set_factors, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, U))
get_factors!(h::Handle, entity, U::Matrix{Float64}) = check(h, ccall((:bdf_get_factors, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, U))

## sample_latent_all2!(rel, dataRefs, procs, mode, mu_u, Lambda_u) — src/sampling.jl:149
function sample_mode!(h::Handle, entity, mu_u::Vector{Float64}, Lambda_u::Matrix{Float64}; z = C_NULL)
  check(h, ccall((:bdf_sample_mode, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                 h.ptr, entity, mu_u, 0, Lambda_u, z))
end
function sample_mode!(h::Handle, entity, mu_u::Matrix{Float64}, Lambda_u::Matrix{Float64}; z = C_NULL)
  check(h, ccall((:bdf_sample_mode, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                 h.ptr, entity, mu_u, size(mu_u, 1), Lambda_u, z))
end

## N, NU, NS of ConditionalNormalWishart — src/sampling.jl:117-119
function nw_stats(h::Handle, entity, D::Int)
  N = Ref{Cdouble}(0.0); NU = zeros(D); NS = zeros(D, D)
  check(h, ccall((:bdf_nw_stats, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, N, NU, NS))
  return N[], NU, NS
end

## rand(ConditionalNormalWishart(U, mu0, b0, Tinv, nu)) with the Bartlett factor A and z drawn by the caller (or C_NULL)
function nw_sample(h::Handle, entity, mu0::Vector{Float64}, b0, Tinv::Matrix{Float64}, nu; A = C_NULL, z = C_NULL)
  D = length(mu0); mu = zeros(D); Lambda = zeros(D, D)
  check(h, ccall((:bdf_nw_sample, LIB), Cint,
                 (Ptr{Void}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                 h.ptr, entity, mu0, b0, Tinv, nu, A, z, mu, Lambda))
  return mu, Lambda
end

## Entity(F = SparseBinMatrix(m, n, rows, cols))
set_features_sbm(h::Handle, entity, m, n, rows::Vector{Int32}, cols::Vector{Int32}) =
  check(h, ccall((:bdf_set_features_sbm, LIB), Cint, (Ptr{Void}, Cint, Int64, Int64, Int64, Ptr{Int32}, Ptr{Int32}),
                 h.ptr, entity, m, n, length(rows), rows, cols))

## Entity(F = ::SparseMatrixCSC{Float64,Int64}) — general sparse features by their CSC fields
set_features_csc(h::Handle, entity, F::SparseMatrixCSC{Float64,Int64}) =
  check(h, ccall((:bdf_set_features_csc, LIB), Cint, (Ptr{Void}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}),
                 h.ptr, entity, size(F, 1), size(F, 2), F.colptr, F.rowval, F.nzval))

## the feature-operator duck type (src/RelationData.jl:314-329): a type CudaSBM can forward *, At_mul_B, AtA_mul_B! here
function spmm(h::Handle, entity, X::Matrix{Float64}, nout::Int; transpose::Bool = false)
  Y = zeros(nout, size(X, 2))
  check(h, ccall((:bdf_spmm, LIB), Cint, (Ptr{Void}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}), h.ptr, entity, transpose, X, size(X, 2), Y))
  return Y
end

## sample_beta(entity, sample .- mu, Lambda_u, lambda_beta, false, tol) — src/sampling.jl:291
function sample_beta!(h::Handle, entity, mu::Vector{Float64}, Lambda::Matrix{Float64}, lambda_beta, tol, numF::Int; E1 = C_NULL, E2 = C_NULL)
  D = length(mu); beta = zeros(numF, D); rhs = zeros(numF, D); iters = zeros(Cint, D)
  check(h, ccall((:bdf_sample_beta, LIB), Cint,
                 (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                 h.ptr, entity, mu, Lambda, lambda_beta, tol, E1, E2, beta, rhs, iters))
  return beta, rhs
end

## the same draw with beta left on the device (get_beta! fetches it when the host wants it): nothing comes back per iteration
function sample_beta_device!(h::Handle, entity, mu::Vector{Float64}, Lambda::Matrix{Float64}, lambda_beta, tol; E1 = C_NULL, E2 = C_NULL)
  iters = zeros(Cint, length(mu))
  check(h, ccall((:bdf_sample_beta, LIB), Cint,
                 (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                 h.ptr, entity, mu, Lambda, lambda_beta, tol, E1, E2, C_NULL, C_NULL, iters))
  return iters
end

function sample_lambda_beta(h::Handle, entity, Lambda::Matrix{Float64}, nu, mu; gamma_variate = NaN)
  out = Ref{Cdouble}(0.0)
  check(h, ccall((:bdf_sample_lambda_beta, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                 h.ptr, entity, Lambda, nu, mu, gamma_variate, out, C_NULL))
  return out[]
end

## pred(rel, test_vec, test_F) without relation features — src/sampling.jl:9-14
function predict(h::Handle, rel, ids::Matrix{Int64})
  yhat = zeros(size(ids, 1))
  check(h, ccall((:bdf_predict, LIB), Cint, (Ptr{Void}, Cint, Int64, Ptr{Int64}, Ptr{Cdouble}), h.ptr, rel, size(ids, 1), ids, yhat))
  return yhat
end

## Entity(F = ::Matrix{Float64}) — a dense feature matrix
set_features_dense(h::Handle, entity, F::Matrix{Float64}) =
  check(h, ccall((:bdf_set_features_dense, LIB), Cint, (Ptr{Void}, Cint, Int64, Int64, Ptr{Cdouble}), h.ptr, entity, size(F, 1), size(F, 2), F))

## reset!: en.FF = full(At_mul_B(en.F, en.F)); en.use_FF = true — src/RelationData.jl:337-339; sample_beta then takes solve_full
compute_ff!(h::Handle, entity) = check(h, ccall((:bdf_compute_ff, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, C_NULL))
set_use_ff!(h::Handle, entity, use_ff::Bool) = check(h, ccall((:bdf_set_use_ff, LIB), Cint, (Ptr{Void}, Cint, Cint), h.ptr, entity, use_ff))

## solve_full(FF, rhs, lambda) — src/sampling.jl:314-320
function solve_full(h::Handle, entity, rhs::Matrix{Float64}, lambda)
  x = zeros(size(rhs))
  check(h, ccall((:bdf_solve_full, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}), h.ptr, entity, rhs, size(rhs, 2), lambda, x))
  return x
end

## sample_alpha — src/macau.jl:84-88: err'err over this handle's training observations on the device (with several ranks: add the ranks'
## (sse, n) up first), then the 1x1 Wishart draw (chi2 = injected variate or NaN for Philox); the relation's alpha is updated on the device
function train_sse(h::Handle, rel)
  sse = Ref{Cdouble}(0.0); n = Ref{Int64}(0)
  check(h, ccall((:bdf_train_sse, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Int64}), h.ptr, rel, sse, n))
  return sse[], Float64(n[])
end
function sample_alpha(h::Handle, rel, alpha_lambda0, alpha_nu0, sse, n; chi2 = NaN)
  alpha = Ref{Cdouble}(0.0)
  check(h, ccall((:bdf_sample_alpha, LIB), Cint, (Ptr{Void}, Cint, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}),
                 h.ptr, rel, alpha_lambda0, alpha_nu0, sse, n, chi2, alpha))
  return alpha[]
end
function sample_alpha!(h::Handle, rel, alpha_lambda0, alpha_nu0; chi2 = NaN)
  sse, n = train_sse(h, rel)
  return sample_alpha(h, rel, alpha_lambda0, alpha_nu0, sse, n, chi2 = chi2)
end

## an entity with an explicit (e.g. work-balanced) shard map instead of the cyclic i:Nprocs:N deal; rank_of_row is 0-based
function add_entity_partitioned(h::Handle, count::Integer, rank_of_row::Vector{Int32})
  e = ccall((:bdf_add_entity_partitioned, LIB), Cint, (Ptr{Void}, Int64, Ptr{Int32}), h.ptr, count, rank_of_row)
  e >= 0 || check(h, e)
  return e                                                  # the entity id, like add_entity
end

## pred_all(r) — src/sampling.jl:92-97 (macau(full_prediction = true), src/macau.jl:145-146)
function predict_all(h::Handle, rel, n1::Integer, n2::Integer)
  out = zeros(n1, n2)
  check(h, ccall((:bdf_predict_all, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, rel, out))
  return out
end

## relation-level features: r.F (numData(r) x nF), sample_beta_rel (src/sampling.jl:322-337), pred(r, probe_vec, F) (:9-14)
set_relation_features(h::Handle, rel, F::Matrix{Float64}) =
  check(h, ccall((:bdf_set_relation_features, LIB), Cint, (Ptr{Void}, Cint, Int64, Int64, Ptr{Cdouble}), h.ptr, rel, size(F, 1), size(F, 2), F))
function sample_beta_rel!(h::Handle, rel, lambda_beta, nF::Int; z1 = C_NULL, z2 = C_NULL)
  beta = zeros(nF)
  check(h, ccall((:bdf_sample_beta_rel, LIB), Cint, (Ptr{Void}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, rel, lambda_beta, z1, z2, beta))
  return beta
end
function predict(h::Handle, rel, ids::Matrix{Int64}, test_F::Matrix{Float64})
  yhat = zeros(size(ids, 1))
  check(h, ccall((:bdf_predict_f, LIB), Cint, (Ptr{Void}, Cint, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, rel, size(ids, 1), ids, test_F, yhat))
  return yhat
end

## ---- the rest of the sweep: features on an entity (src/macau.jl:102-105, 124-129), device-resident sweeps, multi-GPU plumbing ----
set_seed(h::Handle, seed::Integer) = check(h, ccall((:bdf_set_seed, LIB), Cint, (Ptr{Void}, UInt64), h.ptr, seed))
advance_sweep(h::Handle) = check(h, ccall((:bdf_advance_sweep, LIB), Cint, (Ptr{Void},), h.ptr))
synchronize(h::Handle) = check(h, ccall((:bdf_synchronize, LIB), Cint, (Ptr{Void},), h.ptr))

## mj.uhat = F_mul_beta(en)'; mu_matrix = mj.mu .+ mj.uhat (kept on the device for sample_mode_uhat!)
function update_uhat!(h::Handle, entity, mu::Vector{Float64}, uhat::Matrix{Float64})
  check(h, ccall((:bdf_update_uhat, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, mu, uhat))
end
sample_mode_uhat!(h::Handle, entity, Lambda_u::Matrix{Float64}; z = C_NULL) =
  check(h, ccall((:bdf_sample_mode_uhat, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, Lambda_u, z))
function nw_stats_uhat(h::Handle, entity, D::Int)
  N = Ref{Cdouble}(0.0); NU = zeros(D); NS = zeros(D, D)
  check(h, ccall((:bdf_nw_stats_uhat, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, N, NU, NS))
  return N[], NU, NS
end
function beta_gram(h::Handle, entity, D::Int)
  BtB = zeros(D, D)
  check(h, ccall((:bdf_beta_gram, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, BtB))
  return BtB
end
get_beta!(h::Handle, entity, beta::Matrix{Float64}) = check(h, ccall((:bdf_get_beta, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, beta))
set_beta(h::Handle, entity, beta::Matrix{Float64}) = check(h, ccall((:bdf_set_beta, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}), h.ptr, entity, beta))

## AtA_mul_B!(y, F, x, lambda) and cg_AtA / solve_cg2 for a CudaSBM operator type (src/parallel_cg.jl:7-14, 63-94)
function ata_mul(h::Handle, entity, x::Vector{Float64}, lambda)
  y = zeros(length(x))
  check(h, ccall((:bdf_ata_mul, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}), h.ptr, entity, x, lambda, y))
  return y
end
function solve_cg2(h::Handle, entity, rhs::Matrix{Float64}, lambda; tol = NaN, maxiter = 0)
  x = zeros(size(rhs)); iters = zeros(Cint, size(rhs, 2))
  check(h, ccall((:bdf_cg_solve, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Cint, Cdouble, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cint}),
                 h.ptr, entity, rhs, size(rhs, 2), lambda, tol, maxiter, x, iters))
  return x
end

## device-resident sweeps (no host buffers): for each entity {latents, Normal-Wishart statistics, (mu, Lambda) draw}
sweep!(h::Handle, n::Integer = 1) = check(h, ccall((:bdf_sweep, LIB), Cint, (Ptr{Void}, Cint), h.ptr, n))
step_sample!(h::Handle, entity) = check(h, ccall((:bdf_step_sample, LIB), Cint, (Ptr{Void}, Cint), h.ptr, entity))
step_nw_stats!(h::Handle, entity) = check(h, ccall((:bdf_step_nw_stats, LIB), Cint, (Ptr{Void}, Cint), h.ptr, entity))
step_nw_draw!(h::Handle, entity) = check(h, ccall((:bdf_step_nw_draw, LIB), Cint, (Ptr{Void}, Cint), h.ptr, entity))
function get_hyper(h::Handle, entity, D::Int)
  mu = zeros(D); Lambda = zeros(D, D)
  check(h, ccall((:bdf_get_hyper, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, mu, Lambda))
  return mu, Lambda
end

## multi-GPU (one Julia process per GPU): exchange the 64-byte handles of ipc_export between the processes, ipc_import the peers';
## the row kernel then stores every drawn row into all replicas. stats_dev / factors_dev give device pointers for NCCL or CUDA-aware MPI.
function ipc_export(h::Handle, entity)
  buf = zeros(UInt8, 64)
  check(h, ccall((:bdf_ipc_export, LIB), Cint, (Ptr{Void}, Cint, Ptr{UInt8}), h.ptr, entity, buf))
  return buf
end
ipc_import(h::Handle, entity, peer_rank::Integer, handle::Vector{UInt8}) =
  check(h, ccall((:bdf_ipc_import, LIB), Cint, (Ptr{Void}, Cint, Cint, Ptr{UInt8}), h.ptr, entity, peer_rank, handle))
function stats_dev(h::Handle, entity)
  p = Ref{Ptr{Void}}(C_NULL); n = Ref{Int64}(0)
  check(h, ccall((:bdf_stats_dev, LIB), Cint, (Ptr{Void}, Cint, Ptr{Ptr{Void}}, Ptr{Int64}), h.ptr, entity, p, n))
  return p[], n[]
end


## ---- test set and posterior accumulators on the device (src/macau.jl:142-200) ----------------------------------------------------
## setTest! / assignToTest!: ids ntest x K Int64 (1-based), values; test_F = rel.test_F (Matrix) when the relation has features
function set_test(h::Handle, rel, ids::Matrix{Int64}, values::Vector{Float64}, class_cut; test_F = C_NULL)
  check(h, ccall((:bdf_set_test, LIB), Cint, (Ptr{Void}, Cint, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                 h.ptr, rel, size(ids, 1), ids, values, test_F, class_cut))
end
test_reset(h::Handle, rel) = check(h, ccall((:bdf_test_reset, LIB), Cint, (Ptr{Void}, Cint), h.ptr, rel))
## one iteration of the test-set bookkeeping. clamp = Float64[] -> not clamped. The raw sums of THIS handle's test entries:
## [sum sq. error of the running average, sum sq. error of this sample, number of correctly classified, ntest, posterior samples so far]
function predict_accumulate_sums(h::Handle, rel, posterior::Bool, clamp::Vector{Float64})
  out = zeros(5)
  lo, hi = isempty(clamp) ? (NaN, NaN) : (clamp[1], clamp[2])
  check(h, ccall((:bdf_predict_accumulate, LIB), Cint, (Ptr{Void}, Cint, Cint, Cdouble, Cdouble, Ptr{Cdouble}), h.ptr, rel, posterior ? 1 : 0, lo, hi, out))
  return out
end
## (rmse_avg, rmse, err_avg, counter_prob) of src/macau.jl:164-200 from those sums (with several ranks: from the sums added over the ranks)
test_metrics(out::Vector{Float64}) = (sqrt(out[1] / out[4]), sqrt(out[2] / out[4]), out[3] / out[4], round(Int, out[5]))
predict_accumulate(h::Handle, rel, posterior::Bool, clamp::Vector{Float64}) = test_metrics(predict_accumulate_sums(h, rel, posterior, clamp))
## probe_rat_all (unclamped), probe_stdev (sum of squares), probe_rat
function get_test_predictions(h::Handle, rel, ntest::Int)
  avg = zeros(ntest); sq = zeros(ntest); last = zeros(ntest)
  check(h, ccall((:bdf_get_test_predictions, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, rel, avg, sq, last))
  return avg, sq, last
end

## host-side reduction of the Normal-Wishart statistics over ranks (master/worker remotecalls): nw_stats on every rank, add, set_nw_stats
set_nw_stats(h::Handle, entity, N, NU::Vector{Float64}, NS::Matrix{Float64}) =
  check(h, ccall((:bdf_set_nw_stats, LIB), Cint, (Ptr{Void}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, N, NU, NS))

## ---- deferred completion: half-sweeps return once enqueued; the Normal-Wishart draw of one entity overlaps the next entity's rows -----
set_async(h::Handle, on::Bool = true) = check(h, ccall((:bdf_set_async, LIB), Cint, (Ptr{Void}, Cint), h.ptr, on ? 1 : 0))
function nw_sample_async(h::Handle, entity, mu0::Vector{Float64}, b0, Tinv::Matrix{Float64}, nu; A = C_NULL, z = C_NULL)
  check(h, ccall((:bdf_nw_sample_async, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                 h.ptr, entity, mu0, b0, Tinv, nu, A, z))
end
function nw_sample_fetch(h::Handle, entity, D::Int)
  mu = zeros(D); Lambda = zeros(D, D)
  check(h, ccall((:bdf_nw_sample_fetch, LIB), Cint, (Ptr{Void}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h.ptr, entity, mu, Lambda))
  return mu, Lambda
end

## ---- column-split beta solve over the ranks (solve_cg2, src/parallel_matrix.jl:488-507): map every peer's beta once ---------------------
function ipc_export_beta(h::Handle, entity)
  buf = zeros(UInt8, 64)
  check(h, ccall((:bdf_ipc_export_beta, LIB), Cint, (Ptr{Void}, Cint, Ptr{UInt8}), h.ptr, entity, buf))
  return buf
end
ipc_import_beta(h::Handle, entity, peer_rank::Integer, handle::Vector{UInt8}) =
  check(h, ccall((:bdf_ipc_import_beta, LIB), Cint, (Ptr{Void}, Cint, Cint, Ptr{UInt8}), h.ptr, entity, peer_rank, handle))

end # module
