# macau_cuda.jl — the body `macau(data; backend = :cuda)` takes in src/macau.jl once BDFCuda.jl is included: the reference's Gibbs
# loop (src/macau.jl:80-254) with every numeric step routed through libbdf_b200.so. Mirrors bayesiandatafusion.jl_b200/macau.py,
# which is the version the tests run (Julia is not installed in the build image). Julia 0.4/0.5 syntax, like the reference.
#
#   include("julia/BDFCuda.jl"); include("julia/macau_cuda.jl")
#   result = macau_cuda(data; num_latent = 32, burnin = 500, psamples = 200)
using Distributions

function macau_cuda(data::RelationData;
                    num_latent::Int = 10, lambda_beta = NaN, burnin = 500, psamples = 200, verbose::Bool = true,
                    full_lambda_u = true, reset_model = true, compute_ff_size = 6500, tol = NaN, clamp::Vector{Float64} = Float64[],
                    device::Int = 0, seed::Integer = 0, inject_noise::Bool = false)
  reset_model && reset!(data, num_latent, lambda_beta = lambda_beta, compute_ff_size = compute_ff_size)
  D   = num_latent
  h   = BDFCuda.create(D, device = device)
  BDFCuda.set_seed(h, seed)
  ent = Dict{Entity,Cint}()
  for en in data.entities
    ent[en] = BDFCuda.add_entity(h, en.count)
  end
  rid = Cint[]
  for r in data.relations
    f = FastIDF(r.data)                                   # ids::Matrix{Int64} (nnz x K, 1-based), values::Vector{Float64}
    push!(rid, BDFCuda.add_relation(h, Cint[ent[en] for en in r.entities], f.ids, f.values))
    BDFCuda.set_relation_params(h, rid[end], r.model.alpha, r.model.mean_value)
    hasFeatures(r) && BDFCuda.set_relation_features(h, rid[end], full(r.F))
  end
  for en in data.entities
    e = ent[en]
    if hasFeatures(en)
      if isa(en.F, SparseBinMatrix)
        BDFCuda.set_features_sbm(h, e, size(en.F, 1), size(en.F, 2), en.F.rows, en.F.cols)
      elseif isa(en.F, SparseMatrixCSC)
        BDFCuda.set_features_csc(h, e, en.F)
      else
        BDFCuda.set_features_dense(h, e, full(en.F))
      end
      en.use_FF && BDFCuda.compute_ff!(h, e)              # reset!: numF <= compute_ff_size  =>  solve_full instead of CG
    end
  end

  rel = data.relations[1]
  test_ids = convert(Matrix{Int64}, array(rel.test_vec[:, 1:end-1]))
  probe_rat_all = zeros(numTest(rel)); probe_stdev = zeros(numTest(rel)); counter_prob = 1

  for i in 1 : burnin + psamples
    # sample relation model (alpha, relation-level beta) — src/macau.jl:84-93
    for (k, r) in enumerate(data.relations)
      if r.model.alpha_sample
        r.model.alpha = BDFCuda.sample_alpha!(h, rid[k], r.model.alpha_lambda0, r.model.alpha_nu0)
      end
      if hasFeatures(r)
        r.model.beta = BDFCuda.sample_beta_rel!(h, rid[k], r.model.lambda_beta, size(r.F, 2))   # also refreshes linear_values
      end
    end
    # latent vectors and their Normal-Wishart hyper-parameters — src/macau.jl:96-134
    for en in data.entities
      e = ent[en]; mj = en.model
      nu = mj.nu0; Tinv = mj.WI
      if hasFeatures(en)
        BDFCuda.update_uhat!(h, e, mj.mu, mj.uhat)          # mj.uhat = F_mul_beta(en)'; mu_matrix stays on the device
        BDFCuda.sample_mode_uhat!(h, e, mj.Lambda)
        N, NU, NS = BDFCuda.nw_stats_uhat(h, e, D)
        if full_lambda_u
          nu   += size(mj.beta, 1)
          Tinv  = Tinv + BDFCuda.beta_gram(h, e, D) * en.lambda_beta
        end
      else
        BDFCuda.sample_mode!(h, e, mj.mu, mj.Lambda)        # one relation or several: the engine sums them per row
        N, NU, NS = BDFCuda.nw_stats(h, e, D)
      end
      if inject_noise
        # the reference's own variates: Bartlett factor of Wishart(nu + N, .) and the MvNormal normals (src/normal_wishart.jl:38-42)
        A = zeros(D, D)
        for a in 1:D
          A[a, a] = sqrt(rand(Chisq(nu + N - a + 1)))
          for b in 1:a-1; A[a, b] = randn(); end
        end
        mj.mu, mj.Lambda = BDFCuda.nw_sample(h, e, mj.mu0, mj.b0, Tinv, nu, A = A, z = randn(D))
      else
        mj.mu, mj.Lambda = BDFCuda.nw_sample(h, e, mj.mu0, mj.b0, Tinv, nu)
      end
    end
    # update_beta! — src/macau.jl:138-140
    for en in data.entities
      hasFeatures(en) || continue
      e = ent[en]; mj = en.model
      mj.beta, rhs = BDFCuda.sample_beta!(h, e, mj.mu, mj.Lambda, en.lambda_beta, tol, size(en.F, 2))
      if en.lambda_beta_sample
        en.lambda_beta = BDFCuda.sample_lambda_beta(h, e, mj.Lambda, en.nu, en.mu)
      end
    end
    BDFCuda.advance_sweep(h)

    probe_rat = hasFeatures(rel) ? BDFCuda.predict(h, rid[1], test_ids, full(rel.test_F)) : BDFCuda.predict(h, rid[1], test_ids)
    if i > burnin
      if i == burnin + 1
        counter_prob = 1; probe_rat_all = probe_rat; probe_stdev = probe_rat .^ 2
      else
        probe_rat_all = (counter_prob * probe_rat_all + probe_rat) / (counter_prob + 1)
        probe_stdev  += probe_rat .^ 2
        counter_prob += 1
      end
    else
      probe_rat_all = probe_rat
    end
    if verbose && numTest(rel) > 0
      cl = isempty(clamp) ? probe_rat_all : makeClamped(probe_rat_all, clamp)
      @printf("%3d: RMSE=%6.4f\n", i, sqrt(mean((array(rel.test_vec[:, end]) - cl) .^ 2)))
    end
  end

  for en in data.entities                                   # hand the final state back to the host model
    BDFCuda.get_factors!(h, ent[en], en.model.sample)
    hasFeatures(en) && BDFCuda.get_beta!(h, ent[en], en.model.beta)
  end
  result = Dict{AbstractString,Any}()
  result["num_latent"] = num_latent; result["burnin"] = burnin; result["psamples"] = psamples
  if numTest(rel) > 0
    cl = isempty(clamp) ? probe_rat_all : makeClamped(probe_rat_all, clamp)
    result["RMSE"] = sqrt(mean((array(rel.test_vec[:, end]) - cl) .^ 2))
    result["ROC"]  = AUC_ROC(rel.test_label, -vec(probe_rat_all))
  end
  return result
end
