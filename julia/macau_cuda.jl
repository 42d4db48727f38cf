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
                    device::Int = 0, devices::Vector{Int} = Int[], seed::Integer = 0, inject_noise::Bool = false,
                    rank::Int = 0, world::Int = 1, peers = nothing)
  # devices = [0, 1, 2, 3]: one Julia worker process per GPU, the counterpart of latent_pids (src/macau.jl:12,44-66); see macau_cuda_multi below
  if length(devices) > 1 && world == 1
    return macau_cuda_multi(data, devices; num_latent = num_latent, lambda_beta = lambda_beta, burnin = burnin, psamples = psamples, verbose = verbose,
                            full_lambda_u = full_lambda_u, reset_model = reset_model, compute_ff_size = compute_ff_size, tol = tol, clamp = clamp,
                            seed = seed, inject_noise = inject_noise)
  end
  reset_model && reset!(data, num_latent, lambda_beta = lambda_beta, compute_ff_size = compute_ff_size)
  inject_noise && world > 1 && srand(seed)               # host-drawn variates must be the same on every rank (the Philox ones are by construction)
  D   = num_latent
  h   = BDFCuda.create(D, device = device, rank = rank, world = world)
  BDFCuda.set_seed(h, seed)
  ent = Dict{Entity,Cint}()
  for en in data.entities
    ent[en] = BDFCuda.add_entity(h, en.count)
  end
  rid = Cint[]
  for r in data.relations
    f = FastIDF(r.data)                                   # ids::Matrix{Int64} (nnz x K, 1-based), values::Vector{Float64}
    push!(rid, BDFCuda.add_relation(h, Cint[ent[en] for en in r.entities], convert(Matrix{Int64}, f.ids), convert(Vector{Float64}, f.values)))
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
  # the test set goes to the device once; the running posterior mean, the sum of squares and the clamped RMSE of src/macau.jl:164-200 are
  # accumulated there (with several ranks: this rank's share rank+1:world:ntest, the sums are added up by the master)
  mine     = (rank + 1) : world : numTest(rel)
  test_ids = convert(Matrix{Int64}, convert(Array, rel.test_vec[mine, 1:end-1]))   # the conversion FastIDF uses (src/IndexedDF.jl:53)
  if numTest(rel) > 0
    BDFCuda.set_test(h, rid[1], test_ids, convert(Vector{Float64}, convert(Array, rel.test_vec[mine, end])), rel.class_cut,
                     test_F = hasFeatures(rel) ? full(rel.test_F)[mine, :] : C_NULL)
  end
  BDFCuda.set_async(h, true)
  peers != nothing && peers.connect(h, ent)                 # multi-GPU: exchange the IPC handles (factor replicas, beta buffers)
  rmse_avg = NaN; err_avg = NaN

  for i in 1 : burnin + psamples
    # sample relation model (alpha, relation-level beta) — src/macau.jl:84-93
    for (k, r) in enumerate(data.relations)
      if r.model.alpha_sample
        sse, n = BDFCuda.train_sse(h, rid[k])
        if peers != nothing
          sse, n = peers.allreduce((sse, n))                 # every rank holds the observations of the rows it owns
        end
        r.model.alpha = BDFCuda.sample_alpha(h, rid[k], r.model.alpha_lambda0, r.model.alpha_nu0, sse, n)
      end
      if hasFeatures(r)
        r.model.beta = BDFCuda.sample_beta_rel!(h, rid[k], r.model.lambda_beta, size(r.F, 2))   # also refreshes linear_values
      end
    end
    # latent vectors and their Normal-Wishart hyper-parameters — src/macau.jl:96-134. The draw of an entity is first needed by ITS next
    # half-sweep, so it is started asynchronously (side stream) and fetched after the loop: it overlaps the next entity's row kernel.
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
      if peers != nothing                                    # several GPUs: add the ranks' statistics up (master/worker mailboxes)
        N, NU, NS = peers.allreduce((N, NU, NS))
        BDFCuda.set_nw_stats(h, e, N, NU, NS)
      end
      if inject_noise
        # the reference's own variates: Bartlett factor of Wishart(nu + N, .) and the MvNormal normals (src/normal_wishart.jl:38-42)
        A = zeros(D, D)
        for a in 1:D
          A[a, a] = sqrt(rand(Chisq(nu + N - a + 1)))
          for b in 1:a-1; A[a, b] = randn(); end
        end
        BDFCuda.nw_sample_async(h, e, mj.mu0, mj.b0, Tinv, nu, A = A, z = randn(D))
      else
        BDFCuda.nw_sample_async(h, e, mj.mu0, mj.b0, Tinv, nu)
      end
    end
    for en in data.entities
      en.model.mu, en.model.Lambda = BDFCuda.nw_sample_fetch(h, ent[en], D)
    end
    # update_beta! — src/macau.jl:138-140
    for en in data.entities
      hasFeatures(en) || continue
      e = ent[en]; mj = en.model
      BDFCuda.sample_beta_device!(h, e, mj.mu, mj.Lambda, en.lambda_beta, tol)   # beta stays on the device until the end of the run
      if peers != nothing
        # column-split solve: every rank has stored its columns of beta into all replicas once its stream has drained; nobody reads beta
        # (lambda_beta below, uhat in the next iteration) before every rank has got there
        BDFCuda.synchronize(h)
        peers.allreduce(Float64[0.0])
      end
      if en.lambda_beta_sample
        en.lambda_beta = BDFCuda.sample_lambda_beta(h, e, mj.Lambda, en.nu, en.mu)
      end
    end
    BDFCuda.advance_sweep(h)

    if numTest(rel) > 0
      sums = BDFCuda.predict_accumulate_sums(h, rid[1], i > burnin, clamp)      # 40 bytes come back per iteration
      if peers != nothing                                   # every rank holds the sums of its share of the test set
        sums[1:4] = peers.allreduce(sums[1:4])
      end
      rmse_avg, rmse, err_avg, counter_prob = BDFCuda.test_metrics(sums)
      verbose && rank == 0 && @printf("%3d: RMSE=%6.4f\n", i, rmse_avg)
    elseif peers != nothing
      # end-of-iteration rendezvous: no rank may start storing the next iteration's rows into a peer's replica while that peer still reads
      # it (the exchange above is that rendezvous when there is a test set)
      BDFCuda.synchronize(h)
      peers.allreduce(Float64[0.0])
    end
  end

  for en in data.entities                                   # hand the final state back to the host model
    BDFCuda.get_factors!(h, ent[en], en.model.sample)
    hasFeatures(en) && BDFCuda.get_beta!(h, ent[en], en.model.beta)
  end
  result = Dict{AbstractString,Any}()
  result["num_latent"] = num_latent; result["burnin"] = burnin; result["psamples"] = psamples
  if numTest(rel) > 0
    probe_rat_all, probe_stdev, _ = BDFCuda.get_test_predictions(h, rid[1], length(mine))
    if peers != nothing                                     # put the ranks' shares (rank+1 : world : ntest) back into test-set order
      shares = peers.allgather((probe_rat_all, probe_stdev))
      probe_rat_all = zeros(numTest(rel)); probe_stdev = zeros(numTest(rel))
      for q in 1:world
        probe_rat_all[q:world:end] = shares[q][1]
        probe_stdev[q:world:end]   = shares[q][2]
      end
    end
    result["RMSE"] = rmse_avg; result["accuracy"] = err_avg
    result["ROC"] = AUC_ROC(rel.test_label, -vec(probe_rat_all))
    result["probe_rat_all"] = probe_rat_all; result["probe_stdev"] = probe_stdev
  end
  return result
end


# ---- several GPUs: one worker process per device, like the reference's latent_pids workers (src/macau.jl:44-66). Every worker runs
# macau_cuda on its own handle (rank r of `world`) and owns the rows the library deals to it; drawn rows are stored straight into every peer's
# replica by the row kernel once the 64-byte IPC handles have been exchanged (`connect`), the (N, NU, NS) statistics are added up through
# mailboxes on the master (`allreduce`; a CUDA-aware MPI or NCCL binding on BDFCuda.stats_dev does the same without the host hop).
type Peers
  rank::Int
  world::Int
  connect::Function    # (h, ent) -> exchanges ipc_export / ipc_import(_beta) with the other ranks
  allgather::Function  # payload -> the payloads of all ranks, indexed by rank + 1
  allreduce::Function  # a number, an array or a tuple of those -> the sum over the ranks, added in rank order (identical on every rank)
end

addup(a::Tuple, b::Tuple) = map(addup, a, b)
addup(a, b) = a + b

function macau_cuda_multi(data::RelationData, devices::Vector{Int}; kw...)
  world = length(devices)
  length(workers()) >= world - 1 || error("devices = $devices needs $(world - 1) worker processes (addprocs)")
  pids  = vcat(myid(), workers()[1:world-1])
  # rendezvous through RemoteChannels owned by the master: one mailbox per rank
  boxes = [RemoteChannel(() -> Channel{Any}(8 * world)) for r in 1:world]
  function peers_for(r)
    seq = [0]; stash = Any[]
    # all-to-all exchange of one payload per rank. Messages carry a sequence number: a fast rank may already post exchange k+1 while this
    # rank still collects exchange k. Nobody leaves exchange k before every rank has entered it, so it is also a barrier.
    allgather = payload -> begin
      seq[1] += 1
      for q in 1:world; q == r || put!(boxes[q], (seq[1], r, payload)); end
      parts = Array(Any, world); parts[r] = payload
      missing = world - 1
      early = Any[]
      for m in stash
        if m[1] == seq[1]; parts[m[2]] = m[3]; missing -= 1; else push!(early, m); end
      end
      empty!(stash); append!(stash, early)
      while missing > 0
        m = take!(boxes[r])
        if m[1] == seq[1]; parts[m[2]] = m[3]; missing -= 1; else push!(stash, m); end
      end
      parts
    end
    # summed in rank order on every rank: all ranks must end up with bit-identical statistics, or their hyper-parameter draws diverge
    allreduce = x -> reduce(addup, allgather(x))
    connect = (h, ent) -> begin
      mine = Dict{Any,Any}()
      for (en, e) in ent
        mine[en.name] = (BDFCuda.ipc_export(h, e), hasFeatures(en) ? BDFCuda.ipc_export_beta(h, e) : UInt8[])
      end
      everyone = allgather(mine)
      for q in 1:world
        q == r && continue
        for (en, e) in ent
          BDFCuda.ipc_import(h, e, q - 1, everyone[q][en.name][1])
          isempty(everyone[q][en.name][2]) || BDFCuda.ipc_import_beta(h, e, q - 1, everyone[q][en.name][2])
        end
      end
    end
    Peers(r - 1, world, connect, allgather, allreduce)
  end
  refs = [@spawnat pids[r] macau_cuda(data; device = devices[r], rank = r - 1, world = world, peers = peers_for(r), kw...) for r in 2:world]
  result = macau_cuda(data; device = devices[1], rank = 0, world = world, peers = peers_for(1), kw...)
  map(fetch, refs)
  return result
end
