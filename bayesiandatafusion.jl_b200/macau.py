"""`macau()` — the reference's driver (src/macau.jl:3-254) with the hot path swapped for libbdf_b200.so.

The loop body keeps the reference's Gauss-Seidel order (src/macau.jl:96-140): for each entity {latents → Normal-Wishart},
then for each entity {beta, lambda_beta}, then the test-set prediction and the running posterior mean. Every numeric
step is a C-ABI call (sample_latent_all2! → bdf_sample_mode, ConditionalNormalWishart → bdf_nw_stats + bdf_nw_sample,
sample_beta → bdf_sample_beta, pred → bdf_predict); this module only sequences them and keeps the host-side bookkeeping
(RMSE / AUC / result dict) the reference does in Julia. There is no CPU fallback: without the CUDA library or a GPU the
call raises.
"""
from __future__ import annotations

import math
import time
from typing import Callable, Optional

import numpy as np

from .data_reading import write_binary_matrix
from .engine import Engine
from .relation_data import RelationData


def AUC_ROC(truth, score) -> float:
    """src/ROC.jl:1-11 — area under the ROC curve by the rank statistic."""
    truth = np.asarray(truth, dtype=bool)
    score = np.asarray(score, dtype=np.float64)
    npos = int(truth.sum())
    nneg = truth.size - npos
    if npos == 0 or nneg == 0:
        return math.nan
    # average ranks over ties (1-based), vectorised: a tie group occupying sorted positions a..b gets rank (a + b)/2 + 1
    uniq, inv, cnt = np.unique(score, return_inverse=True, return_counts=True)
    last = np.cumsum(cnt)
    ranks = (last - 0.5 * (cnt - 1))[inv]
    return float((ranks[truth].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))


def makeClamped(x, clamp):
    """src/sampling.jl:99-106."""
    if clamp is None or len(clamp) == 0:
        return x
    return np.clip(x, clamp[0], clamp[1])


def bartlett_factor(rng: np.random.Generator, D: int, nu: float):
    """The random part of Distributions.jl's Wishart sampler (Bartlett): lower-triangular A with
    A[i,i] = sqrt(chi2(nu - i)) (0-based i) and A[i>j] ~ N(0,1). Drawn on the host so the Wishart stream stays with the
    caller's generator, exactly as the reference leaves it with Julia's; the device applies it (bdf_nw_sample)."""
    A = np.zeros((D, D), order="F")
    for i in range(D):
        A[i, i] = math.sqrt(rng.chisquare(nu - i))
        A[i, :i] = rng.standard_normal(i)
    return A


def macau(data: RelationData, num_latent: int = 10, lambda_beta: float = math.nan, burnin: int = 500, psamples: int = 200,
          verbose: bool = True, full_lambda_u: bool = True, reset_model: bool = True, compute_ff_size: int = 6500,
          tol: float = math.nan, output: str = "", output_beta: bool = False, output_type: str = "csv", full_prediction: bool = False,
          clamp=(), f: Optional[Callable] = None, rmse_train: bool = False,
          backend: str = "cuda", device: int = 0, devices=None, seed: int = 0, host_noise: Optional[np.random.Generator] = None,
          engine: Optional[Engine] = None):
    """Same keywords as src/macau.jl:3-22 where they apply to this path, plus the switch flag `backend` (only "cuda" exists here; the
    reference's `latent_pids` / `cg_pids` / `latent_blas_threads` select CPU workers and have no meaning on the device), `device` /
    `devices` (the GPUs to use: `devices=[0, 1, 2, 3]` shards the rows of every entity over four GPUs, the counterpart of
    `latent_pids`, src/macau.jl:12,44-66 — one worker process per extra device, see multi.py) and `seed` / `host_noise`: with
    `host_noise` (a numpy Generator) the Normal-Wishart, alpha and lambda_beta variates are drawn on the host and injected, otherwise
    every draw uses the device Philox stream keyed by `seed`.

    State lives on the device during the run: the test set is registered once and the running posterior mean, the sum of squares,
    the clamped RMSE and the accuracy of src/macau.jl:164-200 are accumulated there (bdf_predict_accumulate, 40 bytes back per
    iteration); ROC needs a sort of the test predictions and is evaluated on the host when `verbose` and at the end. The host model
    (`model.sample`, `model.beta`) is refreshed from the device before every call of `f(data)` and at the end."""
    if backend != "cuda":
        raise ValueError('backend must be "cuda": this package is the CUDA engine; the Julia path lives in the reference')
    if output_beta and not output:  # src/macau.jl:26-28
        raise ValueError("To output samples of beta ('output_beta = true') you have to set also output prefix, e.g., output = \"my_model\".")
    if output_type not in ("csv", "binary"):  # src/macau.jl:30
        raise ValueError('output_type must be either "csv" or "binary".')
    if not data.relations:
        raise ValueError("RelationData holds no relation")
    if devices is not None and len(devices) > 1:
        if engine is not None:
            raise ValueError("`engine` and `devices` are mutually exclusive")
        from .multi import macau_multi

        return macau_multi(data, list(devices), dict(
            num_latent=num_latent, lambda_beta=lambda_beta, burnin=burnin, psamples=psamples, verbose=verbose, full_lambda_u=full_lambda_u,
            reset_model=reset_model, compute_ff_size=compute_ff_size, tol=tol, output=output, output_beta=output_beta, output_type=output_type,
            full_prediction=full_prediction, clamp=clamp, f=f, rmse_train=rmse_train, seed=seed, host_noise=host_noise))
    if devices is not None and len(devices) == 1:
        device = int(devices[0])
    eng = engine or Engine(num_latent, device=device)
    try:
        return _macau_loop(data, eng, None, num_latent=num_latent, lambda_beta=lambda_beta, burnin=burnin, psamples=psamples, verbose=verbose,
                           full_lambda_u=full_lambda_u, reset_model=reset_model, compute_ff_size=compute_ff_size, tol=tol, output=output,
                           output_beta=output_beta, output_type=output_type, full_prediction=full_prediction, clamp=clamp, f=f,
                           rmse_train=rmse_train, seed=seed, host_noise=host_noise)
    finally:
        if engine is None:
            eng.close()


def _macau_loop(data, eng, comm, *, num_latent, lambda_beta, burnin, psamples, verbose, full_lambda_u, reset_model, compute_ff_size, tol,
                output, output_beta, output_type, full_prediction, clamp, f, rmse_train, seed, host_noise):
    """The Gibbs loop of src/macau.jl:80-254 over one engine. `comm` is None on one GPU; with several GPUs every rank runs this loop on
    its handle (rank/world set at bdf_create) and `comm` (multi.Comm) carries the all-reduces between the kernels — the statistics of
    ConditionalNormalWishart, the training SSE behind alpha, the test-set sums — so that every rank draws the identical hyper-parameters
    from identical inputs; rank 0 owns the printing, the dumps and the result."""
    rel = data.relations[0]  # predictions / RMSE are reported for the first relation, as in src/macau.jl:142-143
    lead = comm is None or comm.rank == 0
    say = verbose and lead  # `verbose` itself is the same on every rank: it also decides a collective (the ROC gather) below
    if say:
        print("Model setup")
    if reset_model:
        data.reset(num_latent, lambda_beta=lambda_beta, compute_ff_size=compute_ff_size)
    D = num_latent
    K = len(rel.entities)
    eng.set_seed(seed)
    if comm is not None:
        ents = [comm.add_entity(eng, en, data) for en in data.entities]
    else:
        ents = [eng.add_entity(en.count) for en in data.entities]
    eid = {id(en): e for e, en in zip(ents, data.entities)}
    r_ids = []
    for r in data.relations:
        rid = eng.add_relation([eid[id(en)] for en in r.entities], r.data.ids, r.data.values)
        eng.set_relation_params(rid, r.model.alpha, r.model.mean_value)
        if r.hasFeatures():
            if r.F.shape[0] != r.numData():
                raise ValueError("Relation.F must have one row per training observation")
            eng.set_relation_features(rid, r.F)  # temp.FF = F'F, linear_values = mean_value — reset!, src/RelationData.jl:349-353
        r_ids.append(rid)
    r_id = r_ids[0]
    for e, en in zip(ents, data.entities):
        if np.any(en.model.sample):
            eng.set_factors(e, en.model.sample)
        if en.hasFeatures():
            eng.set_features(e, en.F)
            if en.use_FF:
                eng.compute_ff(e)  # en.FF = full(At_mul_B(en.F, en.F)) — reset!, src/RelationData.jl:337-339
    if comm is not None:
        # peer mappings: drawn rows go straight into every replica; the CG of the link matrices is split by column over the ranks
        comm.connect(eng, ents, [e for e, en in zip(ents, data.entities) if en.hasFeatures() and not en.use_FF and not isinstance(en.F, np.ndarray)])

    if say:
        print("Sampling")
    ntest = rel.numTest()
    if ntest:
        # the test set goes to the device once; every rank registers its share (all ranks hold every factor row), the sums are all-reduced below
        sl = slice(None) if comm is None else slice(comm.rank, None, comm.world)
        eng.set_test(r_id, rel.test_ids[sl], rel.test_values[sl], rel.test_F[sl] if rel.hasFeatures() else None, rel.class_cut)
    eng.set_async(True)  # half-sweeps return once enqueued; a numeric failure surfaces at the next call that reads results back
    train_rat_all, train_counter = None, 0
    yhat_full = np.zeros(tuple(rel.data.dims), order="F") if full_prediction else None
    rmse_avg = roc_avg = err_avg = math.nan
    f_output = []
    tol_arg = math.nan if math.isnan(tol) else float(tol)
    stale = True  # the host model lags behind the device
    iter_seconds = []
    pending = {}  # entity -> its model: Normal-Wishart draws started with nw_sample_async and not yet fetched

    def fetch_draws(only=None):
        """(mu, Lambda) of the asynchronous draws, as late as the host needs them: a draw is first used by the SAME entity's next half-sweep,
        so it overlaps the other entities' row kernels and the test-set kernel (across iterations too)."""
        for e in ([only] if only is not None else list(pending)):
            if e in pending:
                mj = pending.pop(e)
                mj.mu, mj.Lambda = eng.nw_sample_fetch(e)

    def refresh_host_model():
        for e, en in zip(ents, data.entities):
            en.model.sample = eng.get_factors(e)
            if en.hasFeatures():
                en.model.beta = eng.get_beta(e)

    def test_predictions():
        """(probe_rat_all, probe_stdev) in test-set order on the lead rank."""
        avg, sq = eng.get_test_predictions(r_id)
        if comm is None:
            return avg, sq
        return comm.gather_strided(avg, ntest), comm.gather_strided(sq, ntest)

    for i in range(1, burnin + psamples + 1):
        time0 = time.time()
        # sample relation model (alpha) — src/macau.jl:84-88
        for rid, r in zip(r_ids, data.relations):
            if r.model.alpha_sample:
                sse, n = eng.train_sse(rid)
                if comm is not None:
                    sse, n = comm.allreduce_scalars([sse, n])
                c2 = host_noise.chisquare(r.model.alpha_nu0 + n) if host_noise is not None else math.nan
                r.model.alpha = eng.sample_alpha(rid, r.model.alpha_lambda0, r.model.alpha_nu0, sse, n, c2)
                eng.set_relation_params(rid, r.model.alpha, r.model.mean_value)
            if r.hasFeatures():  # src/macau.jl:89-92
                z1 = host_noise.standard_normal(r.numData()) if host_noise is not None else None
                z2 = host_noise.standard_normal(r.F.shape[1]) if host_noise is not None else None
                r.model.beta = eng.sample_beta_rel(rid, r.model.lambda_beta, z1, z2)
        # Sampling latent vectors — src/macau.jl:96-134 (entities in several relations: sample_user2_all!, :109-118)
        for e, en in zip(ents, data.entities):
            mj = en.model
            fetch_draws(e)
            if en.hasFeatures():
                eng.update_uhat(e, mj.mu)               # uhat = (F·beta)', mu_matrix = mu .+ uhat, on the device (:102-104)
                eng.sample_mode_uhat(e, mj.Lambda, None)
            else:
                eng.sample_mode(e, mj.mu, mj.Lambda, None)
            nu, Tinv = mj.nu0, mj.WI
            need_N = host_noise is not None
            if en.hasFeatures():
                if comm is None:
                    N, NU, NS = eng.nw_stats_uhat(e)
                else:
                    comm.nw_stats(eng, e, uhat=True)
                    N = float(en.count)
                if full_lambda_u:
                    nu = nu + mj.beta.shape[0]
                    Tinv = Tinv + eng.beta_gram(e) * en.lambda_beta
            elif comm is not None:
                comm.nw_stats(eng, e, uhat=False)       # device statistics, all-reduced over the ranks
                N = float(en.count)
            elif need_N:
                N, NU, NS = eng.nw_stats(e)
            else:
                eng.step_nw_stats(e)                    # the statistics stay on the device: the draw reads them there
                N = float(en.count)
            A = bartlett_factor(host_noise, D, nu + N) if host_noise is not None else None
            z = host_noise.standard_normal(D) if host_noise is not None else None
            # the draw is first needed by THIS entity's next half-sweep: it runs beside the next entity's row kernel
            eng.nw_sample_async(e, mj.mu0, mj.b0, Tinv, nu, A, z)
            pending[e] = mj
        # update_beta! needs this iteration's (mu, Lambda) of the entities with features on the host
        for e, en in zip(ents, data.entities):
            if en.hasFeatures():
                fetch_draws(e)
        # update_beta! — src/macau.jl:138-140, src/sampling.jl:361-370
        for e, en in zip(ents, data.entities):
            if en.hasFeatures():
                _, iters = eng.sample_beta(e, en.model.mu, en.model.Lambda, en.lambda_beta, tol_arg, want_beta=False)  # beta stays on the device
                if comm is not None:
                    comm.allreduce_scalars(iters)  # orders the ranks' stores of their beta columns before anyone reads beta
                if en.lambda_beta_sample:
                    g = float("nan")
                    if host_noise is not None:
                        g = host_noise.standard_gamma((en.nu + en.F.shape[1] * D) / 2.0)
                    en.lambda_beta, _ = eng.sample_lambda_beta(e, en.model.Lambda, en.nu, en.mu, g)
        eng.advance_sweep()
        stale = True

        posterior = i > burnin
        if ntest:
            sums = eng.predict_accumulate(r_id, posterior, clamp)   # running mean, sum of squares, clamped errors: on the device
            if comm is not None:
                sums = comm.allreduce_scalars(list(sums[:4])) + [sums[4]]
            rmse_avg = math.sqrt(sums[0] / sums[3])   # src/macau.jl:196
            err_avg = sums[2] / sums[3]                # :193-194
        if say or (callable(f) and posterior) or i == burnin + psamples:
            fetch_draws()  # the progress line / the callback / the caller read model.mu, model.Lambda
        if posterior:
            if output and lead:
                # saving latent vectors to disk — src/macau.jl:149-162 (Float32, num_latent × count as Julia holds model.sample)
                ndigits = int(math.floor(math.log10(psamples))) + 1
                nstr = str(i - burnin).rjust(ndigits, "0")
                for e, en in zip(ents, data.entities):
                    dumps = [("", eng.get_factors(e).T)]
                    if output_beta and en.hasFeatures():
                        dumps.append((".beta", eng.get_beta(e)))
                    for tag, X in dumps:
                        X32 = np.asarray(X, dtype=np.float32)
                        if output_type == "binary":
                            write_binary_matrix(f"{output}-{en.name}-{nstr}{tag}.binary", X32)
                        else:
                            np.savetxt(f"{output}-{en.name}-{nstr}{tag}.csv", X32, delimiter=",", fmt="%.9g")
            if rmse_train and lead:
                # train_rat = pred(rel) averaged over the posterior samples like probe_rat_all — src/macau.jl:164-178
                train_rat = eng.predict(r_id, rel.data.ids, rel.F if rel.hasFeatures() else None)
                if train_counter == 0:
                    train_rat_all, train_counter = train_rat, 1
                else:
                    train_rat_all = (train_counter * train_rat_all + train_rat) / (train_counter + 1)
                    train_counter += 1
            if full_prediction and lead:
                if rel.hasFeatures():
                    raise ValueError("Prediction of all elements is not possible when Relation has features.")  # src/sampling.jl:93-95
                yhat_full += eng.predict_all(r_id, tuple(rel.data.dims))  # pred_all — src/macau.jl:145-146
            if i == burnin + 1 and say:
                print("--------- Burn-in complete, averaging posterior samples ----------")
            if callable(f) and lead:
                if stale:
                    refresh_host_model()  # f(data) sees the live sample / beta, as in src/macau.jl:186-189
                    stale = False
                f_output.append(f(data))
        if comm is not None:
            comm.fence()  # the other ranks must not start storing the next iteration's rows into replicas rank 0 is still reading
        time1 = time.time()
        iter_seconds.append(time1 - time0)
        if ntest and (verbose or i == burnin + psamples):
            pa, _ = test_predictions()                   # every rank takes part in the gather
            if lead:
                roc_avg = AUC_ROC(rel.test_label, -pa)   # src/macau.jl:198, src/ROC.jl
        if say:
            print(f"{i:3d}: ROC={roc_avg:6.4f} RMSE={rmse_avg:6.4f} | " +
                  " ".join(f"{en.name[:3]}[mu:{np.linalg.norm(en.model.mu):6.2f}]" for en in data.entities) + " | " +
                  " ".join(f"{r.name[:4]}[a={r.model.alpha:2.1f}]" for r in data.relations) + f" [{time1 - time0:1.1f}s]")

    # the device holds the state during the run; hand the final sample back to the host model (model.sample, model.beta)
    if stale:
        refresh_host_model()

    result = {
        "num_latent": num_latent, "burnin": burnin, "psamples": psamples, "lambda_beta": data.entities[0].lambda_beta,
        "RMSE": rmse_avg, "accuracy": err_avg, "ROC": roc_avg, "latent_multi_threading": True,
    }
    if ntest > 0:
        probe_rat_all, probe_stdev = test_predictions()
    if ntest > 0 and lead:
        pred = makeClamped(probe_rat_all, clamp)
        if psamples >= 3:
            tmp = (probe_stdev - probe_rat_all ** 2 * psamples) / (psamples - 1)
            stdev = np.sqrt(np.maximum(tmp, 0.0))
        else:
            stdev = np.full(ntest, math.nan)
        train_count = np.zeros((ntest, K), dtype=np.int64)
        for m in range(K):
            cnt = np.bincount(rel.data.ids[:, m], minlength=rel.data.dims[m] + 1)
            train_count[:, m] = cnt[rel.test_ids[:, m]]
        result["predictions"] = {"ids": rel.test_ids.copy(), "values": rel.test_values.copy(), "pred": pred, "stdev": stdev}
        result["train_counts"] = train_count
    if full_prediction:
        result["predictions_full"] = yhat_full / psamples  # src/macau.jl:228-230
    if rmse_train and train_rat_all is not None:
        result["RMSE_train"] = float(np.sqrt(np.mean((rel.data.values - makeClamped(train_rat_all, clamp)) ** 2)))  # :222-226
    if callable(f):
        result["f_output"] = f_output
    result["gpu_launches"] = eng.launches
    result["seconds_per_iteration"] = float(np.mean(iter_seconds)) if iter_seconds else math.nan  # the "[%1.1fs]" of src/macau.jl:203-207
    return result
