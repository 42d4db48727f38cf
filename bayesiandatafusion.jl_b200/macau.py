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
          tol: float = math.nan, output: str = "", output_beta: bool = False, output_type: str = "binary", full_prediction: bool = False,
          clamp=(), f: Optional[Callable] = None, rmse_train: bool = False,
          backend: str = "cuda", device: int = 0, seed: int = 0, host_noise: Optional[np.random.Generator] = None,
          engine: Optional[Engine] = None):
    """Same keywords as src/macau.jl:3-22 where they apply to this path, plus the switch flag `backend` (only "cuda"
    exists here; the reference's `latent_pids` / `cg_pids` / `latent_blas_threads` select CPU workers and have no
    meaning on the device) and `seed` / `host_noise`: with `host_noise` (a numpy Generator) the Normal-Wishart variates
    are drawn on the host and injected, otherwise every draw uses the device Philox stream keyed by `seed`."""
    if backend != "cuda":
        raise ValueError('backend must be "cuda": this package is the CUDA engine; the Julia path lives in the reference')
    if not data.relations:
        raise ValueError("RelationData holds no relation")
    rel = data.relations[0]  # predictions / RMSE are reported for the first relation, as in src/macau.jl:142-143
    if verbose:
        print("Model setup")
    if reset_model:
        data.reset(num_latent, lambda_beta=lambda_beta, compute_ff_size=compute_ff_size)
    D = num_latent
    K = len(rel.entities)

    eng = engine or Engine(D, device=device)
    eng.set_seed(seed)
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

    if verbose:
        print("Sampling")
    ntest = rel.numTest()
    probe_rat_all = np.zeros(ntest)
    probe_stdev = np.zeros(ntest)
    counter_prob = 1
    yhat_full = np.zeros(tuple(rel.data.dims), order="F") if full_prediction else None
    rmse_avg = roc_avg = err_avg = math.nan
    f_output = []
    if math.isnan(tol):
        tol_arg = math.nan
    else:
        tol_arg = float(tol)

    for i in range(1, burnin + psamples + 1):
        time0 = time.time()
        # sample relation model (alpha) — src/macau.jl:84-88
        for rid, r in zip(r_ids, data.relations):
            if r.model.alpha_sample:
                sse, n = eng.train_sse(rid)
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
            if en.hasFeatures():
                eng.update_uhat(e, mj.mu)               # uhat = (F·beta)', mu_matrix = mu .+ uhat, on the device (:102-104)
                eng.sample_mode_uhat(e, mj.Lambda, None)
            else:
                eng.sample_mode(e, mj.mu, mj.Lambda, None)
            nu, Tinv = mj.nu0, mj.WI
            if en.hasFeatures():
                N, NU, NS = eng.nw_stats_uhat(e)
                if full_lambda_u:
                    nu = nu + mj.beta.shape[0]
                    Tinv = Tinv + eng.beta_gram(e) * en.lambda_beta
            else:
                N, NU, NS = eng.nw_stats(e)
            if host_noise is not None:
                A = bartlett_factor(host_noise, D, nu + N)
                z = host_noise.standard_normal(D)
                mj.mu, mj.Lambda = eng.nw_sample(e, mj.mu0, mj.b0, Tinv, nu, A, z)
            else:
                mj.mu, mj.Lambda = eng.nw_sample(e, mj.mu0, mj.b0, Tinv, nu)
        # update_beta! — src/macau.jl:138-140, src/sampling.jl:361-370
        for e, en in zip(ents, data.entities):
            if en.hasFeatures():
                eng.sample_beta(e, en.model.mu, en.model.Lambda, en.lambda_beta, tol_arg)
                if en.lambda_beta_sample:
                    g = float("nan")
                    if host_noise is not None:
                        g = host_noise.standard_gamma((en.nu + en.F.shape[1] * D) / 2.0)
                    en.lambda_beta, _ = eng.sample_lambda_beta(e, en.model.Lambda, en.nu, en.mu, g)
        eng.advance_sweep()

        probe_rat = eng.predict(r_id, rel.test_ids, rel.test_F if rel.hasFeatures() else None) if ntest else np.zeros(0)
        if i > burnin:
            if output:
                # saving latent vectors to disk — src/macau.jl:149-162 (Float32, num_latent × count as Julia holds model.sample)
                if output_type not in ("binary", "csv"):
                    raise ValueError('output_type must be "binary" or "csv"')
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
            if full_prediction:
                if rel.hasFeatures():
                    raise ValueError("Prediction of all elements is not possible when Relation has features.")  # src/sampling.jl:93-95
                yhat_full += eng.predict_all(r_id, tuple(rel.data.dims))  # pred_all — src/macau.jl:145-146
            if i == burnin + 1:
                if verbose:
                    print("--------- Burn-in complete, averaging posterior samples ----------")
                counter_prob = 1
                probe_rat_all = probe_rat.copy()
                probe_stdev = probe_rat ** 2
            else:
                probe_rat_all = (counter_prob * probe_rat_all + probe_rat) / (counter_prob + 1)
                probe_stdev = probe_stdev + probe_rat ** 2
                counter_prob += 1
        else:
            probe_rat_all = probe_rat
        if callable(f) and i > burnin:
            f_output.append(f(data))
        time1 = time.time()

        haveTest = ntest > 0
        if haveTest:
            correct = rel.test_label == (probe_rat_all < rel.class_cut)
            err_avg = float(correct.mean())
            clamped_all = makeClamped(probe_rat_all, clamp)
            rmse_avg = float(np.sqrt(np.mean((rel.test_values - clamped_all) ** 2)))
            roc_avg = AUC_ROC(rel.test_label, -probe_rat_all)
        if verbose:
            print(f"{i:3d}: ROC={roc_avg:6.4f} RMSE={rmse_avg:6.4f} | " +
                  " ".join(f"{en.name[:3]}[mu:{np.linalg.norm(en.model.mu):6.2f}]" for en in data.entities) + " | " +
                  " ".join(f"{r.name[:4]}[a={r.model.alpha:2.1f}]" for r in data.relations) + f" [{time1 - time0:1.1f}s]")

    # the device holds the state during the run; hand the final sample back to the host model (model.sample)
    for e, en in zip(ents, data.entities):
        en.model.sample = eng.get_factors(e)
        if en.hasFeatures():
            en.model.beta = eng.get_beta(e)

    result = {
        "num_latent": num_latent, "burnin": burnin, "psamples": psamples, "lambda_beta": data.entities[0].lambda_beta,
        "RMSE": rmse_avg, "accuracy": err_avg, "ROC": roc_avg, "latent_multi_threading": True,
    }
    if ntest > 0:
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
    if rmse_train:
        tr = eng.predict(r_id, rel.data.ids)
        result["RMSE_train"] = float(np.sqrt(np.mean((rel.data.values - makeClamped(tr, clamp)) ** 2)))
    if callable(f):
        result["f_output"] = f_output
    result["gpu_launches"] = eng.launches
    if engine is None:
        eng.close()
    return result
