"""Host-side driver logic of `bdf_b200.macau` on CPU, with the engine replaced by the oracle-backed look-alike of
tests/oracle_engine.py: loop order of src/macau.jl:84-140, posterior averaging and its bookkeeping (src/macau.jl:164-203,
222-241), sample dumps, full_prediction, several relations per entity with sampled alpha."""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bdf_b200  # noqa: E402
from bdf_b200 import data_reading as dr  # noqa: E402
from bdf_b200.relation_data import Entity, IndexedDF, Relation, assignToTest  # noqa: E402
from oracle_engine import OracleEngine  # noqa: E402


def test_bpmf_loop_order_averaging_dumps_and_full_prediction(tmp_path):
    rng = np.random.default_rng(1)
    N, M, k = 30, 20, 2
    A, B = rng.standard_normal((N, k)), rng.standard_normal((M, k))
    Y = sp.csc_matrix(np.where(rng.random((N, M)) < 0.5, A @ B.T + 0.05 * rng.standard_normal((N, M)), 0.0))
    rd = bdf_b200.RelationData(Y, class_cut=0.0, alpha=5.0)
    assignToTest(rd.relations[0], 40, rng)
    eng = OracleEngine(3)
    out = str(tmp_path / "run")
    res = bdf_b200.macau(rd, num_latent=3, burnin=8, psamples=12, verbose=False, engine=eng, host_noise=np.random.default_rng(2),
                         output=out, output_type="binary", full_prediction=True, rmse_train=True, clamp=[-10.0, 10.0])
    # Gauss-Seidel order per iteration: entity 1 {latents, stats, draw}, entity 2 {latents, stats, draw}, sweep counter
    per = [("sample", 0), ("stats", 0), ("draw", 0), ("sample", 1), ("stats", 1), ("draw", 1), ("sweep",)]
    assert eng.calls == per * 20
    base = float(np.sqrt(np.mean((rd.relations[0].test_values - rd.relations[0].model.mean_value) ** 2)))
    assert res["RMSE"] < 0.5 * base and res["RMSE_train"] < 0.5 * base
    p = res["predictions"]
    assert p["pred"].shape == (40,) and np.all(p["stdev"] >= 0) and res["train_counts"].shape == (40, 2)
    assert res["predictions_full"].shape == (N, M)
    i, j = p["ids"][0] - 1
    assert abs(res["predictions_full"][i, j] - p["pred"][0]) < 1e-9   # both are posterior means over the same 12 samples
    # 12 dumps per entity, zero-padded to the width of psamples, Float32, num_latent × count; the last one is the final sample
    for en in rd.entities:
        files = sorted(f for f in os.listdir(tmp_path) if f.startswith(f"run-{en.name}-"))
        assert files == [f"run-{en.name}-{s:02d}.binary" for s in range(1, 13)]
        S = dr.read_binary_float32(os.path.join(tmp_path, files[-1]))
        assert S.shape == (3, en.count) and np.allclose(S.T, en.model.sample, rtol=1e-6, atol=1e-6)


def test_two_relations_with_sampled_alpha_on_the_driver():
    rng = np.random.default_rng(5)
    nA, nB, nC, D0 = 40, 25, 15, 2
    A, B, Cm = (rng.standard_normal((n, D0)) for n in (nA, nB, nC))

    def table(X, Y, nnz):
        i = np.stack([rng.integers(1, X.shape[0] + 1, nnz), rng.integers(1, Y.shape[0] + 1, nnz)], 1).astype(np.int64)
        return IndexedDF(i, np.einsum("ij,ij->i", X[i[:, 0] - 1], Y[i[:, 1] - 1]) + 0.3 * rng.standard_normal(nnz), [X.shape[0], Y.shape[0]])

    a, b, c = Entity("a"), Entity("b"), Entity("c")
    r1 = Relation(table(A, B, 700), "ab", [a, b], alpha=1.0)
    r2 = Relation(table(A, Cm, 450), "ac", [a, c], alpha=1.0)
    r1.model.alpha_sample = r2.model.alpha_sample = True
    assignToTest(r1, 70, rng)
    rd = bdf_b200.RelationData()
    rd.addRelation(r1)
    rd.addRelation(r2)
    eng = OracleEngine(3)
    res = bdf_b200.macau(rd, num_latent=3, burnin=10, psamples=10, verbose=False, engine=eng, host_noise=np.random.default_rng(3))
    assert eng.calls[:2] == [("alpha", 0), ("alpha", 1)] and eng.calls[2] == ("sample", 0)   # src/macau.jl:84-88 precede the entity loop
    base = float(np.sqrt(np.mean((r1.test_values - r1.model.mean_value) ** 2)))
    assert res["RMSE"] < 0.6 * base
    assert 4.0 < r1.model.alpha < 25.0 and 4.0 < r2.model.alpha < 25.0   # planted noise precision 1/0.3² ≈ 11


def test_macau_feature_branch_on_the_driver():
    """Entity with side features: uhat before the latents, statistics of U − uhat, nu/Tinv corrections of full_lambda_u
    (src/macau.jl:120-130), beta and lambda_beta after the entity loop (:138-140), compute_ff_size choosing the FF solve."""
    rng = np.random.default_rng(9)
    N, M, numF, k = 60, 25, 6, 2
    F = rng.standard_normal((N, numF))
    Uo, Vo = F @ rng.standard_normal((numF, k)) * 0.5, rng.standard_normal((M, k))
    Y = sp.csc_matrix(np.where(rng.random((N, M)) < 0.3, Uo @ Vo.T + 0.1 * rng.standard_normal((N, M)), 0.0))
    rd = bdf_b200.RelationData(Y, feat1=F, class_cut=0.0, alpha=5.0)
    assignToTest(rd.relations[0], 60, rng)
    eng = OracleEngine(3)
    res = bdf_b200.macau(rd, num_latent=3, burnin=10, psamples=10, verbose=False, engine=eng, host_noise=np.random.default_rng(4))
    assert rd.entities[0].use_FF and eng.calls[0] == ("ff", 0)   # numF <= compute_ff_size: reset! precomputes FF
    per = [("uhat", 0), ("sample", 0), ("stats", 0), ("draw", 0), ("sample", 1), ("stats", 1), ("draw", 1), ("beta", 0, True), ("lambda_beta", 0), ("sweep",)]
    assert eng.calls[1:] == per * 20
    base = float(np.std(Y.data))
    assert res["RMSE"] < 0.6 * base and rd.entities[0].lambda_beta > 0
    assert rd.entities[0].model.beta.shape == (numF, 3) and np.any(rd.entities[0].model.beta)
    # compute_ff_size = 0 keeps the CG path: no FF is formed
    rd2 = bdf_b200.RelationData(Y, feat1=F, class_cut=0.0, alpha=5.0)
    eng2 = OracleEngine(3)
    bdf_b200.macau(rd2, num_latent=3, burnin=1, psamples=1, verbose=False, engine=eng2, compute_ff_size=0)
    assert not rd2.entities[0].use_FF and ("ff", 0) not in eng2.calls and ("beta", 0, False) in eng2.calls


def test_macau_keyword_contract_follows_the_reference():
    """src/macau.jl:24-30: output_beta without an output prefix and an unknown output_type are errors raised before any work is done;
    output_type defaults to "csv"."""
    import inspect

    import pytest

    rng = np.random.default_rng(2)
    Y = sp.csc_matrix(np.where(rng.random((12, 9)) < 0.5, rng.standard_normal((12, 9)), 0.0))
    rd = bdf_b200.RelationData(Y, class_cut=0.0)
    eng = OracleEngine(2)
    with pytest.raises(ValueError, match="output_beta"):
        bdf_b200.macau(rd, num_latent=2, burnin=1, psamples=1, verbose=False, engine=eng, output_beta=True)
    with pytest.raises(ValueError, match="output_type"):
        bdf_b200.macau(rd, num_latent=2, burnin=1, psamples=1, verbose=False, engine=eng, output="x", output_type="hdf5")
    assert eng.calls == []   # both failed before the first sampler call
    assert inspect.signature(bdf_b200.macau).parameters["output_type"].default == "csv"
    with pytest.raises(ValueError, match="backend"):
        bdf_b200.macau(rd, num_latent=2, backend="julia")


def test_f_callback_sees_the_live_model_and_rmse_train_is_averaged(tmp_path):
    """src/macau.jl:186-189: f(data) is called with the current sample after every posterior iteration; :164-178, 222-226: RMSE_train is the
    RMSE of the posterior MEAN of the training predictions, not of the last sample."""
    rng = np.random.default_rng(4)
    N, M, k = 25, 18, 2
    A, B = rng.standard_normal((N, k)), rng.standard_normal((M, k))
    Y = sp.csc_matrix(np.where(rng.random((N, M)) < 0.6, A @ B.T + 0.05 * rng.standard_normal((N, M)), 0.0))
    rd = bdf_b200.RelationData(Y, class_cut=0.0, alpha=5.0)
    assignToTest(rd.relations[0], 30, rng)
    eng = OracleEngine(3)
    seen = []

    def f(data):
        seen.append(data.entities[0].model.sample.copy())
        return float(np.linalg.norm(seen[-1]))

    res = bdf_b200.macau(rd, num_latent=3, burnin=4, psamples=6, verbose=False, engine=eng, host_noise=np.random.default_rng(1), f=f, rmse_train=True)
    assert len(seen) == 6 and len(res["f_output"]) == 6
    assert all(np.any(seen[i] != seen[i + 1]) for i in range(5))          # a fresh sample every time, not the stale initial zeros
    assert np.array_equal(seen[-1], rd.entities[0].model.sample)
    # RMSE_train of the averaged predictions is smaller than that of the last sample alone (averaging removes sampling noise)
    rel = rd.relations[0]
    last = eng.predict(0, rel.data.ids)
    rmse_last = float(np.sqrt(np.mean((rel.data.values - last) ** 2)))
    assert 0.0 < res["RMSE_train"] < rmse_last
