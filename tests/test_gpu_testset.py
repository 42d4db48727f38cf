"""N1 on the device: the test-set kernel and the posterior accumulators (`bdf_set_test`, `bdf_predict_accumulate`, `bdf_get_test_predictions`)
against the reference's host-side bookkeeping restated with the oracle's `pred` (src/macau.jl:142-200: probe_rat_all = running mean over
the posterior samples, probe_stdev = sum of squares, rmse_avg on clamped values, accuracy against class_cut); the asynchronous
Normal-Wishart draw (`bdf_nw_sample_async` / `_fetch`) against the synchronous one; `bdf_set_nw_stats`."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("dims,D", [([60, 45], 10), ([30, 20, 12], 30), ([50, 40], 100)])
def test_device_accumulators_follow_the_reference_bookkeeping(dims, D):
    import bdf_b200

    rng = np.random.default_rng(len(dims) * 100 + D)
    K, nnz, ntest = len(dims), 2000, 700
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1)
    vals = rng.standard_normal(nnz)
    tid = np.stack([rng.integers(1, d + 1, ntest) for d in dims], axis=1)
    tv = rng.standard_normal(ntest) * 0.8
    mean, cut, clamp = 0.15, 0.1, (-0.9, 1.1)
    eng = bdf_b200.Engine(D)
    ents = [eng.add_entity(d) for d in dims]
    rel = eng.add_relation(ents, ids, vals)
    eng.set_relation_params(rel, 2.0, mean)
    eng.set_test(rel, tid, tv, None, cut)
    all_o = sq_o = None
    counter = 0
    for it in range(7):
        U = [rng.standard_normal((d, D)) * 0.4 for d in dims]
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        posterior = it >= 3                                    # three "burn-in" iterations, then four posterior samples
        y = orc.pred(tid, U, mean)
        if not posterior:
            all_o = y
        elif counter == 0:
            all_o, sq_o, counter = y.copy(), y ** 2, 1         # src/macau.jl:167-171
        else:
            all_o = (counter * all_o + y) / (counter + 1)      # :174
            sq_o = sq_o + y ** 2
            counter += 1
        sse, sse_cur, ok, n, cnt = eng.predict_accumulate(rel, posterior, clamp)
        assert n == ntest and cnt == counter
        want = float(np.sum((tv - np.clip(all_o, *clamp)) ** 2))
        want_cur = float(np.sum((tv - np.clip(y, *clamp)) ** 2))
        assert abs(sse - want) <= 1e-10 * want and abs(sse_cur - want_cur) <= 1e-10 * want_cur
        assert ok == float(np.sum((tv < cut) == (all_o < cut)))
    avg, sq, last = eng.get_test_predictions(rel, want_last=True)
    assert rel_err(avg, all_o) <= 1e-12 and rel_err(sq, sq_o) <= 1e-12 and rel_err(last, y) <= 1e-12
    # no clamping (clamp = Float64[]) and a reset
    eng.test_reset(rel)
    sse, _, _, _, cnt = eng.predict_accumulate(rel, True)
    assert cnt == 1 and abs(sse - float(np.sum((tv - y) ** 2))) <= 1e-10 * sse
    # ids outside the entity are an error, not a crash
    bad = tid.copy()
    bad[0, 0] = dims[0] + 1
    with pytest.raises(bdf_b200.BDFError):
        eng.set_test(rel, bad, tv, None, cut)
    with pytest.raises(bdf_b200.BDFError):
        eng.predict_accumulate(rel, True)                      # the failed registration left no test set behind
    eng.close()


def test_test_set_with_relation_features():
    import bdf_b200

    rng = np.random.default_rng(5)
    dims, D, nnz, nF, ntest = [40, 30], 8, 900, 3, 200
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1)
    vals = rng.standard_normal(nnz)
    F = rng.standard_normal((nnz, nF))
    eng = bdf_b200.Engine(D)
    ents = [eng.add_entity(d) for d in dims]
    rel = eng.add_relation(ents, ids, vals)
    eng.set_relation_params(rel, 1.5, 0.3)
    eng.set_relation_features(rel, F)
    beta = rng.standard_normal(nF)
    eng.set_relation_beta(rel, beta)
    U = [rng.standard_normal((d, D)) * 0.5 for d in dims]
    for e, u in zip(ents, U):
        eng.set_factors(e, u)
    tid = np.stack([rng.integers(1, d + 1, ntest) for d in dims], axis=1)
    tv, tF = rng.standard_normal(ntest), rng.standard_normal((ntest, nF))
    with pytest.raises(bdf_b200.BDFError):
        eng.set_test(rel, tid, tv, None, 0.0)                  # "Relation has features, please supply features with test data"
    eng.set_test(rel, tid, tv, tF, 0.0)
    sse = eng.predict_accumulate(rel, False)[0]
    want = orc.pred(tid, U, 0.3) + tF @ beta                   # pred(r, probe_vec, F) = udot + F*beta + mean_value, src/sampling.jl:9-14
    assert abs(sse - float(np.sum((tv - want) ** 2))) <= 1e-10 * sse
    assert rel_err(eng.get_test_predictions(rel, want_last=True)[2], want) <= 1e-12
    eng.close()


@pytest.mark.parametrize("D", [10, 100])
def test_async_draw_and_host_reduced_statistics(D):
    import bdf_b200

    rng = np.random.default_rng(40 + D)
    dims, nnz = [300, 90], 5000
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1)
    vals = rng.standard_normal(nnz)
    outs = []
    for mode in ("sync", "async"):
        eng = bdf_b200.Engine(D)
        eng.set_seed(3)
        ents = [eng.add_entity(d) for d in dims]
        rel = eng.add_relation(ents, ids, vals)
        eng.set_relation_params(rel, 2.0, float(vals.mean()))
        eng.set_async(mode == "async")
        hyper = {e: (np.zeros(D), 5.0 * np.eye(D)) for e in ents}
        for _ in range(3):
            for e in ents:
                eng.sample_mode(e, hyper[e][0], hyper[e][1], None)
                eng.step_nw_stats(e)
                if mode == "sync":
                    hyper[e] = eng.nw_sample(e, np.zeros(D), 2.0, np.eye(D), float(D))
                else:
                    eng.nw_sample_async(e, np.zeros(D), 2.0, np.eye(D), float(D))
            if mode == "async":
                for e in ents:
                    hyper[e] = eng.nw_sample_fetch(e)
            eng.advance_sweep()
        outs.append(([eng.get_factors(e) for e in ents], hyper))
        if mode == "sync":
            # bdf_set_nw_stats: statistics summed on the host (two halves of the rows) give the draw of the device statistics
            e = ents[0]
            N, NU, NS = eng.nw_stats(e)
            Ue = eng.get_factors(e)
            h1, h2 = orc.nw_stats(Ue[:150]), orc.nw_stats(Ue[150:])
            A, z = orc.bartlett_factor(rng, D, D + N), rng.standard_normal(D)
            ref = eng.nw_sample(e, np.zeros(D), 2.0, np.eye(D), float(D), A, z)
            eng.set_nw_stats(e, h1[0] + h2[0], h1[1] + h2[1], h1[2] + h2[2])
            got = eng.nw_sample(e, np.zeros(D), 2.0, np.eye(D), float(D), A, z)
            assert rel_err(got[0], ref[0]) <= 1e-9 and rel_err(got[1], ref[1]) <= 1e-9
        eng.close()
    (Us, hs), (Ua, ha) = outs
    for a, b in zip(Us, Ua):
        assert np.array_equal(a, b)                            # same kernels, same Philox streams: the overlap changes nothing
    for e in hs:
        assert np.array_equal(hs[e][0], ha[e][0]) and np.array_equal(hs[e][1], ha[e][1])
