"""Parity at benchmark scale through properties that do not need the whole oracle run: on a 1/10-scale copy of the C2 workload
(same generator as bench.py: 48k × 17.8k, 10M ratings, heavy-tailed degrees, the top item has ≈200k ratings and is split over
many CTAs) a random sample of rows — always including the heaviest ones — must match the oracle's per-row draw to 1e-10 when the
oracle is fed the very Philox normals the device used; two runs with one seed are bit-identical; the Normal-Wishart count is the
row count."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("D", [32, 100])
def test_sampled_rows_of_a_large_half_sweep_match_the_oracle(D):
    import bdf_b200
    from bench import ALPHA, N_ITEMS, synth

    n1, n2, nnz = 48_000, N_ITEMS, 10_000_000
    ids, vals = synth(n1, n2, nnz, 20161018)
    mean = float(vals.mean())
    rng = np.random.default_rng(D)
    U = [rng.standard_normal((n1, D)) * 0.3, rng.standard_normal((n2, D)) * 0.3]
    G = rng.standard_normal((D, D)) * 0.2
    Lambda, mu = G @ G.T + 3.0 * np.eye(D), rng.standard_normal(D) * 0.1
    eng = bdf_b200.Engine(D)
    eng.set_seed(99)
    ents = [eng.add_entity(n1), eng.add_entity(n2)]
    rel = eng.add_relation(ents, ids, vals)
    eng.set_relation_params(rel, ALPHA, mean)
    for mode in (1, 0):  # items first: their rows are the long, split ones
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        eng.sample_mode(ents[mode], mu, Lambda, None)
        got = eng.get_factors(ents[mode])
        eng.set_factors(ents[mode], U[mode])
        eng.sample_mode(ents[mode], mu, Lambda, None)
        assert np.array_equal(got, eng.get_factors(ents[mode]))  # same seed, same sweep counter: bit-identical
        assert np.all(np.isfinite(got))
        Z = eng.debug_row_noise(ents[mode], 0)
        deg = np.bincount(ids[:, mode] - 1, minlength=[n1, n2][mode])
        rows = np.concatenate([np.argsort(-deg)[:3], np.flatnonzero(deg == 0)[:2], rng.integers(0, len(deg), 30)])
        order = np.argsort(ids[:, mode], kind="stable")   # table order inside a row, as FastIDF keeps it
        start = np.searchsorted(ids[order, mode], np.arange(1, len(deg) + 2))
        other = 1 - mode
        worst = 0.0
        for i in rows:
            sel = order[start[i]:start[i + 1]]
            want = orc.sample_row(D, [{"U": [U[other]], "ids": [ids[sel, other]], "vals": vals[sel], "offset": mean, "alpha": ALPHA}], mu, Lambda, Z[i])
            worst = max(worst, rel_err(got[i], want))
        assert worst <= 1e-10, worst
        N, NU, NS = eng.nw_stats(ents[mode])
        assert N == len(deg) and rel_err(NU, got.sum(0)) <= 1e-11
    eng.close()


def test_edge_shapes():
    """Empty relation (every row is a prior draw), a single observation, D=1 and D=128 (the largest supported)."""
    import bdf_b200

    rng = np.random.default_rng(3)
    for D, dims, nnz in ((1, [9, 4], 30), (128, [12, 5], 40), (7, [6, 3], 0), (16, [5, 4], 1)):
        ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], 1).astype(np.int64).reshape(nnz, 2)
        vals = rng.standard_normal(nnz)
        U = [rng.standard_normal((d, D)) * 0.5 for d in dims]
        G = rng.standard_normal((D, D)) * 0.2
        Lambda, mu = G @ G.T + 2.0 * np.eye(D), rng.standard_normal(D) * 0.1
        eng = bdf_b200.Engine(D)
        ents = [eng.add_entity(d) for d in dims]
        rel = eng.add_relation(ents, ids, vals)
        mean = float(vals.mean()) if nnz else 0.0
        eng.set_relation_params(rel, 1.7, mean)
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        Z = rng.standard_normal((dims[0], D))
        eng.sample_mode(ents[0], mu, Lambda, Z)
        Uo = [u.copy() for u in U]
        orc.sample_latent_all(orc.FastIDF(ids, vals, dims), 0, Uo, 1.7, mean, mu, Lambda, Z)
        assert rel_err(eng.get_factors(ents[0]), Uo[0]) <= 1e-10, (D, nnz)
        assert rel_err(eng.predict(rel, np.array([[1, 1], [dims[0], dims[1]]])), orc.pred(np.array([[1, 1], [dims[0], dims[1]]]), [Uo[0], U[1]], mean)) <= 1e-12
        eng.close()
    with pytest.raises(bdf_b200.BDFError):
        bdf_b200.Engine(129)
