"""Committed golden vectors (tests/golden/make_golden.py: scipy/LAPACK twin of src/sampling.jl:205-211) against the C
oracle (CPU) and against the CUDA path through the C ABI (GPU)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "row_draw_*.npz")))


def load(path):
    g = np.load(path)
    dims = [int(d) for d in g["dims"]]
    U = [g[f"U{m}"] for m in range(len(dims))]
    return g, dims, U


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_golden(path):
    g, dims, U = load(path)
    idf = orc.FastIDF(g["ids"], g["vals"], dims)
    for mode in range(len(dims)):
        Uw = [u.copy() for u in U]
        orc.sample_latent_all(idf, mode, Uw, float(g["alpha"]), float(g["mean"]), g["mu"], g["Lambda"], g[f"Z{mode}"])
        X = g[f"X{mode}"]
        assert np.max(np.abs(Uw[mode] - X)) <= 1e-11 * np.max(np.abs(X))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_cuda_matches_golden(path):
    import bdf_b200

    g, dims, U = load(path)
    D = U[0].shape[1]
    eng = bdf_b200.Engine(D)
    ents = [eng.add_entity(d) for d in dims]
    rel = eng.add_relation(ents, g["ids"], g["vals"])
    eng.set_relation_params(rel, float(g["alpha"]), float(g["mean"]))
    for mode in range(len(dims)):
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        eng.sample_mode(ents[mode], g["mu"], g["Lambda"], g[f"Z{mode}"])
        X = g[f"X{mode}"]
        got = eng.get_factors(ents[mode])
        assert np.max(np.abs(got - X)) <= 1e-10 * np.max(np.abs(X))
    eng.close()
