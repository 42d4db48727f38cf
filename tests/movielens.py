"""TEST INFRASTRUCTURE: loader of the committed MovieLens-1M fixture (tests/golden/movielens_1m.npz, generated from the
reference's data/movielens_1m.mat by tests/golden/make_movielens_fixture.py) and the C1 recipe of docs/index.md:34-60."""
import os

import numpy as np
import scipy.sparse as sp

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "movielens_1m.npz")
SPLIT_SEED = 20161017  # seeded permutation for the 500 000-entry test split


def load():
    z = np.load(FIXTURE)
    out = {}
    for key in ("X", "Fu", "Fv"):
        shape = tuple(int(x) for x in z[f"{key}_shape"])
        out[key] = sp.csc_matrix((z[f"{key}_val"].astype(np.float64), (z[f"{key}_row"].astype(np.int64), z[f"{key}_col"].astype(np.int64))), shape=shape)
    return out


def relation_data(with_features: bool, ntest: int = 500_000, alpha: float = 1.5):
    """docs/index.md:40-60: users/movies entities (F = Fu/Fv when asked), ratings relation with class_cut 2.5, `ntest` observations
    moved to the test set by a seeded permutation, precision alpha."""
    import bdf_b200
    from bdf_b200.relation_data import Entity, Relation, assignToTest, setPrecision

    d = load()
    users = Entity("users", F=d["Fu"] if with_features else None)
    movies = Entity("movies", F=d["Fv"] if with_features else None)
    ratings = Relation(d["X"], "ratings", [users, movies], class_cut=2.5)
    test_id = np.random.default_rng(SPLIT_SEED).permutation(ratings.numData())[:ntest] + 1
    assignToTest(ratings, np.sort(test_id))
    setPrecision(ratings, alpha)
    return bdf_b200.RelationData(ratings)
