"""Host-side mirror of the reference's data model (src/RelationData.jl, src/IndexedDF.jl): the same names, argument
meaning and error behaviour, in Python, so that the drop-in seams of `macau()` read like the Julia code. Nothing here
computes on the hot path — these objects only hold host arrays and validate shapes; the Gibbs sweep itself runs in
libbdf_b200.so.

Conventions kept from the reference: ids are 1-based; a factor matrix `model.sample` is num_latent × count in Julia —
here it is held as a (count, num_latent) C-contiguous numpy array, which is byte-identical.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


class IndexedDF:
    """Observation table of one relation — src/IndexedDF.jl:6-21. `ids` is nnz×K (1-based), `values` nnz."""

    def __init__(self, ids, values, dims=None):
        ids = np.asarray(ids, dtype=np.int64)
        if ids.ndim != 2:
            raise ValueError("ids must be a 2-D array (one column per mode)")
        values = np.asarray(values, dtype=np.float64)
        if values.shape != (ids.shape[0],):
            raise ValueError("values must have one entry per row of ids")
        if dims is None:
            dims = [int(ids[:, m].max()) for m in range(ids.shape[1])]  # IndexedDF(df), src/IndexedDF.jl:25
        dims = [int(d) for d in dims]
        for m, d in enumerate(dims):
            if ids.shape[0] and (ids[:, m].min() < 1 or ids[:, m].max() > d):
                raise IndexError(f"ids of mode {m + 1} must lie in 1..{d}")
        self.ids, self.values, self.dims = ids, values, dims

    def size(self, i=None):
        return tuple(self.dims) if i is None else self.dims[i - 1]

    def nnz(self):
        return self.ids.shape[0]

    def valueMean(self):
        return float(self.values.mean())  # src/IndexedDF.jl:27

    def getCount(self, mode: int, i: int) -> int:
        return int(np.count_nonzero(self.ids[:, mode - 1] == i))

    def removeSamples(self, samples):
        """src/IndexedDF.jl:35-38; `samples` are 1-based row numbers of the table."""
        keep = np.ones(self.nnz(), dtype=bool)
        keep[np.asarray(samples, dtype=np.int64) - 1] = False
        return IndexedDF(self.ids[keep], self.values[keep], self.dims)


@dataclass
class EntityModel:
    """src/RelationData.jl:14-40."""
    sample: np.ndarray
    mu: np.ndarray
    Lambda: np.ndarray
    beta: np.ndarray
    mu0: np.ndarray
    b0: float
    WI: np.ndarray
    nu0: float
    uhat: Optional[np.ndarray] = None


class Entity:
    """src/RelationData.jl:42-64."""

    def __init__(self, name: str, F=None, lambda_beta: float = 1.0):
        self.name = name
        self.F = F
        self.FF = None
        self.use_FF = False
        self.relations: List["Relation"] = []
        self.count = 0
        self.modes: List[int] = []
        self.modes_other: List[List[int]] = []
        self.lambda_beta = float(lambda_beta)
        self.lambda_beta_sample = True
        self.mu = 1.0   # hyper-prior for lambda_beta
        self.nu = 1e-3
        self.model: Optional[EntityModel] = None

    def hasFeatures(self) -> bool:
        return self.F is not None and self.F.shape[0] > 0 and self.F.shape[1] > 0

    def initModel(self, num_latent: int, lambda_beta: float = math.nan):
        """initModel! — src/RelationData.jl:66-90."""
        D = num_latent
        numF = self.F.shape[1] if self.hasFeatures() else 0
        self.model = EntityModel(
            sample=np.zeros((self.count, D)), mu=np.zeros(D), Lambda=5.0 * np.eye(D), beta=np.zeros((numF, D), order="F"),
            mu0=np.zeros(D), b0=2.0, WI=np.eye(D), nu0=float(D), uhat=np.zeros((self.count, D)) if numF else None,
        )
        if not math.isnan(lambda_beta):
            self.lambda_beta = float(lambda_beta)


@dataclass
class RelationModel:
    """src/RelationData.jl:107-118."""
    alpha_sample: bool = False
    alpha_nu0: float = 2.0
    alpha_lambda0: float = 1.0
    lambda_beta: float = 1.0
    alpha: float = 1.0
    beta: np.ndarray = field(default_factory=lambda: np.zeros(0))
    mean_value: float = 0.0


class Relation:
    """src/RelationData.jl:127-160."""

    def __init__(self, data, name: str, entities: Optional[List[Entity]] = None, class_cut: float = 0.0, alpha: float = 1.0):
        if hasattr(data, "tocoo"):  # Relation(data::SparseMatrixCSC, name, entities) — src/RelationData.jl:163-169 (findnz order)
            if entities is not None and len(entities) != 2:
                raise ValueError("For matrix relation the number of entities has to be 2.")
            coo = data.tocsc().tocoo()
            data = IndexedDF(np.stack([coo.row + 1, coo.col + 1], axis=1).astype(np.int64), coo.data, list(data.shape))
        elif not isinstance(data, IndexedDF):
            data = IndexedDF(*data)
        self.data = data
        self.name = name
        self.F = None        # relation-level features: one row per training observation (src/RelationData.jl:127-130)
        self.test_F = None
        self.entities: List[Entity] = list(entities) if entities else []
        K = len(data.dims)
        self.test_ids = np.zeros((0, K), dtype=np.int64)
        self.test_values = np.zeros(0)
        self.test_label = np.zeros(0, dtype=bool)
        self.class_cut = float(class_cut)
        self.model = RelationModel(alpha=float(alpha))
        if self.entities:
            if len(self.entities) != K:
                raise ValueError(f"data has {K + 1} columns but needs to have {len(self.entities) + 1} which is number of entities + 1")
            for i, en in enumerate(self.entities):
                if en.count == 0:
                    en.count = data.dims[i]
                elif en.count > data.dims[i]:
                    data.dims[i] = en.count
                elif en.count < data.dims[i]:
                    raise ValueError(f"Entity {en.name} has smaller count {en.count} than the largest id in the data {data.dims[i]}. "
                                     "Set entity.count manually before creating the relation.")

    def size(self, d=None):
        return self.data.size(d)

    def numData(self):
        return self.data.nnz()

    def numTest(self):
        return self.test_ids.shape[0]

    def hasFeatures(self):
        return self.F is not None and self.F.shape[0] > 0 and self.F.shape[1] > 0  # src/RelationData.jl:178


def assignToTest(r: Relation, ntest_or_ids, rng: Optional[np.random.Generator] = None):
    """assignToTest! — src/RelationData.jl:182-205: move observations (1-based table rows) from train to test."""
    if np.isscalar(ntest_or_ids):
        rng = rng or np.random.default_rng()
        test_id = rng.choice(r.data.nnz(), size=int(ntest_or_ids), replace=False) + 1
    else:
        test_id = np.asarray(ntest_or_ids, dtype=np.int64)
    r.test_ids = r.data.ids[test_id - 1].copy()
    r.test_values = r.data.values[test_id - 1].copy()
    if r.hasFeatures():  # src/RelationData.jl:205-210
        F = np.asarray(r.F, dtype=np.float64)
        train = np.ones(F.shape[0], dtype=bool)
        train[test_id - 1] = False
        r.test_F = F[test_id - 1]
        r.F = F[train]
    r.data = r.data.removeSamples(test_id)
    r.test_label = r.test_values < r.class_cut
    return None


def setTest(r: Relation, test_ids, test_values, test_feat=None):
    """setTest! — src/RelationData.jl:214-233."""
    test_ids = np.asarray(test_ids, dtype=np.int64)
    if r.hasFeatures() and test_feat is None:
        raise ValueError("Relation has features, please supply features with test data:\nsetTest(rel, test_df, test_features")
    if r.hasFeatures() and np.shape(test_feat)[1] != r.F.shape[1]:
        raise ValueError("The test_feat must have the same number of columns as relation.F.")
    if r.hasFeatures() and np.shape(test_feat)[0] != test_ids.shape[0]:
        raise ValueError("The test_feat must have the same number of rows as test_df.")
    if r.hasFeatures():
        r.test_F = np.asarray(test_feat, dtype=np.float64)
    if test_ids.ndim != 2 or test_ids.shape[1] != len(r.data.dims):
        raise ValueError("The number of columns in test_df must be the same as in relation.data.df.")
    r.test_ids = test_ids
    r.test_values = np.asarray(test_values, dtype=np.float64)
    r.test_label = r.test_values < r.class_cut
    return None


def setPrecision(r: Relation, precision: float):
    r.model.alpha = float(precision)  # src/RelationData.jl:178-180


class RelationData:
    """src/RelationData.jl:252-312. Constructors:
      RelationData(M, feat1=…, feat2=…, entity1=…, …)  with M a scipy.sparse matrix (2-mode relation), or
      RelationData((ids, values[, dims]), …)             for a K-mode table (the DataFrame form, :278-289)."""

    def __init__(self, Am=None, feat1=None, feat2=None, entity1="E1", entity2="E2", relation="Rel", class_cut=math.log10(200),
                 alpha=5.0, alpha_sample=False, lambda_beta=1.0, names=None):
        self.entities: List[Entity] = []
        self.relations: List[Relation] = []
        if Am is None:
            return
        if isinstance(Am, Relation):  # RelationData(r::Relation) — src/RelationData.jl:307-311
            self.addRelation(Am)
            return
        if hasattr(Am, "tocoo"):  # sparse matrix: column-major nonzero order like Julia's SparseMatrixCSC (:292-298)
            coo = Am.tocsc().tocoo()
            ids = np.stack([coo.row + 1, coo.col + 1], axis=1).astype(np.int64)
            idf = IndexedDF(ids, coo.data, list(Am.shape))
        elif isinstance(Am, IndexedDF):
            idf = Am
        else:
            idf = IndexedDF(*Am)
        K = len(idf.dims)
        r = Relation(idf, relation, class_cut=class_cut, alpha=alpha)
        r.model.alpha_sample = bool(alpha_sample)
        if K == 2:
            feats = [feat1, feat2]
            enames = list(names) if names else [entity1, entity2]  # the DataFrame form names entities after its columns (:282)
        else:
            feats = [None] * K
            enames = list(names) if names else [f"E{i + 1}" for i in range(K)]
        for m in range(K):
            en = Entity(enames[m], F=feats[m], lambda_beta=lambda_beta)
            en.count = idf.dims[m]
            en.relations.append(r)
            if en.hasFeatures() and feats[m].shape[0] != idf.dims[m]:
                which = "rows" if m == 0 else "columns"
                raise ValueError(f"Number of rows in feat{m + 1} {feats[m].shape[0]} must equal number of {which} in the relation {idf.dims[m]}")
            r.entities.append(en)
            self.entities.append(en)
        self.relations.append(r)

    def addRelation(self, r: Relation):
        """addRelation! — src/RelationData.jl:387-409."""
        if len(r.data.dims) != len(r.entities):
            raise ValueError(f"Relation has {len(r.entities)} entities but its data implies {tuple(r.data.dims)}.")
        self.relations.append(r)
        for i, en in enumerate(r.entities):
            if en.count == 0:
                en.count = r.data.dims[i]
            elif en.count != r.data.dims[i]:
                raise ValueError(f"Entity {en.name} has {en.count} instances, relation {r.name} has data for {r.data.dims[i]}.")
            if not any(e is en for e in self.entities):
                self.entities.append(en)
            if not any(x is r for x in en.relations):
                en.relations.append(r)
        return None

    def reset(self, num_latent: int, lambda_beta=math.nan, compute_ff_size=6500):
        """reset! — src/RelationData.jl:331-355."""
        for en in self.entities:
            en.initModel(num_latent, lambda_beta=lambda_beta)
            en.modes = [r.entities.index(en) + 1 for r in en.relations]
            en.modes_other = [[i + 1 for i, e2 in enumerate(r.entities) if e2 is not en] for r in en.relations]
            if en.hasFeatures():
                en.use_FF = en.F.shape[1] <= compute_ff_size
        for r in self.relations:
            r.model.mean_value = r.data.valueMean()


class SparseBinMatrix:
    """src/parallel_matrix.jl:9-24 — a 0/1 matrix held as COO index lists (Int32, 1-based, no values). m, n default to
    the largest index (:23). Passing it as `feat1=` / `Entity(F=…)` puts the feature products and the beta CG solve on
    the device."""

    def __init__(self, rows, cols, m=None, n=None):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if len(rows) != len(cols):
            raise ValueError("DimensionMismatch: length(rows) must equal length(cols)")
        self.rows, self.cols = rows, cols
        self.m = int(rows.max()) if m is None else int(m)
        self.n = int(cols.max()) if n is None else int(n)

    @property
    def shape(self):
        return (self.m, self.n)

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def __getitem__(self, key):
        """sbm[rows::Vector{Bool}, :] — src/parallel_matrix.jl:26-43: keep the flagged rows (renumbered 1..sum(rows)), list order kept."""
        rows = key[0] if isinstance(key, tuple) else key
        rows = np.asarray(rows, dtype=bool)
        if rows.shape != (self.m,):
            raise ValueError("DimensionMismatch: length(rows) must equal size(sbm,1)")
        idx = rows[self.rows - 1]
        rsum = np.cumsum(rows)
        return SparseBinMatrix(rsum[self.rows[idx] - 1].astype(np.int32), self.cols[idx], int(rows.sum()), self.n)
