"""CPU tests of the host-side mirror of the reference interface (no GPU, no CUDA calls)."""
import math
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

import bdf_b200
from bdf_b200.shard import ShardPlan


def test_indexed_df_known_answers():
    # test/basic.jl:7-49
    X = bdf_b200.IndexedDF(np.array([[2, 1], [2, 3], [3, 4]]), [0.0, -1.0, 0.5], [4, 4])
    assert X.nnz() == 3 and X.size() == (4, 4)
    assert X.getCount(1, 2) == 2 and X.getCount(1, 1) == 0 and X.getCount(2, 2) == 0
    X2a = bdf_b200.IndexedDF(np.array([[2, 1], [2, 1], [3, 4]]), [0.4, -1, -9])
    assert X2a.size() == (3, 4)
    X3 = X.removeSamples([2])
    assert X3.nnz() == 2 and X3.size() == (4, 4) and X3.getCount(1, 2) == 1
    with pytest.raises(IndexError):
        bdf_b200.IndexedDF(np.array([[5, 1]]), [1.0], [4, 4])


def test_relation_and_entities():
    # test/basic.jl:51-90
    ids = np.array([[1, 1], [2, 3], [2, 1], [3, 4], [2, 4]])
    vals = np.array([0.4, 1.0, -1.9, 1.4, 0.85])
    e1, e2, e3 = bdf_b200.Entity("e1"), bdf_b200.Entity("e2"), bdf_b200.Entity("e3")
    r2 = bdf_b200.Relation(bdf_b200.IndexedDF(ids, vals), "r2", [e1, e2])
    assert e1.count == 3 and e2.count == 4 and r2.size() == (3, 4)
    bdf_b200.setTest(r2, np.array([[1, 3], [2, 4]]), [0.1, -0.2])
    assert r2.numTest() == 2 and r2.numData() == 5
    bdf_b200.setPrecision(r2, 1.75)
    assert r2.model.alpha == 1.75
    r3 = bdf_b200.Relation(bdf_b200.IndexedDF(np.array([[1, 5]]), [0.1]), "r3", [e2, e3])
    assert e3.count == 5 and r3.size() == (4, 5)
    with pytest.raises(ValueError):
        bdf_b200.Relation(bdf_b200.IndexedDF(np.array([[5, 4]]), [0.1]), "r4", [e1, e3])


def test_relation_data_from_sparse_and_test_split():
    rng = np.random.default_rng(0)
    Y = sp.random(15, 10, 0.3, random_state=1, format="csc")
    rd = bdf_b200.RelationData(Y, class_cut=0.5)
    rel = rd.relations[0]
    n0 = rel.numData()
    bdf_b200.assignToTest(rel, 2, rng)
    assert rel.numTest() == 2 and len(rel.test_label) == 2 and rel.numData() == n0 - 2
    assert [e.count for e in rd.entities] == [15, 10]
    rd.reset(4)
    m = rd.entities[0].model
    # initModel! — src/RelationData.jl:66-90
    assert m.sample.shape == (15, 4) and not m.sample.any() and np.array_equal(m.Lambda, 5 * np.eye(4))
    assert m.b0 == 2.0 and m.nu0 == 4 and np.array_equal(m.WI, np.eye(4))
    assert rd.entities[0].modes == [1] and rd.entities[1].modes_other == [[1]]
    assert abs(rel.model.mean_value - rel.data.values.mean()) < 1e-15
    with pytest.raises(ValueError):
        bdf_b200.RelationData(Y, feat1=sp.identity(3, format="csc"))
    # tensor form (DataFrame constructor, src/RelationData.jl:278-289)
    ids = np.stack([rng.integers(1, 6, 30), rng.integers(1, 5, 30), rng.integers(1, 3, 30)], 1)
    rdt = bdf_b200.RelationData((ids, rng.standard_normal(30), [5, 4, 2]), names=["A", "B", "C"])
    assert [e.name for e in rdt.entities] == ["A", "B", "C"] and len(rdt.relations[0].entities) == 3


def test_sparse_bin_matrix_and_errors():
    rows = np.concatenate([np.arange(1, 201), np.arange(151, 351)])
    cols = np.concatenate([np.arange(151, 351), np.arange(1, 400, 2)])
    sbm = bdf_b200.SparseBinMatrix(rows, cols)
    assert sbm.shape == (350, 399) and sbm.rows.dtype == np.int32  # test/parallel_matrix.jl:47-48
    with pytest.raises(ValueError):
        bdf_b200.SparseBinMatrix(np.append(rows, 1), cols)          # test/parallel_matrix.jl:64


def test_auc_clamp_and_binary_io():
    assert bdf_b200.AUC_ROC([True, True, False, False], [0.9, 0.8, 0.3, 0.1]) == 1.0
    assert bdf_b200.AUC_ROC([True, False, True, False], [0.9, 0.8, 0.3, 0.1]) == 0.75
    assert bdf_b200.AUC_ROC([True, False], [0.5, 0.5]) == 0.5
    assert math.isnan(bdf_b200.AUC_ROC([True, True], [0.1, 0.2]))
    from bdf_b200.macau import makeClamped

    assert list(makeClamped(np.array([0.0, 3.0, 9.0]), [1.0, 5.0])) == [1.0, 3.0, 5.0]
    # on-disk formats of src/data_reading.jl: Int64 headers, column-major payloads, 1-based Int32 indices
    from bdf_b200 import data_reading as dr

    X = np.arange(12, dtype=np.float32).reshape(3, 4)     # Julia's model.sample: 3 latents × 4 instances
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.binary")
        dr.write_binary_matrix(p, X)
        raw = open(p, "rb").read()
        assert list(np.frombuffer(raw[:16], dtype="<i8")) == [3, 4]          # nrows, ncols — src/data_reading.jl:93-99
        assert np.array_equal(np.frombuffer(raw[16:], dtype="<f4"), X.T.ravel())  # column-major payload
        assert np.array_equal(dr.read_binary_float32(p), X)
        dr.write_binary_matrix(p, X.astype(np.int32))
        assert np.array_equal(dr.read_binary_int32(p), X.astype(np.int32))
        M = sp.random(9, 7, 0.3, random_state=2, format="csc")
        dr.write_sparse_float64(p, M)
        assert (dr.read_sparse_float64(p) != M).nnz == 0
        dr.write_sparse_float32(p, M)
        r, c, v = dr.read_sparse_float32(p)
        assert np.array_equal(sp.csc_matrix((v, (r - 1, c - 1)), shape=M.shape).toarray(), M.astype(np.float32).toarray())
        dr.write_sparse_binary_matrix(p, M)
        hdr = np.fromfile(p, dtype="<i8", count=3)
        assert list(hdr) == [9, 7, M.nnz]                                    # src/data_reading.jl:122-132
        F = dr.read_sparse_binary_matrix(p)
        assert F.shape == (9, 7) and np.array_equal(sp.csc_matrix((np.ones(len(F.rows)), (F.rows - 1, F.cols - 1)), shape=F.shape).toarray(), (M != 0).toarray())


def test_macau_rejects_what_is_not_on_the_device_path():
    Y = sp.random(15, 10, 0.3, random_state=1, format="csc")
    with pytest.raises(ValueError):
        bdf_b200.macau(bdf_b200.RelationData(Y), backend="julia")
    with pytest.raises(ValueError):
        bdf_b200.macau(bdf_b200.RelationData(), burnin=1, psamples=1, verbose=False)


def test_add_relation_shares_entities():
    """addRelation! — src/RelationData.jl:387-409: entities are shared by identity, counts must agree."""
    from bdf_b200.relation_data import Entity, IndexedDF, Relation

    a, b, c = Entity("a"), Entity("b"), Entity("c")
    r1 = Relation(IndexedDF(np.array([[1, 1], [4, 3]]), np.array([1.0, 2.0]), [4, 3]), "ab", [a, b])
    r2 = Relation(IndexedDF(np.array([[2, 5], [4, 1]]), np.array([0.5, 1.5]), [4, 5]), "ac", [a, c])
    rd = bdf_b200.RelationData()
    rd.addRelation(r1)
    rd.addRelation(r2)
    assert [e.name for e in rd.entities] == ["a", "b", "c"] and len(a.relations) == 2 and a.count == 4
    rd.reset(3)
    assert a.modes == [1, 1] and a.modes_other == [[2], [2]] and c.modes == [2]
    with pytest.raises(ValueError):  # entity a has 4 instances, this table has ids up to 6
        rd.addRelation(Relation(IndexedDF(np.array([[6, 1]]), np.array([1.0]), [6, 3]), "ad", [a, Entity("d")]))


def test_shard_plan_is_the_reference_cyclic_split():
    for count, world in ((10, 1), (10, 3), (7, 8), (480000, 8)):
        plan = ShardPlan(count, world)
        rows = [plan.local_rows(r) for r in range(world)]
        # ranges i:Nprocs:N of src/sampling.jl:154 (1-based there)
        for r in range(world):
            assert np.array_equal(rows[r], np.arange(r, count, world))
            assert plan.nlocal(r) == len(rows[r])
        assert sorted(np.concatenate(rows)) == list(range(count))
        slots = plan.slot(np.arange(count))
        assert len(set(slots)) == count and slots.max() < world * plan.nper
        for r in range(world):
            assert np.array_equal(np.sort(slots[rows[r]]), np.arange(r * plan.nper, r * plan.nper + len(rows[r])))
        U = np.random.default_rng(0).standard_normal((count, 3))
        assert np.array_equal(plan.from_slots(plan.to_slots(U)), U)


def test_balanced_partition_levels_heavy_tailed_work():
    from bdf_b200.shard import balanced_partition

    rng = np.random.default_rng(0)
    n, nnz = 2000, 400_000
    deg = np.bincount(np.minimum((n * rng.random(nnz) ** 2.5).astype(np.int64), n - 1), minlength=n)
    for world in (1, 2, 8):
        r = balanced_partition(deg, world, row_cost=20.0)
        assert r.shape == (n,) and r.min() >= 0 and r.max() < world
        load = np.bincount(r, weights=deg + 20.0, minlength=world)
        assert load.max() / load.mean() < 1.01
        cyc = np.bincount(np.arange(n) % world, weights=deg + 20.0, minlength=world)
        if world == 8:
            assert cyc.max() / cyc.mean() > 1.05  # the cyclic deal is what it fixes


def test_sbm_row_subsetting_known_answers():
    """test/sbm.jl:6-32."""
    rows = np.array([1, 2, 3, 2, 3, 4, 1, 2, 3, 4])
    cols = np.array([1, 1, 1, 2, 2, 2, 3, 3, 3, 3])
    m = bdf_b200.SparseBinMatrix(rows, cols)
    assert m.size() == (4, 3)
    m2 = m[np.array([True, False, True, False]), :]
    assert m2.size() == (2, 3) and len(m2.rows) == 5
    assert list(m2.rows) == [1, 2, 2, 1, 2] and list(m2.cols) == [1, 1, 2, 3, 3]
    A = sp.random(100, 50, 0.2, random_state=3, format="csc")
    A.data[:] = 1.0
    coo = A.tocoo()
    Asbm = bdf_b200.SparseBinMatrix(coo.row + 1, coo.col + 1, 100, 50)
    z = np.zeros(100, dtype=bool)
    z[:20] = True
    z[39] = True
    z[59:80] = True
    Az = Asbm[z, :]
    Az2 = sp.coo_matrix((np.ones(len(Az.rows)), (Az.rows - 1, Az.cols - 1)), shape=(int(z.sum()), 50)).toarray()
    assert np.array_equal(A.toarray()[z], Az2)
    with pytest.raises(ValueError):
        Asbm[np.ones(5, dtype=bool)]


def test_custom_rd_jl_host_part():
    """test/custom_rd.jl:6-31,36-41 — entities with features, Relation with entities, addRelation!, RelationData(r)."""
    from bdf_b200.relation_data import Entity, Relation

    rng = np.random.default_rng(0)
    genes, pheno = Entity("genes"), Entity("pheno")
    genes.F = rng.random((100, 5))
    genes.lambda_beta = 3.0
    pheno.F = rng.random((50, 8))
    ids = np.stack([rng.integers(1, 101, 1050), rng.integers(1, 51, 1050)], 1)
    ids[0] = [100, 50]
    r = Relation((ids, rng.random(1050)), "HPO", [genes, pheno], class_cut=0.5)
    rd = bdf_b200.RelationData()
    rd.addRelation(r)
    assert r.class_cut == 0.5 and genes.count == r.size(1) and pheno.count == r.size(2)
    assert len(genes.relations) == 1 and len(pheno.relations) == 1 and len(r.entities) == 2
    assert len(rd.relations) == 1 and len(rd.entities) == 2
    r2 = Relation(sp.random(100, 50, 0.01, random_state=1, format="csc"), "HPO2", [Entity("genes2"), Entity("pheno2")])
    assert r2.size() == (100, 50) and len(r2.entities) == 2
    rd2 = bdf_b200.RelationData(r2)
    assert len(rd2.relations) == 1 and len(rd2.entities) == 2


def test_map_shard_plan_mirrors_the_partitioned_entity_layout():
    from bdf_b200.shard import MapShardPlan

    r = np.array([2, 0, 0, 1, 2, 2, 0, 2])
    plan = MapShardPlan(r, 3)
    assert plan.nper == 4 and [plan.nlocal(k) for k in range(3)] == [3, 1, 4]
    assert list(plan.slot(np.arange(8))) == [8, 0, 1, 4, 9, 10, 2, 11]   # rank·nper + position inside the shard, order kept
    assert list(plan.local_rows(2)) == [0, 4, 5, 7]
    U = np.random.default_rng(0).standard_normal((8, 3))
    S = plan.to_slots(U)
    assert S.shape == (12, 3) and np.array_equal(plan.from_slots(S), U) and np.all(S[3] == 0) and np.all(S[5:8] == 0)


def test_c5_shard_generator_partitions_the_table():
    """tools/bench_c5.py: counter-based chunks, per-rank shard tables. Every observation a rank needs for a mode (its row in that
    mode is local) is in the rank's table, in table order, and the degrees agree with the full table."""
    from bdf_b200.shard import balanced_partition
    from tools.bench_c5 import chunk, degrees, shard_table

    n1, n2, nnz, W, cs = 500, 60, 20_000, 4, 3_000
    parts = [chunk(c, n1, n2, nnz, chunk_size=cs) for c in range((nnz + cs - 1) // cs)]
    i1, i2, v = (np.concatenate([p[k] for p in parts]) for k in range(3))
    assert len(v) == nnz and i1.min() >= 1 and i1.max() <= n1 and i2.max() <= n2
    again = chunk(2, n1, n2, nnz, chunk_size=cs)
    assert np.array_equal(again[0], parts[2][0]) and np.array_equal(again[2], parts[2][2])   # pure function of (seed, chunk)
    d1, d2 = degrees(n1, n2, nnz, chunk_size=cs)
    assert np.array_equal(d1, np.bincount(i1 - 1, minlength=n1)) and np.array_equal(d2, np.bincount(i2 - 1, minlength=n2))
    o1, o2 = balanced_partition(d1, W, 8.0), balanced_partition(d2, W, 8.0)
    total = 0
    for r in range(W):
        ids, vals = shard_table(r, o1, o2, n1, n2, nnz, chunk_size=cs)
        need1, need2 = o1[i1 - 1] == r, o2[i2 - 1] == r
        mine1, mine2 = o1[ids[:, 0] - 1] == r, o2[ids[:, 1] - 1] == r
        assert np.array_equal(ids[mine1], np.stack([i1[need1], i2[need1]], 1)) and np.array_equal(vals[mine1], v[need1])
        assert np.array_equal(ids[mine2], np.stack([i1[need2], i2[need2]], 1)) and np.array_equal(vals[mine2], v[need2])
        total += len(vals)
    assert nnz <= total <= 2 * nnz


def test_balanced_partition_of_a_large_entity_is_balanced_and_deterministic():
    """Entities beyond `exact_top` rows (C5: 10M users): exact LPT for the heaviest rows, snake deal of the light tail."""
    from bdf_b200.shard import balanced_partition

    rng = np.random.default_rng(0)
    n = 400_000
    deg = np.bincount(np.minimum((n * rng.random(3_000_000) ** 2.5).astype(np.int64), n - 1), minlength=n)
    for world in (2, 8):
        r = balanced_partition(deg, world, 64.0, exact_top=50_000)
        assert r.shape == (n,) and r.min() == 0 and r.max() == world - 1
        loads = np.bincount(r, weights=deg + 64.0, minlength=world)
        assert loads.max() / loads.mean() < 1.001
        counts = np.bincount(r, minlength=world)
        assert counts.max() - counts.min() <= 0.02 * n / world
        assert np.array_equal(r, balanced_partition(deg, world, 64.0, exact_top=50_000))
    # below the threshold the exact greedy assignment is unchanged
    small = deg[:5000]
    assert np.array_equal(balanced_partition(small, 4, 1.0), balanced_partition(small, 4, 1.0, exact_top=10**9))
