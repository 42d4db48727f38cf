"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs with
injected noise. Tolerance: 1e-10 relative (max-norm) in FP64, as BASELINE.json's north_star states."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_problem(rng, dims, nnz, D, heavy=None, dup=True, scale=0.3):
    K = len(dims)
    ids = np.stack([rng.integers(1, d + 1, nnz) for d in dims], axis=1).astype(np.int64)
    # leave some rows without observations (prior draw) and add duplicates
    for m, d in enumerate(dims):
        if d > 8:
            ids[ids[:, m] == 3, m] = 4
    if heavy is not None:
        mode, row, cnt = heavy
        ids[:cnt, mode] = row
    if dup and nnz > 10:
        ids[5] = ids[4]
    vals = rng.standard_normal(nnz)
    U = [rng.standard_normal((d, D)) * scale for d in dims]
    G = rng.standard_normal((D, D)) * 0.3
    Lambda = G @ G.T + 2.0 * np.eye(D)
    mu = rng.standard_normal(D) * 0.1
    return ids, vals, U, mu, Lambda


def engine_for(dims, ids, vals, U, D, alpha, mean):
    import bdf_b200

    eng = bdf_b200.Engine(D)
    ents = [eng.add_entity(d) for d in dims]
    rel = eng.add_relation(ents, ids, vals)
    eng.set_relation_params(rel, alpha, mean)
    for e, u in zip(ents, U):
        eng.set_factors(e, u)
    return eng, ents, rel


def check_half_sweeps(dims, nnz, D, seed, alpha=2.0, heavy=None, mu_matrix=False, modes=None):
    rng = np.random.default_rng(seed)
    ids, vals, U, mu, Lambda = make_problem(rng, dims, nnz, D, heavy)
    mean = float(vals.mean())
    eng, ents, rel = engine_for(dims, ids, vals, U, D, alpha, mean)
    idf = orc.FastIDF(ids, vals, dims)
    Uo = [u.copy() for u in U]
    worst = 0.0
    for mode in (modes if modes is not None else range(len(dims))):
        Z = rng.standard_normal((dims[mode], D))
        m = mu if not mu_matrix else mu + rng.standard_normal((dims[mode], D)) * 0.2
        eng.sample_mode(ents[mode], m, Lambda, Z)
        got = eng.get_factors(ents[mode])
        orc.sample_latent_all(idf, mode, Uo, alpha, mean, m, Lambda, Z)
        e = rel_err(got, Uo[mode])
        worst = max(worst, e)
        assert e <= TOL, f"mode {mode}: rel err {e:.3e}"
        # Gauss-Seidel: the next mode must see THIS mode's fresh sample on both sides; keep them bit-identical
        eng.set_factors(ents[mode], Uo[mode])
    eng.close()
    return worst


def test_reference_fixture_shape():
    """test/parallel_latent_basic.jl: 50×10 relation, D=5, alpha=5.0 — every row of both modes."""
    check_half_sweeps([50, 10], 450, 5, seed=1, alpha=5.0)


@pytest.mark.parametrize("D", [1, 2, 7, 8, 10, 16, 24, 30, 32])
def test_small_latent_dims_warp_per_row(D):
    check_half_sweeps([97, 61], 2500, D, seed=100 + D)


@pytest.mark.parametrize("D", [33, 40, 50, 64])
def test_mid_latent_dims(D):
    check_half_sweeps([83, 47], 3000, D, seed=200 + D)


@pytest.mark.parametrize("D", [65, 96, 100, 104, 127, 128])
def test_large_latent_dims_cta_per_row(D):
    check_half_sweeps([70, 45], 4000, D, seed=300 + D)


@pytest.mark.parametrize("D", [10, 32, 64, 100])
def test_heavy_row_is_split_across_ctas(D):
    """One row with 30 000 observations (> the 12 288 split threshold): partials are parked and reduced in chunk order."""
    check_half_sweeps([40, 500], 32000, D, seed=400 + D, heavy=(0, 7, 30000), modes=[0])


@pytest.mark.parametrize("D", [4, 30, 32, 50, 100])
def test_tensor_three_modes(D):
    check_half_sweeps([31, 23, 7], 3000, D, seed=500 + D)


@pytest.mark.parametrize("D", [10, 32, 100])
def test_per_row_mean_matrix(D):
    check_half_sweeps([64, 33], 2000, D, seed=600 + D, mu_matrix=True)


def test_split_result_is_deterministic():
    rng = np.random.default_rng(7)
    dims, D = [30, 400], 32
    ids, vals, U, mu, Lambda = make_problem(rng, dims, 40000, D, heavy=(0, 2, 39000))
    eng, ents, rel = engine_for(dims, ids, vals, U, D, 1.5, 0.0)
    Z = rng.standard_normal((dims[0], D))
    outs = []
    for _ in range(3):
        eng.sample_mode(ents[0], mu, Lambda, Z)
        outs.append(eng.get_factors(ents[0]).copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    eng.close()


@pytest.mark.parametrize("D", [5, 32, 100])
def test_normal_wishart_stats_and_draw(D):
    rng = np.random.default_rng(800 + D)
    dims = [1234, 77]
    ids, vals, U, mu, Lambda = make_problem(rng, dims, 3000, D, scale=1.0)
    eng, ents, rel = engine_for(dims, ids, vals, U, D, 1.0, 0.0)
    for e, u in zip(ents, U):
        N, NU, NS = eng.nw_stats(e)
        n0, NU0, NS0 = orc.nw_stats(u)
        assert N == n0
        assert rel_err(NU, NU0) <= 1e-12 and rel_err(NS, NS0) <= 1e-12
        mu0 = rng.standard_normal(D) * 0.1
        Tinv = np.eye(D) + 0.1 * np.ones((D, D))
        mu_N, beta_N, T_N, nu_N = orc.cond_normal_wishart(n0, NU0, NS0, mu0, 2.0, Tinv, float(D))
        A = orc.bartlett_factor(rng, D, nu_N)
        z = rng.standard_normal(D)
        mu_o, Lam_o = orc.nw_rand(mu_N, beta_N, T_N, A, z)
        mu_g, Lam_g = eng.nw_sample(e, mu0, 2.0, Tinv, float(D), A, z)
        print(f"NW draw D={D} N={int(n0)}: rel err Lambda {rel_err(Lam_g, Lam_o):.2e}, mu {rel_err(mu_g, mu_o):.2e}")
        assert rel_err(Lam_g, Lam_o) <= TOL, rel_err(Lam_g, Lam_o)
        assert rel_err(mu_g, mu_o) <= TOL, rel_err(mu_g, mu_o)
        mu_h, Lam_h = eng.get_hyper(e)
        assert np.array_equal(mu_h, mu_g) and np.array_equal(Lam_h, Lam_g)
    eng.close()


@pytest.mark.parametrize("D", [10, 100])
def test_philox_mode_matches_oracle_fed_with_the_same_noise(D):
    rng = np.random.default_rng(900 + D)
    dims = [120, 80]
    ids, vals, U, mu, Lambda = make_problem(rng, dims, 5000, D)
    eng, ents, rel = engine_for(dims, ids, vals, U, D, 2.0, 0.1)
    eng.set_seed(1234)
    idf = orc.FastIDF(ids, vals, dims)
    Uo = [u.copy() for u in U]
    for mode in range(2):
        eng.sample_mode(ents[mode], mu, Lambda, None)
        Z = eng.debug_row_noise(ents[mode], eng.sweep_counter)
        assert abs(Z.mean()) < 4.0 / np.sqrt(Z.size) and abs(Z.std() - 1.0) < 4.0 / np.sqrt(2 * Z.size)   # 4 sigma
        orc.sample_latent_all(idf, mode, Uo, 2.0, 0.1, mu, Lambda, Z)
        got = eng.get_factors(ents[mode])
        assert rel_err(got, Uo[mode]) <= TOL
        eng.set_factors(ents[mode], Uo[mode])
    eng.close()


def test_predict_matches_oracle():
    rng = np.random.default_rng(11)
    for dims, D in (([40, 30], 10), ([12, 9, 5], 30)):
        ids, vals, U, mu, Lambda = make_problem(rng, dims, 500, D)
        eng, ents, rel = engine_for(dims, ids, vals, U, D, 1.0, 0.7)
        tid = np.stack([rng.integers(1, d + 1, 200) for d in dims], axis=1)
        got = eng.predict(rel, tid)
        want = orc.pred(tid, U, 0.7)
        assert rel_err(got, want) <= 1e-13
        eng.close()


def test_errors_are_reported_not_fatal():
    import bdf_b200

    eng = bdf_b200.Engine(8)
    a, b = eng.add_entity(5), eng.add_entity(4)
    with pytest.raises(bdf_b200.BDFError):
        eng.add_relation([a, b], np.array([[6, 1]]), np.array([1.0]))  # id out of range
    with pytest.raises(bdf_b200.BDFError):
        eng.add_relation([a, a], np.array([[1, 1]]), np.array([1.0]))
    r = eng.add_relation([a, b], np.array([[1, 1], [5, 4]]), np.array([1.0, 2.0]))
    with pytest.raises(bdf_b200.BDFError):
        eng.set_relation_params(r, -1.0, 0.0)
    # a non-PD precision matrix is a numeric error, and the handle stays usable
    with pytest.raises(bdf_b200.BDFError) as ei:
        eng.sample_mode(a, np.zeros(8), -np.eye(8), np.zeros((5, 8)))
    assert ei.value.code == -3
    eng.sample_mode(a, np.zeros(8), np.eye(8), np.zeros((5, 8)))
    assert np.all(np.isfinite(eng.get_factors(a)))
    eng.close()


def test_device_resident_sweeps_recover_planted_model():
    """BPMF end to end on the device (Philox noise, device Normal-Wishart): the posterior-mean test RMSE on planted
    low-rank data approaches the noise level (1/sqrt(alpha))."""
    import bdf_b200

    rng = np.random.default_rng(12)
    N1, N2, D0, D = 400, 300, 4, 8
    A, B = rng.standard_normal((N1, D0)), rng.standard_normal((N2, D0))
    nnz = 30000
    ids = np.stack([rng.integers(1, N1 + 1, nnz), rng.integers(1, N2 + 1, nnz)], axis=1)
    alpha = 4.0
    vals = np.einsum("ij,ij->i", A[ids[:, 0] - 1], B[ids[:, 1] - 1]) + rng.standard_normal(nnz) / np.sqrt(alpha)
    ntest = 3000
    tr, te = slice(ntest, None), slice(0, ntest)
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(N1), eng.add_entity(N2)
    rel = eng.add_relation([e1, e2], ids[tr], vals[tr])
    mean = float(vals[tr].mean())
    eng.set_relation_params(rel, alpha, mean)
    eng.set_seed(99)
    eng.sweep(30)
    acc = np.zeros(ntest)
    for s in range(30):
        eng.sweep(1)
        acc += eng.predict(rel, ids[te])
    rmse = float(np.sqrt(np.mean((acc / 30 - vals[te]) ** 2)))
    assert rmse < 0.75, rmse  # noise floor 0.5; mean predictor ≈ 2.1
    mu, Lam = eng.get_hyper(e1)
    assert np.all(np.isfinite(mu)) and np.all(np.linalg.eigvalsh(Lam) > 0)
    assert eng.launches > 0
    eng.close()


@pytest.mark.parametrize("partition", ["cyclic", "balanced"])
def test_three_emulated_ranks_with_shard_maps(partition):
    """Rows sharded over world=3 handles (all on this GPU): cyclic `i:Nprocs:N` (src/sampling.jl:154) and an explicit
    work-balanced shard map. Every rank samples only its rows; together they reproduce the oracle's half-sweep, for injected
    noise and for Philox noise (keyed by global row id, so the map must not matter)."""
    import bdf_b200
    from bdf_b200.shard import balanced_partition

    rng = np.random.default_rng(321)
    dims, D, W, nnz = [61, 29], 32, 3, 2500
    ids, vals, U, mu, Lambda = make_problem(rng, dims, nnz, D, heavy=(0, 5, 900))
    mean = float(vals.mean())
    maps = [balanced_partition(np.bincount(ids[:, m] - 1, minlength=d), W, 20.0) for m, d in enumerate(dims)]
    if partition == "balanced":
        assert len(set(maps[0].tolist())) == W and not np.array_equal(maps[0], np.arange(dims[0]) % W)
    engs = []
    for r in range(W):
        eng = bdf_b200.Engine(D, rank=r, world=W)
        if partition == "balanced":
            ents = [eng.add_entity_partitioned(d, mp) for d, mp in zip(dims, maps)]
        else:
            ents = [eng.add_entity(d) for d in dims]
        rel = eng.add_relation(ents, ids, vals)
        eng.set_relation_params(rel, 2.0, mean)
        for e, u in zip(ents, U):
            eng.set_factors(e, u)
        engs.append((eng, ents, rel))
    owner = maps if partition == "balanced" else [np.arange(d) % W for d in dims]
    idf = orc.FastIDF(ids, vals, dims)
    # predictions go through the same id → slot map
    tids = np.stack([rng.integers(1, d + 1, 50) for d in dims], 1)
    assert rel_err(engs[1][0].predict(engs[1][2], tids), orc.pred(tids, U, mean)) <= 1e-12
    for mode in (0, 1):
        Z = rng.standard_normal((dims[mode], D))
        got = U[mode].copy()
        for r, (eng, ents, rel) in enumerate(engs):
            eng.set_factors(ents[mode], U[mode])
            eng.sample_mode(ents[mode], mu, Lambda, Z)
            mine = owner[mode] == r
            got[mine] = eng.get_factors(ents[mode])[mine]
        Uo = [u.copy() for u in U]
        orc.sample_latent_all(idf, mode, Uo, 2.0, mean, mu, Lambda, Z)
        assert rel_err(got, Uo[mode]) <= TOL
        # Philox: the noise of a row does not depend on which rank / slot holds it
        eng0, ents0, _ = engs[0]
        Zp = eng0.debug_row_noise(ents0[mode], 0)
        got = U[mode].copy()
        for r, (eng, ents, rel) in enumerate(engs):
            eng.set_factors(ents[mode], U[mode])
            eng.sample_mode(ents[mode], mu, Lambda, None)
            mine = owner[mode] == r
            got[mine] = eng.get_factors(ents[mode])[mine]
            eng.set_factors(ents[mode], U[mode])
        Uo = [u.copy() for u in U]
        orc.sample_latent_all(idf, mode, Uo, 2.0, mean, mu, Lambda, Zp)
        assert rel_err(got, Uo[mode]) <= TOL
    for eng, _, _ in engs:
        eng.close()


def test_warp_specialised_persistent_row_kernel_matches_oracle(monkeypatch):
    """row_kernel_ws.cuh (opt-in, BDF_ROWS_WS=1): two syrk groups feeding three finalise groups per SM through shared-memory tile
    slots. Same arithmetic as the one-CTA-per-row kernel, so the same 1e-10 against the oracle — plain rows, empty rows, duplicates,
    a heavy row that is split into chunks (partials parked in global memory, two-level reduction), per-row mean matrix."""
    monkeypatch.setenv("BDF_ROWS_WS", "1")
    for D in (72, 100, 104):
        check_half_sweeps([700, 60], 40000, D, seed=300 + D, heavy=(1, 7, 21000))
    check_half_sweeps([300, 40], 9000, 100, seed=77, mu_matrix=True)
