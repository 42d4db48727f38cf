"""Thin object wrapper over the C ABI: one `Engine` = one bdf_t handle on one GPU."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib


class BDFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbdf_b200 error {code}: {msg}")
        self.code = code


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp) if a is not None else None


def _f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["C" if order == "C" else "F", "A"])


class Engine:
    """Factor matrices cross this API as N×D C-contiguous numpy arrays — the same bytes as Julia's D×N
    column-major `model.sample`. Index arrays are 1-based like the reference's."""

    def __init__(self, num_latent: int, device: int = 0, rank: int = 0, world: int = 1):
        self.lib = _lib.load()
        self.D = int(num_latent)
        self.rank, self.world = rank, world
        self.counts = []
        self.rel_modes = []
        self.numF, self.rel_nF, self.ntest = {}, {}, {}   # feature columns per entity / per relation, registered test entries per relation
        h = C.c_void_p()
        rc = self.lib.bdf_create(C.byref(h), device, num_latent, rank, world)
        if rc:
            raise BDFError(rc, self.lib.bdf_last_error(None).decode())
        self.h = h

    # -- plumbing ------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc < 0:
            raise BDFError(rc, self.lib.bdf_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.bdf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.bdf_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_seed(self, seed: int):
        self._ck(self.lib.bdf_set_seed(self.h, C.c_uint64(seed)))

    def synchronize(self):
        self._ck(self.lib.bdf_synchronize(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.bdf_launch_count(self.h))

    @property
    def sweep_counter(self) -> int:
        return int(self.lib.bdf_sweep_counter(self.h))

    # -- model registration --------------------------------------------------------------------------------
    def add_entity(self, count: int) -> int:
        e = self._ck(self.lib.bdf_add_entity(self.h, C.c_int64(count)))
        self.counts.append(int(count))
        return e

    def add_entity_partitioned(self, count: int, rank_of_row) -> int:
        """Entity with an explicit shard map (rank of every row, identical on all ranks) instead of the cyclic deal."""
        r = np.ascontiguousarray(rank_of_row, dtype=np.int32)
        if r.shape != (count,):
            raise ValueError("rank_of_row must have one entry per row")
        e = self._ck(self.lib.bdf_add_entity_partitioned(self.h, C.c_int64(count), r.ctypes.data_as(_lib.c_i32p)))
        self.counts.append(int(count))
        return e

    def add_relation(self, entities, ids, vals) -> int:
        ids = np.asfortranarray(ids, dtype=np.int64)
        vals = _f64(vals)
        nnz, K = ids.shape
        if K != len(entities) or len(vals) != nnz:
            raise ValueError("ids must be nnz×K with K = len(entities); vals must have nnz entries")
        ent = (C.c_int * K)(*entities)
        r = self._ck(self.lib.bdf_add_relation(self.h, K, ent, C.c_int64(nnz), ids.ctypes.data_as(_lib.c_i64p), _dp(vals)))
        self.rel_modes.append(list(entities))
        return r

    def set_relation_params(self, rel: int, alpha: float, mean_value: float):
        self._ck(self.lib.bdf_set_relation_params(self.h, rel, alpha, mean_value))

    def set_factors(self, entity: int, U):
        U = _f64(U)
        assert U.shape == (self.counts[entity], self.D)
        self._ck(self.lib.bdf_set_factors(self.h, entity, _dp(U)))

    def get_factors(self, entity: int, out=None):
        if out is None:
            out = np.empty((self.counts[entity], self.D))
        assert out.flags.c_contiguous and out.dtype == np.float64
        self._ck(self.lib.bdf_get_factors(self.h, entity, _dp(out)))
        return out

    def factors_dev(self, entity: int):
        p = C.c_void_p()
        nper = C.c_int64()
        ld = C.c_int64()
        self._ck(self.lib.bdf_factors_dev(self.h, entity, C.byref(p), C.byref(nper), C.byref(ld)))
        return p.value, nper.value, ld.value

    def ipc_export(self, entity: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.bdf_ipc_export(self.h, entity, buf))
        return buf.raw

    def ipc_import(self, entity: int, peer_rank: int, handle: bytes):
        self._ck(self.lib.bdf_ipc_import(self.h, entity, peer_rank, handle))

    def ipc_export_beta(self, entity: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.bdf_ipc_export_beta(self.h, entity, buf))
        return buf.raw

    def ipc_import_beta(self, entity: int, peer_rank: int, handle: bytes):
        self._ck(self.lib.bdf_ipc_import_beta(self.h, entity, peer_rank, handle))

    def stats_dev(self, entity: int):
        p = C.c_void_p()
        n = C.c_int64()
        self._ck(self.lib.bdf_stats_dev(self.h, entity, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- the seams ------------------------------------------------------------------------------------------
    def sample_mode(self, entity: int, mu, Lambda, z=None):
        """sample_latent_all2! — mu: (D,) or (N, D); Lambda: (D, D); z: (N, D) injected normals or None."""
        mu = _f64(mu)
        mu_ld = 0 if mu.ndim == 1 else self.D
        Lambda = np.asfortranarray(Lambda, dtype=np.float64)
        zz = _f64(z) if z is not None else None
        self._ck(self.lib.bdf_sample_mode(self.h, entity, _dp(mu), C.c_int64(mu_ld), _dp(Lambda), _dp(zz)))

    def nw_stats(self, entity: int):
        D = self.D
        N = C.c_double()
        NU = np.zeros(D)
        NS = np.zeros((D, D), order="F")
        self._ck(self.lib.bdf_nw_stats(self.h, entity, C.cast(C.byref(N), _lib.c_dp), _dp(NU), _dp(NS)))
        return N.value, NU, NS

    def set_nw_stats(self, entity: int, N: float, NU, NS):
        """Overwrite the statistics the next nw_sample of `entity` reads (host-side reduction over ranks)."""
        self._ck(self.lib.bdf_set_nw_stats(self.h, entity, float(N), _dp(_f64(NU)), _dp(np.asfortranarray(NS, dtype=np.float64))))

    def nw_sample(self, entity: int, mu0, b0, Tinv, nu, bartlettA=None, z=None):
        D = self.D
        mu = np.zeros(D)
        Lam = np.zeros((D, D), order="F")
        A = np.asfortranarray(bartlettA, dtype=np.float64) if bartlettA is not None else None
        zz = _f64(z) if z is not None else None
        self._ck(self.lib.bdf_nw_sample(self.h, entity, _dp(_f64(mu0)), b0, _dp(np.asfortranarray(Tinv, dtype=np.float64)), nu,
                                        _dp(A), _dp(zz), _dp(mu), _dp(Lam)))
        return mu, Lam

    def set_async(self, on: bool = True):
        """Deferred completion of sample_mode / sample_mode_uhat (bdf_set_async)."""
        self._ck(self.lib.bdf_set_async(self.h, int(bool(on))))

    def nw_sample_async(self, entity: int, mu0, b0, Tinv, nu, bartlettA=None, z=None):
        """Start rand(ConditionalNormalWishart(...)) of `entity` on the side stream; pair with nw_sample_fetch."""
        A = np.asfortranarray(bartlettA, dtype=np.float64) if bartlettA is not None else None
        zz = _f64(z) if z is not None else None
        self._ck(self.lib.bdf_nw_sample_async(self.h, entity, _dp(_f64(mu0)), b0, _dp(np.asfortranarray(Tinv, dtype=np.float64)), nu, _dp(A), _dp(zz)))

    def nw_sample_fetch(self, entity: int):
        mu = np.zeros(self.D)
        Lam = np.zeros((self.D, self.D), order="F")
        self._ck(self.lib.bdf_nw_sample_fetch(self.h, entity, _dp(mu), _dp(Lam)))
        return mu, Lam

    # -- test set on the device (N1) ------------------------------------------------------------------------
    def set_test(self, rel: int, ids, vals, test_F=None, class_cut: float = 0.0):
        ids = np.asfortranarray(ids, dtype=np.int64)
        vals = _f64(vals)
        Fd = np.asfortranarray(test_F, dtype=np.float64) if test_F is not None else None
        self._ck(self.lib.bdf_set_test(self.h, rel, C.c_int64(ids.shape[0]), ids.ctypes.data_as(_lib.c_i64p), _dp(vals), _dp(Fd), float(class_cut)))
        self.ntest[rel] = int(ids.shape[0])

    def test_reset(self, rel: int):
        self._ck(self.lib.bdf_test_reset(self.h, rel))

    def predict_accumulate(self, rel: int, posterior: bool, clamp=()):
        """One iteration of src/macau.jl:143-200 on the device; returns (sse_avg, sse_sample, n_correct, ntest, counter)."""
        lo, hi = (float(clamp[0]), float(clamp[1])) if clamp is not None and len(clamp) else (math.nan, math.nan)
        out = np.zeros(5)
        self._ck(self.lib.bdf_predict_accumulate(self.h, rel, int(bool(posterior)), lo, hi, _dp(out)))
        return tuple(out)

    def get_test_predictions(self, rel: int, want_last: bool = False):
        n = self.ntest[rel]
        avg, sq = np.zeros(n), np.zeros(n)
        last = np.zeros(n) if want_last else None
        self._ck(self.lib.bdf_get_test_predictions(self.h, rel, _dp(avg), _dp(sq), _dp(last)))
        return (avg, sq, last) if want_last else (avg, sq)

    def step_sample(self, entity: int):
        self._ck(self.lib.bdf_step_sample(self.h, entity))

    def step_nw_stats(self, entity: int):
        self._ck(self.lib.bdf_step_nw_stats(self.h, entity))

    def step_nw_draw(self, entity: int):
        self._ck(self.lib.bdf_step_nw_draw(self.h, entity))


    def step_nw_draw_on(self, entity: int, cuda_stream: int):
        self._ck(self.lib.bdf_step_nw_draw_on(self.h, entity, C.c_void_p(cuda_stream)))

    def sweep(self, n: int = 1):
        self._ck(self.lib.bdf_sweep(self.h, n))

    def advance_sweep(self):
        self._ck(self.lib.bdf_advance_sweep(self.h))

    def get_hyper(self, entity: int):
        mu = np.zeros(self.D)
        Lam = np.zeros((self.D, self.D), order="F")
        self._ck(self.lib.bdf_get_hyper(self.h, entity, _dp(mu), _dp(Lam)))
        return mu, Lam

    def set_hyper(self, entity: int, mu, Lambda):
        self._ck(self.lib.bdf_set_hyper(self.h, entity, _dp(_f64(mu)), _dp(np.asfortranarray(Lambda, dtype=np.float64))))

    def debug_row_noise(self, entity: int, sweep: int):
        z = np.zeros((self.counts[entity], self.D))
        self._ck(self.lib.bdf_debug_row_noise(self.h, entity, C.c_uint64(sweep), _dp(z)))
        return z

    # -- Macau side features ------------------------------------------------------------------------------
    def set_features(self, entity: int, F):
        """F: a SparseBinMatrix-like object with .rows/.cols (1-based Int32) and .shape, or a scipy.sparse 0/1 matrix."""
        if isinstance(F, np.ndarray):
            # dense feature matrix (Julia Matrix{Float64}): column-major on the device, products by cuBLAS
            Fd = np.asfortranarray(F, dtype=np.float64)
            m, n = Fd.shape
            self._ck(self.lib.bdf_set_features_dense(self.h, entity, m, n, _dp(Fd)))
            self.numF[entity] = int(n)
            return
        if hasattr(F, "rows") and hasattr(F, "cols"):
            rows, cols, (m, n) = F.rows, F.cols, F.shape
        else:
            csc = F.tocsc()
            csc.sort_indices()
            if not np.all(csc.data == 1):
                # general sparse matrix (Julia SparseMatrixCSC fields, 1-based)
                colptr = np.ascontiguousarray(csc.indptr, dtype=np.int64) + 1
                rowval = np.ascontiguousarray(csc.indices, dtype=np.int64) + 1
                nzval = np.ascontiguousarray(csc.data, dtype=np.float64)
                m, n = F.shape
                self._ck(self.lib.bdf_set_features_csc(self.h, entity, m, n, colptr.ctypes.data_as(_lib.c_i64p), rowval.ctypes.data_as(_lib.c_i64p), _dp(nzval)))
                self.numF[entity] = int(n)
                return
            coo = csc.tocoo()
            rows, cols, (m, n) = coo.row + 1, coo.col + 1, F.shape
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if len(rows) != len(cols):
            raise ValueError("DimensionMismatch: length(rows) must equal length(cols)")  # src/parallel_matrix.jl:20
        self._ck(self.lib.bdf_set_features_sbm(self.h, entity, m, n, len(rows), rows.ctypes.data_as(_lib.c_i32p), cols.ctypes.data_as(_lib.c_i32p)))
        self.numF[entity] = int(n)

    def compute_ff(self, entity: int, want: bool = False):
        """en.FF = full(At_mul_B(en.F, en.F)), en.use_FF = true (src/RelationData.jl:337-339); returns FF when `want`."""
        n = self.numF[entity]
        out = np.zeros((n, n), order="F") if want else None
        self._ck(self.lib.bdf_compute_ff(self.h, entity, _dp(out) if want else None))
        return out

    def set_use_ff(self, entity: int, use_ff: bool):
        self._ck(self.lib.bdf_set_use_ff(self.h, entity, int(bool(use_ff))))

    def solve_full(self, entity: int, rhs, lam: float):
        """solve_full(FF, rhs, lambda) — src/sampling.jl:314-320. rhs: (numF, num_latent)."""
        rhs = np.asfortranarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs, order="F")
        self._ck(self.lib.bdf_solve_full(self.h, entity, _dp(rhs), rhs.shape[1], float(lam), _dp(x)))
        return x

    def debug_features_csr(self, entity: int, transpose: bool, nnz: int):
        nk = self.numF[entity] if transpose else self.counts[entity]
        ptr = np.zeros(nk + 1, dtype=np.int32)
        ind = np.zeros(max(nnz, 1), dtype=np.int32)
        self._ck(self.lib.bdf_debug_features_csr(self.h, entity, int(transpose), ptr.ctypes.data_as(_lib.c_i32p), ind.ctypes.data_as(_lib.c_i32p)))
        return ptr, ind[:nnz]

    def spmm(self, entity: int, X, transpose: bool = False):
        X = np.asfortranarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1, order="F")
        rout = self.numF[entity] if transpose else self.counts[entity]
        Y = np.zeros((rout, X.shape[1]), order="F")
        self._ck(self.lib.bdf_spmm(self.h, entity, int(transpose), _dp(X), X.shape[1], _dp(Y)))
        return Y

    def ata_mul(self, entity: int, x, lam: float):
        x = _f64(x)
        y = np.zeros(self.numF[entity])
        self._ck(self.lib.bdf_ata_mul(self.h, entity, _dp(x), lam, _dp(y)))
        return y

    def cg_solve(self, entity: int, rhs, lam: float, tol: float = float("nan"), maxiter: int = 0):
        rhs = np.asfortranarray(rhs, dtype=np.float64)
        x = np.zeros(rhs.shape, order="F")
        iters = np.zeros(rhs.shape[1], dtype=np.int32)
        self._ck(self.lib.bdf_cg_solve(self.h, entity, _dp(rhs), rhs.shape[1], lam, tol, maxiter, _dp(x), iters.ctypes.data_as(_lib.c_ip)))
        return x, iters

    def set_beta(self, entity: int, beta):
        self._ck(self.lib.bdf_set_beta(self.h, entity, _dp(np.asfortranarray(beta, dtype=np.float64))))

    def get_beta(self, entity: int):
        beta = np.zeros((self.numF[entity], self.D), order="F")
        self._ck(self.lib.bdf_get_beta(self.h, entity, _dp(beta)))
        return beta

    def update_uhat(self, entity: int, mu, want: bool = False):
        out = np.zeros((self.counts[entity], self.D)) if want else None
        self._ck(self.lib.bdf_update_uhat(self.h, entity, _dp(_f64(mu)), _dp(out)))
        return out

    def sample_mode_uhat(self, entity: int, Lambda, z=None):
        zz = _f64(z) if z is not None else None
        self._ck(self.lib.bdf_sample_mode_uhat(self.h, entity, _dp(np.asfortranarray(Lambda, dtype=np.float64)), _dp(zz)))

    def nw_stats_uhat(self, entity: int):
        D = self.D
        N = C.c_double()
        NU = np.zeros(D)
        NS = np.zeros((D, D), order="F")
        self._ck(self.lib.bdf_nw_stats_uhat(self.h, entity, C.cast(C.byref(N), _lib.c_dp), _dp(NU), _dp(NS)))
        return N.value, NU, NS

    def beta_gram(self, entity: int):
        B = np.zeros((self.D, self.D), order="F")
        self._ck(self.lib.bdf_beta_gram(self.h, entity, _dp(B)))
        return B

    def sample_beta(self, entity: int, mu, Lambda, lambda_beta: float, tol: float = float("nan"), E1=None, E2=None, want_rhs: bool = False,
                    want_beta: bool = True):
        """want_beta=False leaves beta on the device (required by the column-split solve of a multi-GPU run, whose columns land in this
        rank's replica only once the ranks have synchronised); returns (None, iters) then."""
        n = self.numF[entity]
        beta = np.zeros((n, self.D), order="F") if want_beta else None
        rhs = np.zeros((n, self.D), order="F") if want_rhs else None
        iters = np.zeros(self.D, dtype=np.int32)
        e1 = _f64(E1) if E1 is not None else None
        e2 = _f64(E2) if E2 is not None else None
        self._ck(self.lib.bdf_sample_beta(self.h, entity, _dp(_f64(mu)), _dp(np.asfortranarray(Lambda, dtype=np.float64)), lambda_beta, tol,
                                          _dp(e1), _dp(e2), _dp(beta), _dp(rhs), iters.ctypes.data_as(_lib.c_ip)))
        return (beta, rhs, iters) if want_rhs else (beta, iters)

    def sample_lambda_beta(self, entity: int, Lambda, nu: float, mu: float, gamma_variate: float = float("nan")):
        out = C.c_double()
        shape = C.c_double()
        self._ck(self.lib.bdf_sample_lambda_beta(self.h, entity, _dp(np.asfortranarray(Lambda, dtype=np.float64)), nu, mu, gamma_variate,
                                                 C.cast(C.byref(out), _lib.c_dp), C.cast(C.byref(shape), _lib.c_dp)))
        return out.value, shape.value

    def debug_ata_time(self, entity: int, reps: int = 20, ncols: int = 0) -> float:
        ms = C.c_double()
        self._ck(self.lib.bdf_debug_ata_time_window(self.h, entity, reps, ncols, C.cast(C.byref(ms), _lib.c_dp)))
        return ms.value

    def debug_phase_clocks(self, entity: int):
        out = np.zeros(7)
        n = C.c_int64()
        self._ck(self.lib.bdf_debug_phase_clocks(self.h, entity, _dp(out), C.byref(n)))
        return dict(zip(["setup", "syrk", "split", "build", "factor", "solve", "total"], out)), n.value

    def predict_all(self, rel: int, shape):
        """pred_all(r) — src/sampling.jl:92-97: every cell of a 2-mode relation, (N1, N2)."""
        out = np.zeros(shape, order="F")
        self._ck(self.lib.bdf_predict_all(self.h, rel, _dp(out)))
        return out

    def train_sse(self, rel: int):
        """err'err of src/macau.jl:86 over this rank's share of the training table; returns (sse, count)."""
        sse = C.c_double()
        n = C.c_int64()
        self._ck(self.lib.bdf_train_sse(self.h, rel, C.byref(sse), C.byref(n)))
        return sse.value, n.value

    def sample_alpha(self, rel: int, alpha_lambda0: float, alpha_nu0: float, sse: float, count: float, chi2: float = math.nan) -> float:
        """sample_alpha (src/sampling.jl:129-134); `chi2` is the injected chi-square(alpha_nu0 + n) variate, NaN → Philox."""
        out = C.c_double()
        self._ck(self.lib.bdf_sample_alpha(self.h, rel, alpha_lambda0, alpha_nu0, sse, float(count), chi2, C.byref(out)))
        return out.value

    def set_relation_features(self, rel: int, F):
        """Relation.F (nnz × nF, rows in the order of the table given to add_relation); FF = F'F, beta = 0 on the device."""
        Fd = np.asfortranarray(F, dtype=np.float64)
        self._ck(self.lib.bdf_set_relation_features(self.h, rel, Fd.shape[0], Fd.shape[1], _dp(Fd)))
        self.rel_nF[rel] = int(Fd.shape[1])

    def sample_beta_rel(self, rel: int, lambda_beta: float, z1=None, z2=None):
        """sample_beta_rel (src/sampling.jl:322-337) + linear_values update (src/macau.jl:90-91); returns beta (nF)."""
        beta = np.zeros(self.rel_nF[rel])
        z1 = _f64(z1) if z1 is not None else None
        z2 = _f64(z2) if z2 is not None else None
        self._ck(self.lib.bdf_sample_beta_rel(self.h, rel, float(lambda_beta), _dp(z1), _dp(z2), _dp(beta)))
        return beta

    def set_relation_beta(self, rel: int, beta):
        self._ck(self.lib.bdf_set_relation_beta(self.h, rel, _dp(_f64(beta))))

    def predict(self, rel: int, ids, test_F=None):
        if test_F is not None:
            ids = np.asfortranarray(ids, dtype=np.int64)
            Fd = np.asfortranarray(test_F, dtype=np.float64)
            out = np.zeros(ids.shape[0])
            self._ck(self.lib.bdf_predict_f(self.h, rel, C.c_int64(ids.shape[0]), ids.ctypes.data_as(_lib.c_i64p), _dp(Fd), _dp(out)))
            return out
        return self._predict_plain(rel, ids)

    def _predict_plain(self, rel: int, ids):
        ids = np.asfortranarray(ids, dtype=np.int64)
        out = np.zeros(ids.shape[0])
        self._ck(self.lib.bdf_predict(self.h, rel, C.c_int64(ids.shape[0]), ids.ctypes.data_as(_lib.c_i64p), _dp(out)))
        return out
