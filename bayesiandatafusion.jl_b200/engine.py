"""Thin object wrapper over the C ABI: one `Engine` = one bdf_t handle on one GPU."""
from __future__ import annotations

import ctypes as C

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

    def nw_sample(self, entity: int, mu0, b0, Tinv, nu, bartlettA=None, z=None):
        D = self.D
        mu = np.zeros(D)
        Lam = np.zeros((D, D), order="F")
        A = np.asfortranarray(bartlettA, dtype=np.float64) if bartlettA is not None else None
        zz = _f64(z) if z is not None else None
        self._ck(self.lib.bdf_nw_sample(self.h, entity, _dp(_f64(mu0)), b0, _dp(np.asfortranarray(Tinv, dtype=np.float64)), nu,
                                        _dp(A), _dp(zz), _dp(mu), _dp(Lam)))
        return mu, Lam

    def step_sample(self, entity: int):
        self._ck(self.lib.bdf_step_sample(self.h, entity))

    def step_nw_stats(self, entity: int):
        self._ck(self.lib.bdf_step_nw_stats(self.h, entity))

    def step_nw_draw(self, entity: int):
        self._ck(self.lib.bdf_step_nw_draw(self.h, entity))

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

    def debug_phase_clocks(self, entity: int):
        out = np.zeros(7)
        n = C.c_int64()
        self._ck(self.lib.bdf_debug_phase_clocks(self.h, entity, _dp(out), C.byref(n)))
        return dict(zip(["setup", "syrk", "split", "build", "factor", "solve", "total"], out)), n.value

    def predict(self, rel: int, ids):
        ids = np.asfortranarray(ids, dtype=np.int64)
        out = np.zeros(ids.shape[0])
        self._ck(self.lib.bdf_predict(self.h, rel, C.c_int64(ids.shape[0]), ids.ctypes.data_as(_lib.c_i64p), _dp(out)))
        return out
