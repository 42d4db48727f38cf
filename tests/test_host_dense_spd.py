"""The dense SPD solve behind solve_full (A9, src/sampling.jl:314-320) and the relation-feature solve (src/sampling.jl:322-334), checked
WITHOUT a GPU: csrc/dense_spd.cuh is plain CUDA C (threadIdx / blockIdx / __shared__ / __syncthreads), so tests/cpu_shim compiles the very
same kernels and launch sequence for the host — the threads of a block are std::threads, __syncthreads a barrier — and the result is
compared with LAPACK. Covers block-size edges (n = 1, < 64, = 64, 65, several blocks with a ragged last one), more right-hand sides than
one CTA takes, pitch > nrhs (the pad columns must stay untouched), the strict upper triangle never being read, and the pivot check."""
import ctypes
import pathlib
import subprocess

import numpy as np
import pytest

HERE = pathlib.Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def host_solver(tmp_path_factory):
    out = tmp_path_factory.mktemp("spd") / "dense_spd_host.so"
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(out), str(HERE / "cpu_shim" / "dense_spd_host.cpp")], check=True)
    lib = ctypes.CDLL(str(out))
    lib.spd_solve_host.restype = ctypes.c_int

    def solve(A, B, ld):
        n, nrhs = B.shape
        Ac = np.asfortranarray(A.copy())
        Ac[np.triu_indices(n, 1)] = np.nan          # only the lower triangle may be read
        Bp = np.full((n, ld), 7.25)
        Bp[:, :nrhs] = B
        info, blocks = ctypes.c_int(-1), ctypes.c_long(0)
        launches = lib.spd_solve_host(Ac.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n), Bp.ctypes.data_as(ctypes.c_void_p), ld, nrhs,
                                      ctypes.byref(info), ctypes.byref(blocks))
        return Bp, Ac, info.value, launches

    return solve


@pytest.mark.parametrize("n,nrhs,ld", [(1, 1, 1), (18, 10, 12), (29, 10, 12), (64, 5, 8), (65, 32, 32), (75, 32, 32), (130, 100, 104), (200, 1, 1), (129, 130, 132)])
def test_blocked_cholesky_solve_matches_lapack(host_solver, n, nrhs, ld):
    rng = np.random.default_rng(1000 * n + nrhs)
    G = rng.standard_normal((n, n + 3))
    A = G @ G.T + 0.5 * np.eye(n)
    B = rng.standard_normal((n, nrhs))
    X, L, info, launches = host_solver(A, B, ld)
    ref = np.linalg.solve(A, B)
    assert info == 0
    assert np.max(np.abs(X[:, :nrhs] - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(np.tril(L) - np.linalg.cholesky(A))) <= 1e-12 * np.max(np.abs(A))
    assert np.all(X[:, nrhs:] == 7.25)              # pitch > nrhs: the pad columns are not touched
    nb = (n + 63) // 64
    assert launches == nb + 2 * (nb - 1) + 2 * (nb + nb - 1)   # potf2 + (trsm, syrk) per block with rows below; 2 × (trsv + update) passes


def test_pivot_check_reports_the_first_bad_column(host_solver):
    A = np.eye(70)
    A[66, 66] = -1.0
    _, _, info, _ = host_solver(A, np.ones((70, 2)), 2)
    assert info == 67
