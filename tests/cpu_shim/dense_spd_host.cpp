// Test infrastructure: the dense SPD solve of csrc/dense_spd.cuh compiled for the host through cuda_shim.h (tests/test_host_dense_spd.py).
#include "cuda_shim.h"
#include "../../bayesiandatafusion.jl_b200/csrc/dense_spd.cuh"

extern "C" int spd_solve_host(double* A, long n, double* B, int ld, int nrhs, int* info, long* blocks) {
  *info = 0;
  shim_blocks_run = 0;
  const int launches = bdf::spd::solve(0, A, (int64_t)n, B, ld, nrhs, info);
  *blocks = shim_blocks_run;
  return launches;
}
