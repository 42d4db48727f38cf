// L2 read bandwidth of this GPU: every SM streams a buffer that fits the L2 (default 48 MB of the 126 MB) over and over with 128-bit
// loads, the access pattern of a gather whose operand matrix is L2-resident. The roofline denominator of the sparse-binary gather
// kernels (C3), whose operand rows are L2 hits — MEASURED_PEAKS.json only has the HBM copy figure.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_peak tools/l2_peak.cu && tools/l2_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__global__ void __launch_bounds__(512) rd(const double2* __restrict__ a, size_t n, int reps, double* out) {
  double s = 0.0;
  for (int r = 0; r < reps; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      const double2 v = __ldcg(a + ((i + (size_t)r * 977) % n));
      s += v.x + v.y;
    }
  if (s == 1.2345) out[0] = s;
}
int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 48;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const size_t n = mb * 1024 * 1024 / 16;
  double2* a; double* out; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&out, 8)); CK(cudaMemset(a, 0, n * 16));
  const int reps = 200;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int t = 0; t < 5; t++) {
    CK(cudaEventRecord(e0));
    rd<<<p.multiProcessorCount * 4, 512>>>(a, n, reps, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double gbs = (double)n * 16 * reps / ms / 1e6;
    if (gbs > best) best = gbs;
  }
  printf("{\"gpu\":\"%s\",\"l2_mb\":%d,\"buffer_mb\":%zu,\"l2_read_gbs\":%.1f}\n", p.name, p.l2CacheSize >> 20, mb, best);
  return 0;
}
