// Micro-benchmark of the row kernel's inner loop shape: per k4-step NF fragment loads from shared memory, then NT DMMAs on
// distinct accumulators with the fragment pairing of a 3x4 tile block. Variants: barrier every 4 steps or not; 1-3 CTAs/SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NF, int NT, bool BAR, bool PIPE>
__global__ void __launch_bounds__(256) k(double* out, int steps) {
  extern __shared__ double sm[];
  constexpr int S = 108;
  for (int i = threadIdx.x; i < 16 * S; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double* base = sm + (lane & 3) * S + (lane >> 2);
  double acc[NT][2];
#pragma unroll
  for (int t = 0; t < NT; t++) acc[t][0] = acc[t][1] = 0.0;
  double f[NF], g[NF];
  if (PIPE) {
#pragma unroll
    for (int r = 0; r < NF; r++) f[r] = base[8 * r];
  }
  for (int s = 0; s < steps; s++) {
    const int k4 = s & 3;
    if (!PIPE) {
#pragma unroll
      for (int r = 0; r < NF; r++) f[r] = base[k4 * 4 * S + 8 * r];
    } else {
#pragma unroll
      for (int r = 0; r < NF; r++) g[r] = base[((s + 1) & 3) * 4 * S + 8 * r];
    }
#pragma unroll
    for (int t = 0; t < NT; t++) dmma(acc[t], f[t % 3], f[3 + (t / 3) % (NF - 3)]);
    if (PIPE) {
#pragma unroll
      for (int r = 0; r < NF; r++) f[r] = g[r];
    }
    if (BAR && k4 == 3) __syncthreads();
  }
  double r = 0;
#pragma unroll
  for (int t = 0; t < NT; t++) r += acc[t][0] + acc[t][1];
  if (r == 123.456) out[0] = r;
}
template <int NF, int NT, bool BAR, bool PIPE>
void run(const char* name, int sms, int cps, double* out) {
  const int steps = 40000;
  size_t smem = 16 * 108 * 8;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<NF, NT, BAR, PIPE><<<sms * cps, 256, smem>>>(out, steps);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<NF, NT, BAR, PIPE><<<sms * cps, 256, smem>>>(out, steps);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double fl = 512.0 * NT * steps * 8.0 * sms * cps;
  printf("%-34s ctas/sm=%d: %.2f TFLOP/s\n", name, cps, fl / ms / 1e9);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* out; CK(cudaMalloc(&out, 8));
  int sms = p.multiProcessorCount;
  for (int cps = 1; cps <= 3; cps++) {
    run<7, 12, false, false>("NF7 NT12 nobar", sms, cps, out);
    run<7, 12, true, false>("NF7 NT12 bar/4", sms, cps, out);
    run<7, 12, true, true>("NF7 NT12 bar/4 sw-pipelined", sms, cps, out);
    run<7, 12, false, true>("NF7 NT12 nobar sw-pipelined", sms, cps, out);
  }
  return 0;
}
