// FP64 peak micro-benchmark for B200 (sm_100a): dependent-free DFMA loop and DMMA.8x8x4 (mma.sync f64) loop.
// Writes one JSON line to stdout. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(1024) dfma_k(double* out, int iters, double s) {
  double acc[ILP];
#pragma unroll
  for (int j=0;j<ILP;j++) acc[j] = threadIdx.x + j;
  double a = s, b = 1.0 - s;
  for (int i=0;i<iters;i++) {
#pragma unroll
    for (int j=0;j<ILP;j++) acc[j] = fma(acc[j], a, b);
  }
  double r=0;
#pragma unroll
  for (int j=0;j<ILP;j++) r += acc[j];
  if (r == 123.456) out[0] = r;
}

template<int ILP>
__global__ void __launch_bounds__(1024) dmma_k(double* out, int iters, double s) {
  double c[ILP][2];
#pragma unroll
  for (int j=0;j<ILP;j++){ c[j][0]=threadIdx.x; c[j][1]=j; }
  double a = s, b = 1.0 - s;
  for (int i=0;i<iters;i++) {
#pragma unroll
    for (int j=0;j<ILP;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double r=0;
#pragma unroll
  for (int j=0;j<ILP;j++) r += c[j][0]+c[j][1];
  if (r == 123.456) out[0] = r;
}

template<typename F>
double time_ms(F launch, int reps) {
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r=0;r<reps;r++){
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if (ms<best) best=ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 20000;
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d", p.name, sms, p.clockRate);
  // DFMA: ILP 8, threads/SM sweep
  {
    int cfgs[][2] = {{256,1},{512,1},{1024,1},{1024,2}};
    for (auto& c : cfgs) {
      int thr=c[0], bps=c[1];
      double ms = time_ms([&]{ dfma_k<8><<<sms*bps, thr>>>(out, iters, 0.5); }, 5);
      double fl = 2.0*8*iters*(double)thr*sms*bps;
      printf(",\"dfma_ilp8_t%d_b%d_tflops\":%.3f", thr, bps, fl/ms/1e9);
    }
    double ms = time_ms([&]{ dfma_k<16><<<sms*2, 512>>>(out, iters, 0.5); }, 5);
    printf(",\"dfma_ilp16_t512_b2_tflops\":%.3f", 2.0*16*iters*512.0*sms*2/ms/1e9);
  }
  // DMMA 8x8x4: 512 flop per warp instr
  {
    int cfgs[][2] = {{128,1},{256,1},{512,1},{1024,1},{1024,2}};
    for (auto& c : cfgs) {
      int thr=c[0], bps=c[1];
      double ms = time_ms([&]{ dmma_k<8><<<sms*bps, thr>>>(out, iters, 0.5); }, 5);
      double fl = 512.0*8*iters*(double)(thr/32)*sms*bps;
      printf(",\"dmma884_ilp8_t%d_b%d_tflops\":%.3f", thr, bps, fl/ms/1e9);
    }
    double ms = time_ms([&]{ dmma_k<1><<<sms, 128>>>(out, iters, 0.5); }, 5);
    printf(",\"dmma884_ilp1_t128_tflops\":%.3f", 512.0*1*iters*4.0*sms/ms/1e9);
    ms = time_ms([&]{ dmma_k<2><<<sms, 128>>>(out, iters, 0.5); }, 5);
    printf(",\"dmma884_ilp2_t128_tflops\":%.3f", 512.0*2*iters*4.0*sms/ms/1e9);
    ms = time_ms([&]{ dmma_k<4><<<sms, 128>>>(out, iters, 0.5); }, 5);
    printf(",\"dmma884_ilp4_t128_tflops\":%.3f", 512.0*4*iters*4.0*sms/ms/1e9);
    ms = time_ms([&]{ dmma_k<16><<<sms, 256>>>(out, iters, 0.5); }, 5);
    printf(",\"dmma884_ilp16_t256_tflops\":%.3f", 512.0*16*iters*8.0*sms/ms/1e9);
  }
  // sustained (≈3 s) DFMA and DMMA to see the power-capped clock
  {
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n=0; for (; n<60; n++) dfma_k<8><<<sms*2,1024>>>(out, iters*4, 0.5);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
    printf(",\"dfma_sustained_tflops\":%.3f,\"dfma_sustained_s\":%.2f", 2.0*8*iters*4*1024.0*sms*2*n/ms/1e9, ms/1e3);
    CK(cudaEventRecord(e0));
    for (n=0; n<60; n++) dmma_k<8><<<sms*2,1024>>>(out, iters*4, 0.5);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
    printf(",\"dmma_sustained_tflops\":%.3f,\"dmma_sustained_s\":%.2f", 512.0*8*iters*4*32.0*sms*2*n/ms/1e9, ms/1e3);
  }
  printf("}\n");
  return 0;
}
