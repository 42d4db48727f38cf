// Test infrastructure: just enough of the CUDA execution model to run a plain CUDA C kernel (threadIdx / blockIdx / __shared__ /
// __syncthreads, no PTX, no warp intrinsics) on the host. Blocks run one after another; the threads of a block are std::threads and
// __syncthreads is a std::barrier, so shared-memory hazards and index arithmetic behave as on the device. `__shared__` becomes a
// function-local static (one copy, shared by the threads of the running block).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct shim_uint3 {
  unsigned x, y, z;
};
inline thread_local shim_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
inline thread_local std::barrier<>* shim_barrier = nullptr;
inline std::atomic<long> shim_blocks_run{0};

inline void __syncthreads() { shim_barrier->arrive_and_wait(); }
inline int atomicCAS(int* p, int cmp, int val) {
  static std::mutex m;
  std::lock_guard<std::mutex> g(m);
  const int old = *p;
  if (old == cmp) *p = val;
  return old;
}

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

template <class F>
void shim_launch(dim3 grid, dim3 block, F body) {
  const unsigned nt = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        std::barrier<> bar(nt);
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; t++)
          th.emplace_back([&, t] {
            threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = {bx, by, bz};
            blockDim = block;
            gridDim = grid;
            shim_barrier = &bar;
            body();
            bar.arrive_and_drop();  // a thread that has left the kernel no longer takes part in its barriers
          });
        for (auto& x : th) x.join();
        shim_blocks_run++;
      }
}

#define BDF_LAUNCH(kernel, grid, block, stream, ...) shim_launch((grid), (block), [&] { kernel(__VA_ARGS__); })
