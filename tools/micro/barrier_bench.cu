// Micro-benchmark: cost of one grid-wide barrier in a persistent cooperative kernel (148 x T threads).
//   variant 0: engine.cuh grid_barrier (atom.add.release + last arriver publishes, ld.acquire poll)
//   variant 1: cooperative_groups grid.sync()
//   variant 2: atom.add.acq_rel arrive, relaxed poll + one fence
//   variant 3: like 2, but only warp 0 of each block waits on the flag; others wait at bar.sync (same)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../include -I ../../ecneproject_b200/csrc barrier_bench.cu -o barrier_bench
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include "engine.cuh"
namespace cg = cooperative_groups;
using namespace ecne;

__device__ __forceinline__ unsigned int barrier_v2(unsigned int* bar, unsigned int& epoch, const unsigned int* payload_src) {
  __shared__ unsigned int s_payload;
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    unsigned long long* rel = (unsigned long long*)(bar + 32);
    unsigned int payload;
    if (old == epoch * gridDim.x - 1) {
      payload = payload_src ? *((volatile const unsigned int*)payload_src) : 0u;
      unsigned long long v = ((unsigned long long)payload << 32) | epoch;
      asm volatile("st.release.gpu.u64 [%0], %1;" ::"l"(rel), "l"(v) : "memory");
    } else {
      unsigned long long v;
      do {
        asm volatile("ld.relaxed.gpu.u64 %0, [%1];" : "=l"(v) : "l"(rel) : "memory");
      } while ((unsigned int)(v & 0xffffffffu) < epoch);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      payload = (unsigned int)(v >> 32);
    }
    s_payload = payload;
  }
  __syncthreads();
  return s_payload;
}

__global__ void k(int variant, int iters, unsigned int* bar, unsigned int* payload, unsigned long long* out, int work) {
  unsigned int epoch = 0;
  cg::grid_group g = cg::this_grid();
  unsigned int acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (work && (threadIdx.x & 31) == 0) atomicAdd(payload + 1 + (blockIdx.x & 7), 1u);  // some global writes before the barrier
    if (variant == 0) acc += grid_barrier(bar, epoch, payload);
    else if (variant == 1) g.sync();
    else acc += barrier_v2(bar, epoch, payload);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (acc & 0);
}

int main() {
  unsigned int *bar, *payload; unsigned long long* out;
  cudaMalloc(&bar, 1024); cudaMalloc(&payload, 1024); cudaMalloc(&out, 148 * 8);
  int iters = 2000;
  for (int threads : {1024, 256}) for (int work : {0, 1}) for (int variant : {0, 1, 2}) {
    cudaMemset(bar, 0, 1024); cudaMemset(payload, 0, 1024);
    void* args[] = {&variant, &iters, &bar, &payload, &out, &work};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(148), dim3(threads), args, 0, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %4d work %d variant %d: %s  %.3f us/barrier, %.0f cycles/barrier (block 0)\n", threads, work, variant,
           cudaGetErrorString(e), 1e3 * ms / iters, (double)h[0] / iters);
  }
  return 0;
}
