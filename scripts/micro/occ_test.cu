// Does using tcgen05 (TMEM) limit occupancy to one CTA per SM?  Prints cudaOccupancyMaxActiveBlocksPerMultiprocessor.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../titok_video_b200/csrc/common.cuh"
using namespace ttk;
__global__ void __launch_bounds__(384, 2) k_plain(float* o) { o[threadIdx.x] = threadIdx.x; }
__global__ void __launch_bounds__(384, 2) k_tmem(float* o, int cols) {
  __shared__ uint32_t tptr;
  if (threadIdx.x < 32) tmem_alloc(&tptr, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  o[threadIdx.x + blockIdx.x * 384] = tptr;
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tptr, 256);
}
int main() {
  int a = -1, b = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_plain, 384, 16384);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_tmem, 384, 16384);
  printf("occupancy plain %d, tmem %d\n", a, b);
  float* d; cudaMalloc(&d, 4 * 384 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {148, 296, 592}) {
    cudaEventRecord(e0);
    k_tmem<<<grid, 384, 16384>>>(d, 256);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("k_tmem grid %d: %.3f ms (%s)  [each CTA spins ~1 ms: 2 resident CTAs/SM => 296 CTAs take about as long as 148]\n", grid, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
