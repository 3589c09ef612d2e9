// Microbenchmark: tensor-memory read / write bandwidth per SM as seen by tcgen05.ld / tcgen05.st (32x32b shapes),
// for 4, 8 and 16 warps. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../titok_video_b200/csrc/common.cuh"
using namespace ttk;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
  tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
}

template <int MODE>  // 0: ld x32, wait each; 1: 4 x ld x32 then one wait; 2: st x32
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_ld_32x32b_x32(base + ((it * 32) & 511), v);
      tmem_ld_wait();
      acc += v[0] ^ v[13] ^ v[31];
    } else if (MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      const uint32_t col = (it * 128) & 511;
      tmem_ld_32x32b_x32(base + col, a);
      tmem_ld_32x32b_x32(base + col + 32, b);
      tmem_ld_32x32b_x32(base + col + 64, c);
      tmem_ld_32x32b_x32(base + col + 96, d);
      tmem_ld_wait();
      acc += a[0] ^ b[7] ^ c[19] ^ d[31];
    } else {
      tmem_st_32x32b_x32(base + ((it * 32) & 511), v);
      if ((it & 3) == 3) tmem_st_wait();
    }
  }
  if (MODE == 2) tmem_st_wait();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 2048 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 16}) {
      for (int grid : {1, 148}) {
        if (mode == 0) k<0><<<grid, warps * 32>>>(d, iters);
        if (mode == 1) k<1><<<grid, warps * 32>>>(d, iters);
        if (mode == 2) k<2><<<grid, warps * 32>>>(d, iters);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * warps * 4096.0 * (mode == 1 ? 4 : 1);
        printf("mode %d (%s) warps %2d grid %3d: %lld cycles, %.1f B/clk/SM  (%s)\n", mode,
               mode == 0 ? "ld x32 + wait" : mode == 1 ? "4 x ld x32, one wait" : "st x32", warps, grid, h, bytes / h,
               cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
