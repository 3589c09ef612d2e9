// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) for the operand sources / layouts used by the
// attention kernel. One thread issues `n` MMAs back to back into one accumulator, commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../titok_video_b200/csrc/common.cuh"
using namespace ttk;

// mode 0: SS, A K-major, B K-major | 1: SS, B MN-major | 2: TS (A in TMEM), B MN-major | 3: TS, B K-major
template <int MODE, int N>
__global__ void k(long long* out, int n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tptr, 512);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, (MODE == 1 || MODE == 2) ? 1 : 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const int kk = i & 7;
      const uint64_t db = (MODE == 1 || MODE == 2) ? umma_smem_desc_sw128(sb + kk * 2048, 1024, 64 * 128)
                                                   : umma_smem_desc_sw128(sb + (kk & 3) * 32, 1024, 0);
      if (MODE >= 2)
        umma_bf16_ts(tptr, tptr + 256 + kk * 8, db, idesc, 1u);
      else
        umma_bf16_ss(tptr, umma_smem_desc_sw128(sa + (kk & 3) * 32, 1024, 0), db, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

template <int MODE, int N>
void run(long long* d, const char* name) {
  const int n = 4096;
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int grid : {1, 148}) {
    k<MODE, N><<<grid, 128, 80 * 1024>>>(d, n);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d grid %3d: %.1f cycles / MMA (%s)\n", name, N, grid, (double)h / n, cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 2048 * 8);
  run<0, 128>(d, "SS  A K-major, B K-major");
  run<0, 64>(d, "SS  A K-major, B K-major");
  run<1, 64>(d, "SS  A K-major, B MN-major");
  run<1, 128>(d, "SS  A K-major, B MN-major");
  run<2, 64>(d, "TS  A tmem,    B MN-major");
  run<3, 64>(d, "TS  A tmem,    B K-major");
  run<3, 128>(d, "TS  A tmem,    B K-major");
  return 0;
}
