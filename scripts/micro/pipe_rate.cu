// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions of the attention
// softmax loop: MUFU.EX2, FFMA2 / FADD2 (packed fp32), FFMA, FMNMX3, F2FP (cvt.rn.bf16x2.f32), LEA-style shift+add,
// and of the loop's mixes. One CTA of 128 * W threads = W warps per scheduler, 16 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../titok_video_b200/csrc/common.cuh"
using namespace ttk;

#define REP16(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)

template <int OP>
__global__ void k(long long* out, float* sink, int iters) {
  float a[16];
  uint64_t q[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = 0.001f * (threadIdx.x + i);
    q[i] = f32x2_pack(a[i], a[i] * 0.5f);
  }
  const uint64_t c2 = f32x2_pack(0.999f, 1.001f), d2 = f32x2_pack(1e-3f, -1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(c2), "l"(d2));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(d2));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(1e-3f));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(a[(i + 2) & 15]));
      if (OP == 5) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        a[i] = __uint_as_float(r);
      }
      if (OP == 6) {
        uint32_t r = __float_as_uint(a[i]);
        asm volatile("{ .reg .u32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1; }" : "+r"(r) : "r"(__float_as_uint(a[(i + 1) & 15])));
        a[i] = __uint_as_float(r);
      }
      if (OP == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(-126.f));
      if (OP == 9) {
        uint32_t r = __float_as_uint(a[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        a[i] = __uint_as_float(r);
      }
      if (OP == 10) {
        uint32_t r = __float_as_uint(a[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r));
        a[i] = __uint_as_float(r);
      }
      if (OP == 11) {
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        a[i] = __uint_as_float(r);
      }
      if (OP == 12) {
        uint32_t r = __float_as_uint(a[i]);
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(__float_as_uint(a[(i + 1) & 15])));
        a[i] = __uint_as_float(r);
      }
      if (OP == 13) {  // candidate loop per pair: FFMA2, F2FP(f16x2), MUFU.f16x2, HADD2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(c2), "l"(d2));
        float x, y;
        f32x2_unpack(q[i], x, y);
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        uint32_t acc = __float_as_uint(a[i]);
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(r));
        a[i] = __uint_as_float(acc);
      }
      if (OP == 8) {  // the MUFU path of the loop per pair: FFMA2, 2 MUFU, FADD2, F2FP
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(c2), "l"(d2));
        float x, y;
        f32x2_unpack(q[i], x, y);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
        uint64_t e = f32x2_pack(x, y);
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 1) & 15]) : "l"(e));
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
        a[i] = __uint_as_float(r);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x, y;
    f32x2_unpack(q[i], x, y);
    s += a[i] + x + y;
  }
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(long long* d, float* sink, const char* name, int per_iter) {
  const int iters = 2048;
  for (int w : {1, 2, 4}) {
    k<OP><<<1, 128 * w>>>(d, sink, iters);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %d warp(s)/scheduler: %.2f cycles per warp-instruction per scheduler\n", name, w,
           (double)h / (iters * 16.0 * per_iter * w));
  }
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 64);
  cudaMalloc(&sink, 64);
  run<0>(d, sink, "MUFU.EX2", 1);
  run<1>(d, sink, "FFMA2 (fma.f32x2)", 1);
  run<2>(d, sink, "FADD2 (add.f32x2)", 1);
  run<3>(d, sink, "FFMA", 1);
  run<4>(d, sink, "FMNMX3 (max.f32 3-input)", 1);
  run<5>(d, sink, "F2FP (cvt.rn.bf16x2.f32)", 1);
  run<6>(d, sink, "SHL+IADD (LEA)", 1);
  run<7>(d, sink, "FMNMX imm", 1);
  run<8>(d, sink, "loop mix per pair (5 instr)", 5);
  run<9>(d, sink, "MUFU.EX2 f16x2", 1);
  run<10>(d, sink, "MUFU.EX2 bf16x2", 1);
  run<11>(d, sink, "F2FP (cvt.rn.f16x2.f32)", 1);
  run<12>(d, sink, "HADD2 (add.f16x2)", 1);
  run<13>(d, sink, "f16x2 loop mix per pair (4)", 4);
  return 0;
}
