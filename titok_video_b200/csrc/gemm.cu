// tcgen05 / TMEM / TMA GEMM family for the packed-token linear layers of the TiTok encoder/decoder.
//
//   C[M,N] = A[M,K] * W[N,K]^T      A, W bf16 (K contiguous), fp32 accumulation in TMEM
//
// One persistent CTA per SM. Warp roles:
//   warp 0      TMA producer  (A tile 128x64, W tile BNx64 per stage, SWIZZLE_128B)
//   warp 1      MMA issuer    (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN, K=16)
//   warp 2      TMEM allocator
//   warps 4..   epilogue      (tcgen05.ld 32x32b: thread == output row), double-buffered accumulators
//
// Epilogues (all round to bf16 exactly where the reference's op boundaries do):
//   EPI_STORE  out = bf16(acc + bias)                    nn.Linear         (blocks.py:93,103,165,173; transformer.py:104,55)
//   EPI_QKV    q|gate|k|v split + interleaved RoPE(q,k)  Attn.forward      (transformer.py:85-98, rope.py:19-27)
//   EPI_GEGLU  h = gelu(gate) * value                    GEGLU.forward     (transformer.py:47-52)
//   EPI_RESID  x' = x + y | RMSNorm(alpha*x + y); xn = RMSNorm(x')   ResidualAttentionBlock (transformer.py:126-146)
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

enum { EPI_STORE = 0, EPI_QKV = 1, EPI_GEGLU = 2, EPI_RESID = 3 };

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  // generic output
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* bias;   // [N] or null
  const int32_t* out_row_map;  // [M] or null; negative entries are skipped
  // EPI_QKV
  const float* rope;  // [M, 60] (cos,sin) pairs for the first 30 complex lanes of every head
  int width;          // q width (= gate width)
  int gqa;            // k width (= v width)
  // EPI_GEGLU
  int inner;
  // EPI_RESID  (N == BN == width)
  const __nv_bfloat16* resid;  // x [M, ldr]
  int64_t ldr;
  __nv_bfloat16* x_out;  // x' [M, ldo]
  __nv_bfloat16* xn_out; // RMSNorm(x') * w_next [M, ldo] (may be null)
  const float* w_post;   // post-norm weight (mode 1)
  const float* w_next;   // next pre-norm weight
  float alpha;
  int mode;  // 0: x + y ; 1: RMSNorm(alpha*x + y)*w_post
};

template <int BN, bool B_MN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN, int EPI, int EPI_WARPS, bool B_MN>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BN, B_MN>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          if constexpr (B_MN) {
            // W given as [K, N] (N contiguous): one [64 k][64 n] box per 64-wide N block
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          } else if constexpr (EPI == EPI_GEGLU) {
            // value rows [n*BN/2, +BN/2) and gate rows [inner + n*BN/2, +BN/2) stacked into one B tile
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * (BN / 2));
            tma_load_2d(sb + (BN / 2) * 128, &tmB, &full_bar[stage], kb * BK, p.inner + n_blk * (BN / 2));
          } else {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc_sw128(sa + k * 32, 1024, 0);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + k * 2048, 1024, BK * 128)
                                     : umma_smem_desc_sw128(sb + k * 32, 1024, 0);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[as]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;               // column split when EPI_WARPS == 8
    constexpr int CSPLIT = EPI_WARPS / 4;    // 1 or 2
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile % p.num_n_tiles;
      const int row = m_blk * BM + row_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;

      if constexpr (EPI == EPI_STORE) {
        int orow = row;
        if (row_ok && p.out_row_map) orow = p.out_row_map[row];
        const bool st_ok = row_ok && orow >= 0;
        constexpr int CPW = BN / CSPLIT;  // columns per warp
#pragma unroll 1
        for (int c0 = chalf * CPW; c0 < (chalf + 1) * CPW; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c0;
          if (st_ok && col0 < p.N) {
            __nv_bfloat16* dst = p.out + static_cast<int64_t>(orow) * p.ldo + col0;
            if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a = __uint_as_float(v[j * 8 + e * 2]);
                  float b = __uint_as_float(v[j * 8 + e * 2 + 1]);
                  if (p.bias) {
                    a += __bfloat162float(p.bias[col0 + j * 8 + e * 2]);
                    b += __bfloat162float(p.bias[col0 + j * 8 + e * 2 + 1]);
                  }
                  w[e] = pack_bf16x2(a, b);
                }
                stg16(dst + j * 8, make_uint4(w[0], w[1], w[2], w[3]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < p.N) {
                  float a = __uint_as_float(v[j]);
                  if (p.bias) a += __bfloat162float(p.bias[col0 + j]);
                  dst[j] = __float2bfloat16_rn(a);
                }
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_QKV) {
        // column classes: [0,w) q (RoPE) | [w,2w) gate | [2w,2w+g) k (RoPE) | [2w+g,2w+2g) v
        constexpr int CPW = BN / CSPLIT;
#pragma unroll 1
        for (int c0 = chalf * CPW; c0 < (chalf + 1) * CPW; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c0;
          if (row_ok && col0 < p.N) {
            const bool is_rope = (col0 < p.width) || (col0 >= 2 * p.width && col0 < 2 * p.width + p.gqa);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = bf16r(__uint_as_float(v[j]));  // Linear output is bf16
            if (is_rope) {
              const int d0 = col0 & 63;  // 0 or 32: first head dim of this chunk
              const float* cs = p.rope + static_cast<int64_t>(row) * 60 + d0;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (d0 + 2 * j < 60) {  // complex lanes 30,31 (dims 60..63) are not rotated
                  const float2 t = *reinterpret_cast<const float2*>(cs + 2 * j);
                  const float x0 = f[2 * j], x1 = f[2 * j + 1];
                  f[2 * j] = x0 * t.x - x1 * t.y;
                  f[2 * j + 1] = x0 * t.y + x1 * t.x;
                }
              }
            }
            __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              stg16(dst + j * 8, make_uint4(pack_bf16x2(f[j * 8], f[j * 8 + 1]), pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]),
                                            pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]),
                                            pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7])));
            }
          }
        }
      } else if constexpr (EPI == EPI_GEGLU) {
        constexpr int HALF = BN / 2;         // value columns [0,HALF), gate columns [HALF,BN)
        constexpr int CPW = HALF / CSPLIT;
#pragma unroll 1
        for (int c0 = chalf * CPW; c0 < (chalf + 1) * CPW; c0 += 32) {
          uint32_t xv[32], gv[32];
          tmem_ld_32x32b_x32(t_row + c0, xv);
          tmem_ld_32x32b_x32(t_row + HALF + c0, gv);
          tmem_ld_wait();
          const int col0 = n_blk * HALF + c0;
          if (row_ok && col0 < p.inner) {
            __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0;
            float h[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = bf16r(__uint_as_float(xv[j]));
              const float g = bf16r(__uint_as_float(gv[j]));
              h[j] = bf16r(gelu_erf(g)) * x;
            }
            if (col0 + 32 <= p.inner) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                stg16(dst + j * 8,
                      make_uint4(pack_bf16x2(h[j * 8], h[j * 8 + 1]), pack_bf16x2(h[j * 8 + 2], h[j * 8 + 3]),
                                 pack_bf16x2(h[j * 8 + 4], h[j * 8 + 5]), pack_bf16x2(h[j * 8 + 6], h[j * 8 + 7])));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.inner) dst[j] = __float2bfloat16_rn(h[j]);
            }
          }
        }
      } else if constexpr (EPI == EPI_RESID) {
        // full output row in this tile (N == BN); thread == row. Values kept as packed bf16 in registers.
        static_assert(EPI != EPI_RESID || EPI_WARPS == 4, "row epilogue uses one thread per row");
        uint32_t xs[BN / 2];
        float ss = 0.f;
        const __nv_bfloat16* xr = p.resid + static_cast<int64_t>(row_ok ? row : 0) * p.ldr;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r = ldg16(xr + c0 + j * 8);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float y0 = bf16r(__uint_as_float(v[j * 8 + e * 2]));
              const float y1 = bf16r(__uint_as_float(v[j * 8 + e * 2 + 1]));
              float a0 = bf16_lo(rr[e]), a1 = bf16_hi(rr[e]);
              if (p.mode == 1) {
                a0 = bf16r(a0 * p.alpha);
                a1 = bf16r(a1 * p.alpha);
              }
              const float s0 = bf16r(a0 + y0), s1 = bf16r(a1 + y1);
              ss += s0 * s0 + s1 * s1;
              xs[(c0 + j * 8 + e * 2) / 2] = pack_bf16x2(s0, s1);
            }
          }
        }
        if (p.mode == 1) {
          const float rstd = 1.0f / sqrtf(ss * (1.0f / BN) + 1e-5f);
          ss = 0.f;
#pragma unroll
          for (int i = 0; i < BN / 2; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(p.w_post + 2 * i);
            const float s0 = bf16r(bf16_lo(xs[i]) * rstd * w.x);
            const float s1 = bf16r(bf16_hi(xs[i]) * rstd * w.y);
            ss += s0 * s0 + s1 * s1;
            xs[i] = pack_bf16x2(s0, s1);
          }
        }
        if (row_ok) {
          __nv_bfloat16* xo = p.x_out + static_cast<int64_t>(row) * p.ldo;
#pragma unroll
          for (int i = 0; i < BN / 8; ++i) stg16(xo + i * 8, make_uint4(xs[4 * i], xs[4 * i + 1], xs[4 * i + 2], xs[4 * i + 3]));
          if (p.xn_out) {
            const float rstd = 1.0f / sqrtf(ss * (1.0f / BN) + 1e-5f);
            __nv_bfloat16* no = p.xn_out + static_cast<int64_t>(row) * p.ldo;
#pragma unroll
            for (int i = 0; i < BN / 8; ++i) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 w = *reinterpret_cast<const float2*>(p.w_next + 8 * i + 2 * e);
                w4[e] = pack_bf16x2(bf16_lo(xs[4 * i + e]) * rstd * w.x, bf16_hi(xs[4 * i + e]) * rstd * w.y);
              }
              stg16(no + i * 8, make_uint4(w4[0], w4[1], w4[2], w4[3]));
            }
          }
        }
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
template <int BN, int EPI, int EPI_WARPS, bool B_MN>
static int launch_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BN, B_MN>;
  if (int e = check_device_sm100()) return e;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return TTK_ERR_BAD_SHAPE;
  CUtensorMap tmA, tmB;
  if (int e = make_tmap_bf16_2d(&tmA, A, p.M, p.K, lda, BM)) return e;
  int e2;
  if (B_MN) {
    e2 = make_tmap_bf16_2d(&tmB, W, p.K, p.N, ldw, BK);  // [K, N] row-major, box [64 k][64 n]
  } else if (EPI == EPI_GEGLU) {
    e2 = make_tmap_bf16_2d(&tmB, W, 2 * static_cast<uint64_t>(p.inner), p.K, ldw, BN / 2);
  } else {
    e2 = make_tmap_bf16_2d(&tmB, W, p.N, p.K, ldw, BN);
  }
  if (e2) return e2;
  p.num_m_tiles = (p.M + BM - 1) / BM;
  if (EPI == EPI_GEGLU)
    p.num_n_tiles = (p.inner + BN / 2 - 1) / (BN / 2);
  else
    p.num_n_tiles = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BK - 1) / BK;
  auto kern = gemm_kernel<BN, EPI, EPI_WARPS, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return TTK_ERR_CUDA;
    attr_done = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 128 + 32 * EPI_WARPS, S::TOTAL, stream>>>(tmA, tmB, p);
  return launch_status();
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// nn.Linear: out[M,N] = A[M,K] @ W[N,K]^T + bias. w_is_kn != 0: W is given as [K,N] (N contiguous).
int ttk_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const void* bias,
                  void* out, int64_t ldo, const int32_t* out_row_map, int w_is_kn, cudaStream_t stream) {
  if (!A || !W || !out) return TTK_ERR_BAD_ARG;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out_row_map = out_row_map;
  if (w_is_kn) {
    if (N % 8 != 0) return TTK_ERR_ALIGNMENT;
    return launch_gemm<128, EPI_STORE, 4, true>(A, lda, W, ldw, p, stream);
  }
  if (N > 128) return launch_gemm<256, EPI_STORE, 8, false>(A, lda, W, ldw, p, stream);
  return launch_gemm<128, EPI_STORE, 4, false>(A, lda, W, ldw, p, stream);
}

// Attn.to_qkv + split + RoPE(q), RoPE(k): out[M, 2w+2g] = [rope(q) | gate | rope(k) | v]
int ttk_gemm_qkv_rope(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, int width, int gqa,
                      const float* rope, void* out, int64_t ldo, cudaStream_t stream) {
  if (!A || !W || !out || !rope) return TTK_ERR_BAD_ARG;
  if (width % 64 != 0 || gqa % 64 != 0 || ldo % 8 != 0) return TTK_ERR_BAD_SHAPE;
  GemmParams p{};
  p.M = M;
  p.N = 2 * width + 2 * gqa;
  p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.rope = rope;
  p.width = width;
  p.gqa = gqa;
  return launch_gemm<256, EPI_QKV, 8, false>(A, lda, W, ldw, p, stream);
}

// GEGLU.w12 + chunk + gelu(gate)*value: out[M, inner]; W12 is [2*inner, K] (value rows first).
int ttk_gemm_geglu(const void* A, int64_t lda, const void* W12, int64_t ldw, int M, int inner, int K, void* out,
                   int64_t ldo, cudaStream_t stream) {
  if (!A || !W12 || !out) return TTK_ERR_BAD_ARG;
  if (inner % 8 != 0 || ldo % 8 != 0) return TTK_ERR_BAD_SHAPE;
  GemmParams p{};
  p.M = M;
  p.N = 2 * inner;
  p.K = K;
  p.inner = inner;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  return launch_gemm<128, EPI_GEGLU, 8, false>(A, lda, W12, ldw, p, stream);
}

// Linear(K -> 256, no bias) fused with the residual / KEEL post-norm and the next pre-norm (width 256 only):
//   y = A @ W^T ; mode 0: x' = x + y ; mode 1: x' = RMSNorm(alpha*x + y) * w_post
//   x_out = x' ; xn_out = RMSNorm(x') * w_next (optional)
int ttk_gemm_resid_norm256(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const void* x,
                           int64_t ldx, int mode, float alpha, const float* w_post, const float* w_next, void* x_out,
                           void* xn_out, int64_t ldo, cudaStream_t stream) {
  if (!A || !W || !x || !x_out) return TTK_ERR_BAD_ARG;
  if (mode == 1 && !w_post) return TTK_ERR_BAD_ARG;
  if (xn_out && !w_next) return TTK_ERR_BAD_ARG;
  if (ldx % 8 != 0 || ldo % 8 != 0) return TTK_ERR_ALIGNMENT;
  GemmParams p{};
  p.M = M;
  p.N = 256;
  p.K = K;
  p.resid = static_cast<const __nv_bfloat16*>(x);
  p.ldr = ldx;
  p.mode = mode;
  p.alpha = alpha;
  p.w_post = w_post;
  p.w_next = w_next;
  p.x_out = static_cast<__nv_bfloat16*>(x_out);
  p.xn_out = static_cast<__nv_bfloat16*>(xn_out);
  p.ldo = ldo;
  return launch_gemm<256, EPI_RESID, 4, false>(A, lda, W, ldw, p, stream);
}

}  // extern "C"
