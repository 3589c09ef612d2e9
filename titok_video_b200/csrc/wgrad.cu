// Weight-gradient GEMM of the packed-token linear layers (backward of nn.Linear w.r.t. its weight):
//
//   dW[N_out, K_in] (fp32, +=) = dY[M, N_out]^T * X[M, K_in]        dY, X bf16 row-major (token rows)
//
// what autograd's mm backward computes for every nn.Linear of the reference under bf16 autocast
// (transformer.py:45,55,83,104; blocks.py:49,67,125,143), with the fp32 result the optimizer's fp32 master
// parameters receive.
//
// The contraction runs over the TOKEN dimension (10^3..10^5 rows) while the output is tiny (256..1408 x 256..768),
// so the kernel is a split-K GEMM: grid = (out tiles of 128) x (in tiles of BN) x (token splits); every CTA
// streams its share of 64-row blocks of both operands with TMA ([64 rows][64 cols] SWIZZLE_128B boxes), feeds them
// to tcgen05.mma as MN-major operands (the token dimension is K, the feature dimension is contiguous), keeps the
// [128 x BN] fp32 partial in tensor memory and reduces it into dW with fp32 atomics.
//   warp 0  TMA producer      warp 1  MMA issuer      warp 2  TMEM allocator      warps 4-7  epilogue
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

constexpr int WG_TM = 128;  // out features per tile (MMA M)
constexpr int WG_BK = 64;   // token rows per pipeline stage (MMA K = 4 x 16)
constexpr int WG_BOX = WG_BK * 128;  // one [64 rows][64 cols] bf16 box

struct WgradParams {
  float* dw;
  int64_t ldw;
  int n_out, k_in;
  int tiles_m, tiles_n, splits;
  int num_k_blocks;  // ceil(M / 64)
  int vec4;          // dw and ldw allow 16-byte vector reductions
};

template <int BN>
struct WgradSmem {
  static constexpr int A_BYTES = (WG_TM / 64) * WG_BOX;
  static constexpr int B_BYTES = (BN / 64) * WG_BOX;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 256 + 1024;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using S = WgradSmem<BN>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x % p.splits;
  const int tile = blockIdx.x / p.splits;
  const int m_blk = tile / p.tiles_n;
  const int n_blk = tile % p.tiles_n;
  // this CTA's share of the 64-row token blocks
  const int per = (p.num_k_blocks + p.splits - 1) / p.splits;
  const int kb0 = split * per;
  const int kb1 = min(p.num_k_blocks, kb0 + per);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (nkb > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int i = 0; i < nkb; ++i) {
          const int stage = i % STAGES;
          const uint32_t phase = (i / STAGES) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int row = (kb0 + i) * WG_BK;
#pragma unroll
          for (int j = 0; j < WG_TM / 64; ++j) tma_load_2d(sa + j * WG_BOX, &tmA, &full_bar[stage], m_blk * WG_TM + j * 64, row);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * WG_BOX, &tmB, &full_bar[stage], n_blk * BN + j * 64, row);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bf16(WG_TM, BN, 1, 1);  // both operands MN-major (token dim = K)
        for (int i = 0; i < nkb; ++i) {
          const int stage = i % STAGES;
          const uint32_t phase = (i / STAGES) & 1;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {
            // 16 token rows = 2048 bytes inside a box; 64-wide feature blocks are WG_BOX bytes apart (LBO)
            const uint64_t da = umma_smem_desc_sw128(sa + k * 2048, 1024, WG_BOX);
            const uint64_t db = umma_smem_desc_sw128(sb + k * 2048, 1024, WG_BOX);
            umma_bf16_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(acc_full);
      }
      __syncwarp();
    } else if (warp >= 4) {
      const int quarter = warp & 3;
      const int row = m_blk * WG_TM + quarter * 32 + lane;  // out feature
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n_blk * BN + c0;
        if (col0 >= p.k_in) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0, v);
        tmem_ld_wait();
        if (row < p.n_out) {
          float* dst = p.dw + static_cast<int64_t>(row) * p.ldw + col0;
          if (p.vec4 && col0 + 32 <= p.k_in) {
            // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the atomic instructions
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.k_in) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

template <int BN>
static int launch_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, int M, int n_out, int k_in, float* dw,
                        int64_t ldw, cudaStream_t stream) {
  using S = WgradSmem<BN>;
  CUtensorMap tmA, tmB;
  if (int e = make_tmap_bf16_2d(&tmA, dy, M, n_out, ldy, WG_BK)) return e;
  if (int e = make_tmap_bf16_2d(&tmB, x, M, k_in, ldx, WG_BK)) return e;
  WgradParams p{};
  p.dw = dw;
  p.ldw = ldw;
  p.n_out = n_out;
  p.k_in = k_in;
  p.tiles_m = (n_out + WG_TM - 1) / WG_TM;
  p.tiles_n = (k_in + BN - 1) / BN;
  p.num_k_blocks = (M + WG_BK - 1) / WG_BK;
  const int tiles = p.tiles_m * p.tiles_n;
  // splits: fill the SMs, but every CTA should contract at least 8 token blocks (512 rows) -- each split pays a full
  // [128 x BN] atomic epilogue, which dominates when the token range per CTA is short
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  const int max_splits = (p.num_k_blocks + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.vec4 = ((reinterpret_cast<uintptr_t>(dw) & 15u) == 0 && ldw % 4 == 0) ? 1 : 0;
  // no empty splits: every CTA gets at least one token block
  const int per = (p.num_k_blocks + splits - 1) / splits;
  splits = (p.num_k_blocks + per - 1) / per;
  p.splits = splits;
  auto kern = wgrad_kernel<BN>;
  static PerDeviceOnce once;  // one per template instantiation
  if (int e = set_smem_attr_once(once, reinterpret_cast<const void*>(kern), S::TOTAL)) return e;
  kern<<<tiles * splits, 256, S::TOTAL, stream>>>(tmA, tmB, p);
  return launch_status();
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// dW[n_out, k_in] (fp32, row pitch ldw) += dY[M, n_out]^T @ X[M, k_in]. n_out, k_in multiples of 8.
int ttk_gemm_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, int M, int n_out, int k_in, float* dw,
                   int64_t ldw, cudaStream_t stream) {
  if (!dy || !x || !dw) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (n_out <= 0 || k_in <= 0 || n_out % 8 || k_in % 8 || ldy % 8 || ldx % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  if (k_in > 128) return launch_wgrad<256>(dy, ldy, x, ldx, M, n_out, k_in, dw, ldw, stream);
  return launch_wgrad<128>(dy, ldy, x, ldx, M, n_out, k_in, dw, ldw, stream);
}

}  // extern "C"
