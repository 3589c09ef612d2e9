// tcgen05 / TMEM / TMA GEMM family for the packed-token linear layers of the TiTok encoder/decoder.
//
//   C[M,N] = A[M,K] * W[N,K]^T      A, W bf16 (K contiguous), fp32 accumulation in TMEM
//
// One persistent CTA per SM. Warp roles:
//   warp 0      TMA producer  (A tile 128x64, W tile BNx64 per stage, SWIZZLE_128B; EPI_RESID: + residual tile)
//   warp 1      MMA issuer    (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN, K=16)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue      (tcgen05.ld 32x32b: thread == output row; two warps per TMEM lane quarter split the
//                              columns), double-buffered accumulators; results leave through swizzled shared-memory
//                              boxes [32 rows x 64 cols] and TMA stores, so every global write is a full 128-byte line
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
// epilogue warps per kind: 8 (two per TMEM lane quarter) for the store-only epilogues, 16 (four per quarter) for the
// arithmetic-heavy ones (GELU, residual + norms) so that each scheduler has four warps to hide latencies with
// the weight-stationary qkv kernel (BN = 192: three 64-column heads per tile) gives every head box its own warp (12 warps)
__host__ __device__ constexpr int epi_warps(int epi, int bn = 256) {
  return (epi == 2 || epi == 3) ? 16 : ((epi == 1 && bn == 192) ? 12 : 8);
}
constexpr int BOX_BYTES = 32 * 128;  // one [32 rows x 64 bf16] staging box

constexpr int WS_MAX_N = 12;  // n blocks a weight-stationary launch may have (K <= 256 shapes: N <= 3072)

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  // generic output
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* bias;   // [N] or null
  const int32_t* out_row_map;  // [M] or null; negative entries are skipped (direct-store path only)
  int direct;                  // EPI_STORE: 1 = per-thread global stores (row map / unaligned N), 0 = TMA stores
  // EPI_QKV
  const float* rope;  // [M, 60] (cos,sin) pairs for the first 30 complex lanes of every head
  int width;          // q width (= gate width)
  int gqa;            // k width (= v width)
  float* knorm2;      // optional [gqa / 64][M]: |k|^2 of every row and kv head (for the attention kernel's score bound)
  // EPI_GEGLU
  int inner;
  // EPI_RESID  (N == BN == 256)
  const float* w_post;  // post-norm weight (mode 1)
  const float* w_next;  // next pre-norm weight (null: no xn output)
  float alpha;
  int mode;  // 0: x + y ; 1: RMSNorm(alpha*x + y)*w_post
  long long* trace;  // optional [gridDim][64] clock64 stamps (ttk_debug_set_trace), null in production
  // weight-stationary kernels: CTAs [ws_begin[n], ws_begin[n+1]) share n block n (sized by the block's share of the work)
  int ws_begin[WS_MAX_N + 1];
};

// development aid: per-CTA timeline of the three pipelines (slot = 1 + 8 * local tile + event)
__device__ __forceinline__ void trace_stamp(const GemmParams& p, int slot) {
  if (p.trace && slot < 64) p.trace[blockIdx.x * 64 + slot] = clock64();
}

// WS ("weight-stationary", K <= 256): the CTA keeps the whole [BN x K] weight tile of ONE n block resident in shared
// memory and streams only A tiles. At K = 256 a streamed 128x256 tile needs 192 KB of operands for 2048 MMA cycles,
// more than the L2 can feed every SM (~50 B/clk/SM measured); with the weights resident it is 64 KB per tile.
template <int BN, int EPI, bool WS>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int RES_KB = 4;  // resident k blocks (K <= 256)
  static constexpr int RES_BYTES = WS ? RES_KB * B_BYTES : 0;
  static constexpr int STAGE_BYTES = WS ? A_BYTES : A_BYTES + B_BYTES;
  // EPI_RESID: the [128 x 256] residual tile (4 swizzled boxes of 128 rows x 64 cols), updated in place and stored
  // from there. Other epilogues: staging boxes per epilogue warp (double-buffered unless the weights are resident).
  static constexpr int OUT_BUFS = WS ? 1 : 2;
  static constexpr int OUT_BOXES = (EPI == EPI_QKV) ? epi_warps(EPI, BN) : 8;  // staging boxes (GEGLU: one per warp pair)
  static constexpr int OUT_BYTES = (EPI == EPI_RESID) ? BM * 256 * 2 : OUT_BOXES * OUT_BUFS * BOX_BYTES;
  static constexpr int AUX_BYTES = (EPI == EPI_RESID) ? 2 * 256 * 4 + 2 * 4 * 128 * 4 : 0;  // norm weights + partial sums
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 227 * 1024 - 1024 - RES_BYTES - OUT_BYTES - AUX_BYTES - BAR_BYTES;
  static constexpr int STAGES = WS ? 4 : ((BUDGET / STAGE_BYTES) > 6 ? 6 : (BUDGET / STAGE_BYTES));
  static constexpr int TOTAL = RES_BYTES + STAGES * STAGE_BYTES + OUT_BYTES + AUX_BYTES + BAR_BYTES + 1024;  // +1024: alignment
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// Which tiles a CTA walks. Streaming kernels: tile = blockIdx.x + i * gridDim.x over the (m, n) grid. Weight-stationary
// kernels: one n block for the CTA's lifetime (CTAs [ws_begin[n], ws_begin[n+1]) own n block n: a partly filled last block
// -- GEGLU: inner = 704 = 5.5 blocks of 128 -- gets fewer CTAs), m blocks strided over the CTAs of that n block.
template <bool WS>
struct TileWalk {
  int m_blk, n_blk, step, num_m, num_n, tile;
  __device__ __forceinline__ TileWalk(int num_m_tiles, int num_n_tiles, const int* ws_begin)
      : num_m(num_m_tiles), num_n(num_n_tiles) {
    if (WS) {
      n_blk = 0;
      while (n_blk + 1 < num_n && static_cast<int>(blockIdx.x) >= ws_begin[n_blk + 1]) ++n_blk;
      m_blk = blockIdx.x - ws_begin[n_blk];
      step = ws_begin[n_blk + 1] - ws_begin[n_blk];  // CTAs that share this n block
    } else {
      tile = blockIdx.x;
      step = gridDim.x;
      m_blk = tile / num_n;
      n_blk = tile % num_n;
    }
  }
  __device__ __forceinline__ bool valid() const { return WS ? (m_blk < num_m) : (tile < num_m * num_n); }
  __device__ __forceinline__ void next() {
    if (WS) {
      m_blk += step;
    } else {
      tile += step;
      m_blk = tile / num_n;
      n_blk = tile % num_n;
    }
  }
};

// gelu(g) = 0.5 g (1 + erf(g / sqrt 2)) for two values at once, erf from Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7; two MUFU ops per value: rcp and ex2), evaluated with packed FFMA2 / FMUL2:
//   t = 1 / (1 + p |g| / sqrt 2),  erf(|x|) = 1 - (a1 t + .. + a5 t^5) exp(-g^2 / 2),  gelu = 0.5 g + 0.5 |g| erf(|x|)
// Against the exact function rounded to bf16 it differs on 117 of the 65280 finite bf16 inputs, all of them
// g < -3 with |gelu| < 3e-3 (torch's own CPU bf16 kernel differs on 765); see tests/test_gpu_kernels.py.
__device__ __forceinline__ void gelu_erf_x2(float g0, float g1, float& o0, float& o1) {
  const uint64_t g = f32x2_pack(g0, g1);
  const uint64_t ag = f32x2_pack(fabsf(g0), fabsf(g1));
  float d0, d1, a0, a1, t0, t1, e0, e1;
  f32x2_unpack(f32x2_fma(ag, f32x2_pack(0.23164189f, 0.23164189f), f32x2_pack(1.0f, 1.0f)), d0, d1);  // p / sqrt 2
  f32x2_unpack(f32x2_mul(f32x2_mul(g, g), f32x2_pack(-0.72134752f, -0.72134752f)), a0, a1);           // -log2(e) / 2
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const uint64_t t = f32x2_pack(t0, t1), e = f32x2_pack(e0, e1);
  // negated coefficients: q = -(a1 t + ... + a5 t^5)
  uint64_t q = f32x2_fma(f32x2_pack(-1.061405429f, -1.061405429f), t, f32x2_pack(1.453152027f, 1.453152027f));
  q = f32x2_fma(q, t, f32x2_pack(-1.421413741f, -1.421413741f));
  q = f32x2_fma(q, t, f32x2_pack(0.284496736f, 0.284496736f));
  q = f32x2_fma(q, t, f32x2_pack(-0.254829592f, -0.254829592f));
  q = f32x2_mul(q, t);
  const uint64_t y = f32x2_fma(q, e, f32x2_pack(1.0f, 1.0f));  // erf(|x|)
  const uint64_t half = f32x2_pack(0.5f, 0.5f);
  f32x2_unpack(f32x2_fma(f32x2_mul(ag, half), y, f32x2_mul(g, half)), o0, o1);
}

// Per-warp staging of [32 rows x 64 cols] bf16 boxes for TMA stores (double-buffered).
template <int NBUF>
struct BoxStager {
  uint8_t* base;  // NBUF * BOX_BYTES, 1024-byte aligned
  int buf;
  __device__ __forceinline__ uint8_t* acquire(int lane) {
    if (lane == 0) tma_store_wait_read<NBUF - 1>();  // the store that last read this buffer has finished
    __syncwarp();
    return base + buf * BOX_BYTES;
  }
  // all lanes have written their rows
  __device__ __forceinline__ void flush(const CUtensorMap* tm, int lane, int col0, int row0) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm, base + buf * BOX_BYTES, col0, row0);
      tma_store_commit();
    }
    if (NBUF > 1) buf ^= 1;
  }
};

// thread == row: write 64 bf16 (32 packed words) of this lane's row into a swizzled box
__device__ __forceinline__ void box_write_row(uint8_t* box, int lane, const uint32_t (&pk)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(box + sw128_offset(lane, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

template <int BN, int EPI, bool B_MN, bool WS>
__global__ void __launch_bounds__(128 + 32 * epi_warps(EPI, BN), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO,   // out / x_out: box [32 rows x 64 cols]
            const __grid_constant__ CUtensorMap tmO2,  // EPI_RESID: xn_out, same box
            const __grid_constant__ CUtensorMap tmR,   // EPI_RESID: residual x, box [128 rows x 64 cols]
            const GemmParams p) {
  using S = GemmSmem<BN, EPI, WS>;
  constexpr int STAGES = S::STAGES;
  constexpr int EPI_WARPS = epi_warps(EPI, BN);
  static_assert(!WS || (!B_MN && EPI != EPI_RESID), "weight-stationary mode: K-major weights, store-type epilogues");
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of the 64-column store box");
  static_assert(EPI != EPI_RESID || BN == 256, "row epilogue owns complete 256-wide rows");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* sres = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem = sres + S::RES_BYTES;  // operand stage ring (behind the resident weight tile, if any)
  uint8_t* smem_out = smem + STAGES * S::STAGE_BYTES;
  float* wsm = reinterpret_cast<float*>(smem_out + S::OUT_BYTES);  // [2][256]
  float* ssm = wsm + 512;                                          // [2 exchanges][4 column parts][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + S::OUT_BYTES + S::AUX_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* resid_full = bars + 2 * STAGES + 4;
  uint64_t* resid_empty = bars + 2 * STAGES + 5;
  uint64_t* res_full = bars + 2 * STAGES + 6;  // WS: the resident weight tile has landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 7);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if constexpr (EPI == EPI_RESID) {
      tma_prefetch_desc(&tmO2);
      tma_prefetch_desc(&tmR);
    }
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
    mbar_init(resid_full, 1);
    mbar_init(resid_empty, EPI_WARPS);
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) trace_stamp(p, 0);
  // everything above overlapped the predecessor's tail (programmatic dependent launch); its results are needed from here
  pdl_wait();
  pdl_launch_dependents();  // persistent grid: every CTA is resident, the successor may set itself up whenever an SM frees

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      TileWalk<WS> tw(p.num_m_tiles, p.num_n_tiles, p.ws_begin);
      if constexpr (WS) {
        // the CTA's weight tile: all k blocks of its n block, once
        if (tw.valid()) {
          mbar_arrive_expect_tx(res_full, p.num_k_blocks * S::B_BYTES);
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
            uint8_t* sb = sres + kb * S::B_BYTES;
            if constexpr (EPI == EPI_GEGLU) {
              tma_load_2d(sb, &tmB, res_full, kb * BK, tw.n_blk * (BN / 2));
              tma_load_2d(sb + (BN / 2) * 128, &tmB, res_full, kb * BK, p.inner + tw.n_blk * (BN / 2));
            } else {
              tma_load_2d(sb, &tmB, res_full, kb * BK, tw.n_blk * BN);
            }
          }
        }
      }
      for (; tw.valid(); tw.next(), ++it) {
        const int m_blk = tw.m_blk;
        const int n_blk = tw.n_blk;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kb == 0) trace_stamp(p, 1 + 8 * it + 0);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          if constexpr (WS) {
            (void)sb;
            (void)n_blk;
          } else
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
          if constexpr (EPI == EPI_RESID) {
            // residual tile of this M block. First tile: right behind the first operand stage (nothing to wait for).
            // Later tiles: behind the LAST operand stage, because the load has to wait until the previous tile's
            // epilogue has stored its rows and must not hold back the operands that let this tile's MMAs overlap it.
            // (the load of a later tile is issued up to a whole epilogue after this point: an L2 prefetch four k blocks
            // ahead takes the HBM latency out of the chain residual load -> epilogue -> store that paces the kernel;
            // K = 256: 57.8 -> 54.6 us per launch at 64 clips. Prefetching earlier than that at K = 704 cost 1.5 %.)
            if (it != 0 && kb == (p.num_k_blocks > 4 ? p.num_k_blocks - 4 : 0)) {
#pragma unroll
              for (int b = 0; b < 4; ++b) tma_prefetch_l2_2d(&tmR, b * 64, m_blk * BM);
            }
            if (kb == (it == 0 ? 0 : p.num_k_blocks - 1)) {
              mbar_wait(resid_empty, (it & 1) ^ 1);
              mbar_arrive_expect_tx(resid_full, BM * 256 * 2);
#pragma unroll
              for (int b = 0; b < 4; ++b) tma_load_2d(smem_out + b * (BM * 128), &tmR, resid_full, b * 64, m_blk * BM);
            }
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
      int itm = 0;
      TileWalk<WS> tw(p.num_m_tiles, p.num_n_tiles, p.ws_begin);
      if constexpr (WS) {
        if (tw.valid()) mbar_wait(res_full, 0);
      }
      for (; tw.valid(); tw.next(), ++itm) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        trace_stamp(p, 1 + 8 * itm + 1);
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0) trace_stamp(p, 1 + 8 * itm + 2);
          if (kb == p.num_k_blocks - 1) trace_stamp(p, 1 + 8 * itm + 3);
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = WS ? smem_u32(sres + kb * S::B_BYTES) : sa + S::A_BYTES;
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
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int cpart = ew >> 2;     // column part of this warp (0..EPI_WARPS/4-1)
    const int row_in_tile = quarter * 32 + lane;
    // store-only epilogues: a private double-buffered staging box per warp. GEGLU: warps (cpart 2b, 2b+1) of a lane
    // quarter fill the two halves of one shared 64-column box.
    BoxStager<S::OUT_BUFS> stg{smem_out + (EPI == EPI_GEGLU ? (quarter * 2 + (cpart >> 1)) : ew) * S::OUT_BUFS * BOX_BYTES, 0};
    if constexpr (EPI == EPI_RESID) {
      for (int te = threadIdx.x - 128; te < 256; te += 32 * EPI_WARPS) {
        wsm[te] = (p.mode == 1) ? p.w_post[te] : 1.0f;
        wsm[256 + te] = p.w_next ? p.w_next[te] : 1.0f;
      }
      named_bar_sync(1, 32 * EPI_WARPS);
    }
    int as = 0;
    uint32_t aphase = 0;
    uint32_t it = 0;
    for (TileWalk<WS> tw(p.num_m_tiles, p.num_n_tiles, p.ws_begin); tw.valid(); tw.next(), ++it) {
      const int m_blk = tw.m_blk;
      const int n_blk = tw.n_blk;
      const int row0 = m_blk * BM + quarter * 32;  // first row of this warp's 32-row slice
      const int row = row0 + lane;
      const bool row_ok = row < p.M;

      if constexpr (EPI == EPI_STORE) {
        constexpr int CPW = BN / 2;  // columns per warp
        if (p.direct) {
          if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 4);
          mbar_wait(&tmem_full[as], aphase);
          tc_fence_after();
          if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 5);
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
          int orow = row;
          if (row_ok && p.out_row_map) orow = p.out_row_map[row];
          const bool st_ok = row_ok && orow >= 0;
#pragma unroll 1
          for (int c0 = cpart * CPW; c0 < (cpart + 1) * CPW; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c0, v);
            tmem_ld_wait();
            const int col0 = n_blk * BN + c0;
            if (st_ok && col0 < p.N) {
              __nv_bfloat16* dst = p.out + static_cast<int64_t>(orow) * p.ldo + col0;
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
        } else {
          // bias of this warp's first box is fetched before the accumulator is ready (off the critical path)
          uint32_t bb[32];
          auto load_bias = [&](int col0) {
            const uint32_t* b2 = reinterpret_cast<const uint32_t*>(p.bias + col0);  // col0 % 64 == 0: 4-byte aligned
#pragma unroll
            for (int j = 0; j < 32; ++j) bb[j] = (p.bias && col0 + 2 * j < p.N) ? __ldg(b2 + j) : 0u;
          };
          load_bias(n_blk * BN + cpart * CPW);
          if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 4);
          mbar_wait(&tmem_full[as], aphase);
          tc_fence_after();
          if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 5);
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
          for (int cb = cpart * CPW; cb < (cpart + 1) * CPW; cb += 64) {
            const int col0 = n_blk * BN + cb;
            if (col0 >= p.N) break;  // warp-uniform
            uint32_t v0[32], v1[32], pk[32];
            tmem_ld_32x32b_x32(t_row + cb, v0);
            tmem_ld_32x32b_x32(t_row + cb + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pk[j] = pack_bf16x2(__uint_as_float(v0[2 * j]) + bf16_lo(bb[j]), __uint_as_float(v0[2 * j + 1]) + bf16_hi(bb[j]));
              pk[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]) + bf16_lo(bb[16 + j]),
                                       __uint_as_float(v1[2 * j + 1]) + bf16_hi(bb[16 + j]));
            }
            if (cb + 64 < (cpart + 1) * CPW) load_bias(col0 + 64);
            uint8_t* box = stg.acquire(lane);
            box_write_row(box, lane, pk);
            stg.flush(&tmO, lane, col0, row0);
          }
        }
      } else if constexpr (EPI == EPI_QKV) {
        // column classes: [0,w) q (RoPE) | [w,2w) gate | [2w,2w+g) k (RoPE) | [2w+g,2w+2g) v; one 64-col box == one head
        constexpr int CPW = BN / (EPI_WARPS / 4);  // columns per warp: 128 (BN = 256, 8 warps) or 64 (BN = 192, 12 warps)
        // RoPE second pass runs with lane == complex pair over the staged box; the (cos, sin) of the 32 rows of this
        // warp's slice are fetched up front (one contiguous 240-byte run per row), before the accumulator is ready.
        // Complex lanes 30, 31 (head dims 60..63) are not rotated (rope.py:22-24).
        const int colw0 = n_blk * BN + cpart * CPW;
        const bool any_rope = (colw0 < p.width) || (colw0 + CPW > 2 * p.width && colw0 < 2 * p.width + p.gqa);
        float2 cs[32];
        if (any_rope && lane < 30) {
          const float2* src = reinterpret_cast<const float2*>(p.rope) + static_cast<int64_t>(row0) * 30 + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) cs[rr] = (row0 + rr < p.M) ? __ldg(src + rr * 30) : make_float2(1.f, 0.f);
        }
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 4);
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 5);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
        for (int cb = cpart * CPW; cb < (cpart + 1) * CPW; cb += 64) {
          const int col0 = n_blk * BN + cb;
          if (col0 >= p.N) break;
          uint32_t v0[32], v1[32], pk[32];
          tmem_ld_32x32b_x32(t_row + cb, v0);
          tmem_ld_32x32b_x32(t_row + cb + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {  // Linear output is bf16
            pk[j] = pack_bf16x2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
            pk[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
          }
          if (p.knorm2 && col0 >= 2 * p.width && col0 < 2 * p.width + p.gqa) {
            // |k|^2 of this row's head (one box == one head), from the fp32 accumulators. (The bf16 rounding and the
            // rotation that follow move the norm by < 2^-8 relative; the consumer's bound carries a 2 % margin.)
            float ss = 0.f, ss1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              ss = fmaf(__uint_as_float(v0[j]), __uint_as_float(v0[j]), ss);
              ss1 = fmaf(__uint_as_float(v1[j]), __uint_as_float(v1[j]), ss1);
            }
            ss += ss1;
            if (row_ok) p.knorm2[static_cast<int64_t>((col0 - 2 * p.width) >> 6) * p.M + row] = ss;
          }
          uint8_t* box = stg.acquire(lane);
          box_write_row(box, lane, pk);
          const bool is_rope = (col0 < p.width) || (col0 >= 2 * p.width && col0 < 2 * p.width + p.gqa);
          if (is_rope) {
            __syncwarp();
            if (lane < 30) {
              uint32_t x[32];
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                x[rr] = *reinterpret_cast<const uint32_t*>(box + sw128_offset(rr, lane >> 2) + (lane & 3) * 4);
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const float x0 = bf16_lo(x[rr]), x1 = bf16_hi(x[rr]);
                *reinterpret_cast<uint32_t*>(box + sw128_offset(rr, lane >> 2) + (lane & 3) * 4) =
                    pack_bf16x2(x0 * cs[rr].x - x1 * cs[rr].y, x0 * cs[rr].y + x1 * cs[rr].x);
              }
            }
          }
          stg.flush(&tmO, lane, col0, row0);
        }
      } else if constexpr (EPI == EPI_GEGLU) {
        constexpr int HALF = BN / 2;  // value columns [0,HALF), gate columns [HALF,BN) of the accumulator
        static_assert(HALF == 128, "four column parts of 32 output columns");
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 4);
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 5);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
        const int cb = cpart * 32;
        const int colbox = n_blk * HALF + (cpart >> 1) * 64;  // first output column of the pair's shared box
        const uint32_t pair_bar = 2 + quarter * 2 + (cpart >> 1);
        if (colbox < p.inner) {  // uniform over the pair
          uint32_t pk[16];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {  // 16 columns at a time: 640 threads leave 96 registers per thread
            uint32_t xv[16], gv[16];
            tmem_ld_32x32b_x16(t_row + cb + hh * 16, xv);
            tmem_ld_32x32b_x16(t_row + HALF + cb + hh * 16, gv);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // Linear output is bf16: round value and gate pairs with one packed convert each; the product
              // bf16(gelu) * value is one packed bf16 multiply (exact product, one rounding, like the reference's op)
              const uint32_t x2 = pack_bf16x2(__uint_as_float(xv[2 * j]), __uint_as_float(xv[2 * j + 1]));
              const uint32_t g2 = pack_bf16x2(__uint_as_float(gv[2 * j]), __uint_as_float(gv[2 * j + 1]));
              float a0, a1;
              gelu_erf_x2(bf16_lo(g2), bf16_hi(g2), a0, a1);
              const uint32_t a2 = pack_bf16x2(a0, a1);
              const __nv_bfloat162 h2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&a2), *reinterpret_cast<const __nv_bfloat162*>(&x2));
              pk[hh * 8 + j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          if ((cpart & 1) == 0 && lane == 0) tma_store_wait_read<S::OUT_BUFS - 1>();  // the pair's buffer is free again
          named_bar_sync(pair_bar, 64);
          uint8_t* box = stg.base + stg.buf * BOX_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(box + sw128_offset(lane, (cpart & 1) * 4 + c)) =
                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          fence_proxy_async_smem();
          named_bar_sync(pair_bar, 64);
          if ((cpart & 1) == 0 && lane == 0) {
            tma_store_2d(&tmO, box, colbox, row0);
            tma_store_commit();
          }
          if (S::OUT_BUFS > 1) stg.buf ^= 1;
        }
      } else if constexpr (EPI == EPI_RESID) {
        // thread == (row, column part): 64 columns as packed bf16 in registers; the four parts of a row exchange their
        // partial sums of squares through shared memory. bf16 adds / multiplies run on the packed bf16x2 pipe: one
        // rounding of the exact result, which is what the reference's bf16 ops (fp32 op-math, then round) produce.
        uint32_t xs[32];
        float ss = 0.f;
        uint8_t* rbox = smem_out + cpart * (BM * 128) + quarter * BOX_BYTES;  // this warp's 32 rows of 64-col box `cpart`
        const __nv_bfloat162 alpha2 = __float2bfloat162_rn(p.alpha);
        mbar_wait(resid_full, it & 1);
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 4);
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 5);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // 32 columns at a time (register footprint)
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + cpart * 64 + hh * 32, v);
          tmem_ld_wait();
          if (hh == 1) {
            // accumulator in registers: hand the TMEM stage back before the norm / store phase
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 r = *reinterpret_cast<const uint4*>(rbox + sw128_offset(lane, hh * 4 + c));
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = c * 4 + e;  // packed pair index within this 32-column half
              const uint32_t y = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
              __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&rr[e]);
              if (p.mode == 1) a = __hmul2(a, alpha2);
              const __nv_bfloat162 sv = __hadd2(a, *reinterpret_cast<const __nv_bfloat162*>(&y));
              const uint32_t su = *reinterpret_cast<const uint32_t*>(&sv);
              const float s0 = bf16_lo(su), s1 = bf16_hi(su);
              ss = fmaf(s0, s0, ss);
              ss = fmaf(s1, s1, ss);
              xs[hh * 16 + i] = su;
            }
          }
        }
        if (threadIdx.x == 128 && it == 0) trace_stamp(p, 48);
        // x' goes out in place over the residual rows through TMA (rows >= M are clipped). For long k loops it is issued
        // BEFORE the barrier of the exchange that follows, so that the store's read of shared memory runs under the
        // barrier wait and the staging rows are free again when xn wants them.
        auto store_x = [&]() {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rbox + sw128_offset(lane, c)) = make_uint4(xs[4 * c], xs[4 * c + 1], xs[4 * c + 2], xs[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO, rbox, cpart * 64, row0);
            tma_store_commit();
          }
        };
        // exchange slot: mode 1 uses slots 0 and 1 within a tile; mode 0 alternates them between tiles, so a slot is
        // never rewritten before the partners' reads of its previous value are ordered by a barrier
        float* ex0 = ssm + ((p.mode == 1) ? 0 : static_cast<int>(it & 1)) * 512;
        ex0[cpart * 128 + row_in_tile] = ss;
        // (measured at 64 clips, same box: K = 704 73.7 -> 71.5 us with the early store, K = 256 54.2 -> 56.4 us: there the
        // store competes with the operand stream of the next tile, which is already under way)
        const bool early = p.num_k_blocks > 4;
        if (p.mode != 1 && early) store_x();  // mode 0: x' = x + y is final already
        named_bar_sync(1, 32 * EPI_WARPS);
        ss = (ex0[row_in_tile] + ex0[128 + row_in_tile]) + (ex0[256 + row_in_tile] + ex0[384 + row_in_tile]);
        if (p.mode == 1) {
          const float rstd = 1.0f / sqrtf(ss * (1.0f / 256) + 1e-5f);
          float part = 0.f;
          const float* wp = wsm + cpart * 64;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(wp + 2 * i);
            const uint32_t o = pack_bf16x2(bf16_lo(xs[i]) * rstd * w.x, bf16_hi(xs[i]) * rstd * w.y);
            const float s0 = bf16_lo(o), s1 = bf16_hi(o);
            part = fmaf(s0, s0, part);
            part = fmaf(s1, s1, part);
            xs[i] = o;
          }
          float* ex1 = ssm + 512;
          ex1[cpart * 128 + row_in_tile] = part;
          if (early) store_x();
          named_bar_sync(1, 32 * EPI_WARPS);  // (also orders this tile's reads of slot 0 before the next tile's writes)
          ss = (ex1[row_in_tile] + ex1[128 + row_in_tile]) + (ex1[256 + row_in_tile] + ex1[384 + row_in_tile]);
          if (!early) store_x();
        } else if (!early) {
          store_x();
        }
        if (threadIdx.x == 128 && it == 0) trace_stamp(p, 50);
        if (p.w_next) {
          const float rstd = 1.0f / sqrtf(ss * (1.0f / 256) + 1e-5f);
          const float* wn = wsm + 256 + cpart * 64;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 w = *reinterpret_cast<const float2*>(wn + 2 * i);
            xs[i] = pack_bf16x2(bf16_lo(xs[i]) * rstd * w.x, bf16_hi(xs[i]) * rstd * w.y);
          }
          if (lane == 0) tma_store_wait_read<0>();  // x' has left shared memory
          __syncwarp();
        if (threadIdx.x == 128 && it == 0) trace_stamp(p, 51);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rbox + sw128_offset(lane, c)) = make_uint4(xs[4 * c], xs[4 * c + 1], xs[4 * c + 2], xs[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO2, rbox, cpart * 64, row0);
            tma_store_commit();
          }
        }
        if (threadIdx.x == 128 && it == 0) trace_stamp(p, 52);
        if (lane == 0) {
          tma_store_wait_read<0>();
          mbar_arrive(resid_empty);  // the producer may overwrite this warp's rows with the next residual tile
        }
        __syncwarp();
      }

      if constexpr (EPI != EPI_RESID) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
      }
      if (threadIdx.x == 128) trace_stamp(p, 1 + 8 * it + 6);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait<0>();  // shared memory must outlive the bulk stores that read it
    if (threadIdx.x == 128) trace_stamp(p, 63);
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
struct GemmIo {
  const void* A;
  int64_t lda;
  const void* W;
  int64_t ldw;
  void* out2 = nullptr;         // EPI_RESID: xn_out
  const void* resid = nullptr;  // EPI_RESID: x
  int64_t ldr = 0;
};

#ifndef GEGLU_TAIL_BASE
#define GEGLU_TAIL_BASE 0.6  // share of a tile's cost that does not shrink with its valid columns (GEGLU, 64 clips: 0.0 122 us, 0.3 94.7, 0.6 88.9; equal shares 91.6)
#endif

template <int BN, int EPI, bool B_MN, bool WS = false>
static int launch_gemm(const GemmIo& io, GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BN, EPI, WS>;
  p.trace = g_trace;
  if (int e = check_device_sm100()) return e;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return TTK_ERR_BAD_SHAPE;
  CUtensorMap tmA, tmB, tmO, tmO2, tmR;
  if (int e = make_tmap_bf16_2d(&tmA, io.A, p.M, p.K, io.lda, BM)) return e;
  int e2;
  if (B_MN) {
    e2 = make_tmap_bf16_2d(&tmB, io.W, p.K, p.N, io.ldw, BK);  // [K, N] row-major, box [64 k][64 n]
  } else if (EPI == EPI_GEGLU) {
    e2 = make_tmap_bf16_2d(&tmB, io.W, 2 * static_cast<uint64_t>(p.inner), p.K, io.ldw, BN / 2);
  } else {
    e2 = make_tmap_bf16_2d(&tmB, io.W, p.N, p.K, io.ldw, BN);
  }
  if (e2) return e2;
  const int out_cols = (EPI == EPI_GEGLU) ? p.inner : p.N;
  if (p.direct) {
    tmO = tmA;  // unused
  } else if (int e = make_tmap_bf16_2d(&tmO, p.out, p.M, out_cols, p.ldo, 32)) {
    return e;
  }
  tmO2 = tmO;
  tmR = tmO;
  if (EPI == EPI_RESID) {
    if (io.out2)
      if (int e = make_tmap_bf16_2d(&tmO2, io.out2, p.M, 256, p.ldo, 32)) return e;
    if (int e = make_tmap_bf16_2d(&tmR, io.resid, p.M, 256, io.ldr, BM)) return e;
  }
  p.num_m_tiles = (p.M + BM - 1) / BM;
  if (EPI == EPI_GEGLU)
    p.num_n_tiles = (p.inner + BN / 2 - 1) / (BN / 2);
  else
    p.num_n_tiles = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BK - 1) / BK;
  auto kern = gemm_kernel<BN, EPI, B_MN, WS>;
  static PerDeviceOnce once;  // one per template instantiation
  if (int e = set_smem_attr_once(once, reinterpret_cast<const void*>(kern), S::TOTAL)) return e;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (WS) {
    // CTAs per n block in proportion to the block's work. A partly valid last block (GEGLU: columns past `inner`) keeps the
    // MMA and the barrier traffic of a full tile but only part of the epilogue: weight = GEGLU_TAIL_BASE + the rest by columns.
    if (p.num_n_tiles > WS_MAX_N || grid < p.num_n_tiles) return TTK_ERR_BAD_SHAPE;
    const int bw = (EPI == EPI_GEGLU) ? BN / 2 : BN, cols = (EPI == EPI_GEGLU) ? p.inner : p.N;
    double wt[WS_MAX_N], tot = 0;
    for (int n = 0; n < p.num_n_tiles; ++n) {
      const int valid = cols - n * bw < bw ? cols - n * bw : bw;
      wt[n] = GEGLU_TAIL_BASE + (1.0 - GEGLU_TAIL_BASE) * valid / bw;
      tot += wt[n];
    }
    int used = 0;
    double acc = 0;
    p.ws_begin[0] = 0;
    for (int n = 0; n < p.num_n_tiles; ++n) {
      acc += wt[n];
      int end = static_cast<int>(acc / tot * grid + 0.5);
      const int left = p.num_n_tiles - 1 - n;
      if (end < used + 1) end = used + 1;
      if (end > grid - left) end = grid - left;
      p.ws_begin[n + 1] = used = end;
    }
    p.ws_begin[p.num_n_tiles] = grid;
  }
  return cuda_status(launch_pdl(kern, dim3(grid), dim3(128 + 32 * epi_warps(EPI, BN)), S::TOTAL, stream, tmA, tmB, tmO, tmO2, tmR, p));
}

// weight-stationary mode pays off when K fits the resident tile and every n group gets a few m tiles per CTA
static bool use_ws(int M, int N_tiles_of_256, int K) {
  const int m_tiles = (M + BM - 1) / BM;
  return K <= 4 * BK && m_tiles * N_tiles_of_256 >= 2 * num_sms();
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// Development aid: buf = device int64 [148][64] (zeroed by the caller) or NULL to switch tracing off. Every GEMM launch
// after this call records clock64 stamps of its pipelines for the first local tiles of each CTA (see trace_stamp).
int ttk_debug_set_trace(void* buf) {
  g_trace = static_cast<long long*>(buf);
  return TTK_OK;
}

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
  // TMA stores need 16-byte aligned rows; a row map scatters rows, which a tiled store cannot express
  p.direct = (out_row_map != nullptr || (ldo % 8) != 0 || (N % 2) != 0 || (reinterpret_cast<uintptr_t>(out) & 15u) != 0 ||
              (bias && (reinterpret_cast<uintptr_t>(bias) & 3u) != 0))
                 ? 1
                 : 0;
  GemmIo io{A, lda, W, ldw};
  if (w_is_kn) {
    if (N % 8 != 0) return TTK_ERR_ALIGNMENT;
    // input-gradient GEMMs of the training path (N = 256 / 704 / 768): 256-wide tiles read the A operand half as often
    if (N > 128) return launch_gemm<256, EPI_STORE, true>(io, p, stream);
    return launch_gemm<128, EPI_STORE, true>(io, p, stream);
  }
  if (N > 128) {
    if (!p.direct && use_ws(M, (N + 255) / 256, K)) return launch_gemm<256, EPI_STORE, false, true>(io, p, stream);
    return launch_gemm<256, EPI_STORE, false>(io, p, stream);
  }
  return launch_gemm<128, EPI_STORE, false>(io, p, stream);
}

// Attn.to_qkv + split + RoPE(q), RoPE(k): out[M, 2w+2g] = [rope(q) | gate | rope(k) | v]
int ttk_gemm_qkv_rope(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, int width, int gqa,
                      const float* rope, void* out, int64_t ldo, float* k_norm2, cudaStream_t stream) {
  if (!A || !W || !out || !rope) return TTK_ERR_BAD_ARG;
  if (width % 64 != 0 || gqa % 64 != 0 || ldo % 8 != 0) return TTK_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(rope) & 7u) != 0) return TTK_ERR_ALIGNMENT;
  GemmParams p{};
  p.M = M;
  p.N = 2 * width + 2 * gqa;
  p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.rope = rope;
  p.width = width;
  p.gqa = gqa;
  p.knorm2 = k_norm2;
  GemmIo io{A, lda, W, ldw};
  // Weight-stationary variants measured slower here, twice: 256-wide tiles (two boxes per warp through one staging buffer
  // serialise the RoPE pass and the store) and 192-wide tiles with a warp per head box (0.78 ms vs 0.53 ms per step: four
  // n-blocks of 192 re-read every A tile four times and leave only 37 CTAs per n-block). Streaming 128x256 tiles it is.
  return launch_gemm<256, EPI_QKV, false>(io, p, stream);
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
  GemmIo io{A, lda, W12, ldw};
  if (use_ws(M, (inner + 127) / 128, K)) return launch_gemm<256, EPI_GEGLU, false, true>(io, p, stream);
  return launch_gemm<256, EPI_GEGLU, false>(io, p, stream);
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
  p.mode = mode;
  p.alpha = alpha;
  p.w_post = w_post;
  p.w_next = xn_out ? w_next : nullptr;
  p.out = static_cast<__nv_bfloat16*>(x_out);
  p.ldo = ldo;
  GemmIo io{A, lda, W, ldw};
  io.out2 = xn_out;
  io.resid = x;
  io.ldr = ldx;
  return launch_gemm<256, EPI_RESID, false>(io, p, stream);
}

}  // extern "C"
