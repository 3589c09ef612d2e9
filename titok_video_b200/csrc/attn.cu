// Variable-length, non-causal, grouped-query attention (head_dim 64) on tcgen05 / TMEM / TMA,
// fused with the sigmoid output gate of Attn.forward.
//
// Replaces flash_attn_varlen_func(q, k, v, cu_seqlens, ...) * sigmoid(gate)
//   reference: model/base/transformer.py:100-103 (call), :87 (q|gate|k|v split)
//
// Layout: one packed activation buffer qkv[M, ld] bf16 with column blocks
//   [0,w) q (RoPE applied) | [w,2w) gate | [2w,2w+g) k (RoPE applied) | [2w+g,2w+2g) v
// Rows of one clip are contiguous (latent rows then patch rows); attention never crosses clips.
//
// Work item (one CTA): TWO 128-row query tiles that share one K/V stream (two query heads of the
// same kv group, or two consecutive row tiles of one head). Warp roles:
//   warp 0        TMA producer: Q tiles once, K and V tiles through 4-stage rings
//   warp 1        MMA issuer:   S_t = Q_t K^T  (SS: M128 N128 K64)
//                               O_t += P_t V    (TS: P read from tensor memory, V MN-major in smem, M128 N64 K128)
//   warp 2        TMEM allocator (512 columns: S0 S1 | O0 O1 | P0 P1)
//   warps 4-11    softmax for query tile 0: thread == (query row, half of the 128 keys of a kv tile); the two
//   warps 12-19   softmax for query tile 1   half-row threads agree on the row maximum through shared memory
// Pipeline: S_t(j+1) is issued as soon as the softmax warps have pulled S_t(j) into registers, so the next score
// tile is ready before the exponentials of the current one are done; P_t(j) goes back to tensor memory as bf16
// (tcgen05.st) and the P V product accumulates into O_t in tensor memory. The running maximum is updated lazily:
// O_t / l are rescaled (by the softmax warps themselves) only when the row maximum grew by more than 2^8, which
// after the first one or two kv tiles practically never happens; the final division by l makes the result exact.
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

struct AttnWork {
  int q_row0[2];   // first packed row of each query tile
  int q_valid[2];  // rows of the tile that belong to the clip (0 => tile unused)
  int q_head[2];   // query head of each tile
  int kv_head;
  int kv_row0;  // first packed row of the clip
  int kv_len;   // rows in the clip
  int pad[3];
};
static_assert(sizeof(AttnWork) == 48, "AttnWork is mirrored in titok_video_b200/plan.py");

struct AttnParams {
  const AttnWork* work;
  const __nv_bfloat16* gate;  // qkv + width
  int64_t ld;                 // row pitch of qkv (elements)
  __nv_bfloat16* out;         // [M, ldo]
  int64_t ldo;
  float scale_log2;  // softmax_scale * log2(e)
  __nv_bfloat16* o_save;  // training: un-gated attention output [M, ldo] (null in inference)
  float* lse;             // training: [heads][M] log2-domain log-sum-exp of the scaled scores (null in inference)
  int M;
  long long* trace;  // development aid (ttk_debug_set_trace): [CTA][64] clock64 stamps of kv iterations 3..7
};

// slot = (j - 3) * 12 + event for kv iterations 3..7; events 0-5 softmax warp 4 (tile 0), 6-7 softmax warp 12 (tile 1),
// 8-11 MMA issuer
// (compiled in only with -DAT_TRACE: the stamps cost ~15 % in the softmax loop)
__device__ __forceinline__ void at_stamp(const AttnParams& p, int j, int ev) {
#ifdef AT_TRACE
  if (p.trace && j >= 3 && j < 7) p.trace[blockIdx.x * 64 + (j - 3) * 12 + ev] = clock64();
#endif
}

constexpr int AT_BM = 128;  // query rows per tile
constexpr int AT_BN = 128;  // keys per kv tile
constexpr int AT_D = 64;
constexpr int AT_KST = 4;  // K / V ring depth
constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;   // 16 KB
constexpr int AT_KV_BYTES = AT_BN * AT_D * 2;  // 16 KB
constexpr int AT_XCH_BYTES = 3 * 2 * 2 * 128 * 4;  // half-row exchange: row max (two parities) and row sum
constexpr int AT_SMEM = 2 * AT_Q_BYTES + 2 * AT_KST * AT_KV_BYTES + 512 + AT_XCH_BYTES + 1024;
constexpr int AT_THREADS = 128 + 16 * 32;
// TMEM columns
constexpr uint32_t AT_TM_S = 0;    // S0 [0,128)   S1 [128,256)
constexpr uint32_t AT_TM_O = 256;  // O0 [256,320) O1 [320,384)
constexpr uint32_t AT_TM_P = 384;  // P0 [384,448) P1 [448,512)   (128 keys x bf16 = 64 columns)
#ifndef AT_STAGGER_NS
#define AT_STAGGER_NS 600
#endif
constexpr float AT_RESCALE_LOG2 = 8.0f;  // rescale O only when the row maximum grew by more than 2^8

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef AT_EXP_NOMUFU  // timing experiment only (scripts/attn_bench.py): how much of the kernel is MUFU time?
  return fmaf(x, 1e-3f, 1e-2f);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

template <bool TRAIN>  // TRAIN: also store the un-gated output and the log-sum-exp (compiled out of the inference kernel)
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [2][128][64]
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;        // [KST][128][64]
  uint8_t* sV = sK + AT_KST * AT_KV_BYTES;  // [KST][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT_KST * AT_KV_BYTES);
  uint64_t* q_full = bars;              // [1]
  uint64_t* k_full = bars + 1;          // [KST]
  uint64_t* k_empty = k_full + AT_KST;  // [KST]
  uint64_t* v_full = k_empty + AT_KST;  // [KST]
  uint64_t* v_empty = v_full + AT_KST;  // [KST]
  uint64_t* s_full = v_empty + AT_KST;  // [2]  S_t(j) is in tensor memory
  uint64_t* s_empty = s_full + 2;       // [2]  the softmax warps hold S_t(j) in registers
  uint64_t* p_full = s_empty + 2;       // [2]  P_t(j) is in tensor memory (and O_t has been rescaled if needed)
  uint64_t* pv_done = p_full + 2;       // [2]  O_t += P_t(j) V(j) has retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [3][tile][half][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const AttnWork w = p.work[blockIdx.x];
  const int n_kv = (w.kv_len + AT_BN - 1) / AT_BN;
  const bool act0 = w.q_valid[0] > 0, act1 = w.q_valid[1] > 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], 8);
      mbar_init(&p_full[t], 8);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Register re-distribution: 640 threads start with 96 registers; the control warpgroup drops to 40 and the 16
  // softmax warps (64 scores per thread live) grow to 104 (4 x 32 x 56 freed >= 16 x 32 x 8 taken).
  // setmaxnreg sits at the top of each role branch.
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      const uint32_t q_bytes = (act0 ? AT_Q_BYTES : 0) + (act1 ? AT_Q_BYTES : 0);
      mbar_arrive_expect_tx(q_full, q_bytes);
      if (act0) tma_load_2d(sQ, &tmQ, q_full, w.q_head[0] * AT_D, w.q_row0[0]);
      if (act1) tma_load_2d(sQ + AT_Q_BYTES, &tmQ, q_full, w.q_head[1] * AT_D, w.q_row0[1]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % AT_KST;
        const uint32_t ph = (j / AT_KST) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], AT_KV_BYTES);
        tma_load_2d(sK + st * AT_KV_BYTES, &tmK, &k_full[st], w.kv_head * AT_D, w.kv_row0 + j * AT_BN);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], AT_KV_BYTES);
        tma_load_2d(sV + st * AT_KV_BYTES, &tmV, &v_full[st], w.kv_head * AT_D, w.kv_row0 + j * AT_BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_D, 0, 1);   // P (tmem, K-major) x V (MN-major)
      const bool act[2] = {act0, act1};
      auto issue_s = [&](int t, int st) {
        const uint32_t sa = smem_u32(sQ + t * AT_Q_BYTES);
        const uint32_t sb = smem_u32(sK + st * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16_ss(tmem_base + AT_TM_S + t * AT_BN, umma_smem_desc_sw128(sa + k * 32, 1024, 0),
                       umma_smem_desc_sw128(sb + k * 32, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int t, int st, bool accumulate) {
        const uint32_t sb = smem_u32(sV + st * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)  // 16 keys == 8 packed columns of P
          umma_bf16_ts(tmem_base + AT_TM_O + t * AT_D, tmem_base + AT_TM_P + t * (AT_BN / 2) + k * 8,
                       umma_smem_desc_sw128(sb + k * 2048, 1024, 0), idesc_o, (accumulate || k != 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < 2; ++t) {
        if (!act[t]) continue;
        // Stagger the two query tiles by roughly half a kv iteration: their softmax groups share the four MUFU
        // pipes, and nothing else would ever move them out of phase (measured: in phase, every exponential phase
        // runs at half speed while the pipes idle during the load / max / store phases of both tiles).
        if (t == 1 && act[0]) __nanosleep(AT_STAGGER_NS);
        issue_s(t, 0);
        umma_commit(&s_full[t]);
      }
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % AT_KST;
        const uint32_t ph = (j / AT_KST) & 1;
        const uint32_t par = j & 1;
        if (j + 1 < n_kv) {
          // ---- S_t(j+1) = Q_t K(j+1)^T as soon as S_t(j) has been read out of tensor memory
          const int st1 = (j + 1) % AT_KST;
          mbar_wait(&k_full[st1], ((j + 1) / AT_KST) & 1);
          for (int t = 0; t < 2; ++t) {
            if (!act[t]) continue;
            mbar_wait(&s_empty[t], par);
            tc_fence_after();
            at_stamp(p, j, 8 + t);
            issue_s(t, st1);
            umma_commit(&s_full[t]);
          }
          umma_commit(&k_empty[st1]);
        }
        // ---- O_t += P_t(j) V(j)
        mbar_wait(&v_full[st], ph);
        for (int t = 0; t < 2; ++t) {
          if (!act[t]) continue;
          mbar_wait(&p_full[t], par);
          tc_fence_after();
          at_stamp(p, j, 10 + t);
          issue_pv(t, st, j > 0);
          umma_commit(&pv_done[t]);
        }
        umma_commit(&v_empty[st]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;        // 0..15
    const int t = sw >> 3;          // query tile of this softmax group
    const int half = (sw >> 2) & 1; // key half [64*half, +64) of every kv tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row within the tile == TMEM lane
    const bool active = t == 0 ? act0 : act1;
    if (active) {
      const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t t_s = tmem_base + lane_off + AT_TM_S + t * AT_BN + half * 64;
      const uint32_t t_o = tmem_base + lane_off + AT_TM_O + t * AT_D + half * 32;
      const uint32_t t_p = tmem_base + lane_off + AT_TM_P + t * (AT_BN / 2) + half * 32;
      const uint32_t pair_bar = 1 + t;  // named barrier of this tile's 8 softmax warps
      float* x_mine = xch + (t * 2 + half) * 128 + r;        // + parity * 512 (max) / + 1024 (sum)
      float* x_other = xch + (t * 2 + (half ^ 1)) * 128 + r;
      const float c = p.scale_log2;
      const uint64_t c2 = f32x2_pack(c, c);
      float m_ref = 0.f, l_run = 0.f;  // l_run: this half's share of the row sum

      for (int j = 0; j < n_kv; ++j) {
        const uint32_t par = j & 1;
        // ---- S(j): this thread's 64 scores into registers with one wait, then hand S back to the tensor core
        mbar_wait(&s_full[t], par);
        tc_fence_after();
        if (threadIdx.x == 128) at_stamp(p, j, 0);
        if (threadIdx.x == 384) at_stamp(p, j, 6);
        // the epilogue's gate values: pull this thread's 64-byte segment towards L2 a few kv tiles ahead of its use
        // (measured: the epilogue of a CTA drops from ~9k to ~5.5k cycles)
        if (j == n_kv - 3 || (n_kv < 3 && j == 0)) {
          const __nv_bfloat16* gp = p.gate + static_cast<int64_t>(w.q_row0[t] + min(r, w.q_valid[t] - 1)) * p.ld +
                                    w.q_head[t] * AT_D + half * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
        }
        uint32_t sv[64];
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t]);  // S_t(j+1) may overwrite the accumulator once all 8 warps arrived
        if (threadIdx.x == 128) at_stamp(p, j, 1);

        const int kv_valid = w.kv_len - j * AT_BN - half * 64;  // valid keys of this half; < 64 only in the last kv tile
        if (kv_valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_valid) sv[i] = 0xff800000u;  // -inf: exp2 -> 0, never the max
        }
        float m0 = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
        float m1 = fmaxf(__uint_as_float(sv[2]), __uint_as_float(sv[3]));
#pragma unroll
        for (int i = 4; i < 64; i += 4) {
          m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        }
        // both halves of a row take the same decisions: exchange the half maxima (parity-alternating slots)
#ifdef AT_EXP_NOBAR  // timing experiment only: cost of the half-row exchange
        const float m_tile = fmaxf(m0, m1);
#else
        x_mine[par * 512] = fmaxf(m0, m1);
        named_bar_sync(pair_bar, 256);
        const float m_tile = fmaxf(fmaxf(m0, m1), x_other[par * 512]);  // the first half always holds >= 1 valid key
#endif

        if (threadIdx.x == 128) at_stamp(p, j, 2);
        // ---- lazy maximum: keep the old reference unless some row of this warp outgrew it by 2^8
        if (j == 0) {
          m_ref = m_tile;
        } else if (__any_sync(0xffffffffu, (m_tile - m_ref) * c > AT_RESCALE_LOG2)) {
          const float m_new = fmaxf(m_ref, m_tile);
          const float alpha = ex2_approx((m_ref - m_new) * c);
          mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t holds every product up to kv tile j-1
          tc_fence_after();
          uint32_t o[32];  // this half rescales 32 of the 64 output columns
          tmem_ld_32x32b_x32(t_o, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(t_o, o);
          tmem_st_wait();
          l_run *= alpha;
          m_ref = m_new;
        }

        // ---- P(j) = 2^(s*c - m*c): packed FMA, MUFU ex2, packed row-sum, bf16 pairs (in place, sv[0..31])
        const float nmc = -m_ref * c;
        const uint64_t nmc2 = f32x2_pack(nmc, nmc);
        uint64_t la = f32x2_pack(0.f, 0.f), lb = la;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const uint64_t ta = f32x2_fma(f32x2_pack(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), c2, nmc2);
          const uint64_t tb = f32x2_fma(f32x2_pack(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), c2, nmc2);
          float a0, a1, b0, b1;
          f32x2_unpack(ta, a0, a1);
          f32x2_unpack(tb, b0, b1);
          a0 = ex2_approx(a0);
          a1 = ex2_approx(a1);
          b0 = ex2_approx(b0);
          b1 = ex2_approx(b1);
          la = f32x2_add(la, f32x2_pack(a0, a1));
          lb = f32x2_add(lb, f32x2_pack(b0, b1));
          sv[i >> 1] = pack_bf16x2(a0, a1);
          sv[(i >> 1) + 1] = pack_bf16x2(b0, b1);
        }
        {
          float x0, x1;
          f32x2_unpack(f32x2_add(la, lb), x0, x1);
          l_run += x0 + x1;
        }
        // ---- P(j) -> tensor memory once P V(j-1) no longer reads the buffer
        if (threadIdx.x == 128) at_stamp(p, j, 3);
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
        }
        if (threadIdx.x == 128) at_stamp(p, j, 4);
        tmem_st_32x32b_x32(t_p, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (threadIdx.x == 128) at_stamp(p, j, 5);
        if (threadIdx.x == 384) at_stamp(p, j, 7);
      }

      // ---- epilogue: out = bf16(O / l) * bf16(sigmoid(gate)); this half writes 32 of the 64 head dims
      x_mine[1024] = l_run;
      named_bar_sync(pair_bar, 256);
      const float l_tot = l_run + x_other[1024];
      const float inv_l = __fdividef(1.0f, l_tot);
      if (TRAIN && half == 0 && r < w.q_valid[t])  // the backward kernels recompute P = 2^(s*c - lse)
        p.lse[static_cast<int64_t>(w.q_head[t]) * p.M + w.q_row0[t] + r] = fmaf(m_ref, c, __log2f(l_tot));
      mbar_wait(&pv_done[t], (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_wait();
      const int qv = w.q_valid[t];
      if (r < qv) {
        const int row = w.q_row0[t] + r;
        const int col = w.q_head[t] * AT_D + half * 32;
        const __nv_bfloat16* g = p.gate + static_cast<int64_t>(row) * p.ld + col;
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col;
        __nv_bfloat16* osv = TRAIN ? p.o_save + static_cast<int64_t>(row) * p.ldo + col : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 gv = ldg16(g + q * 8);
          const uint32_t gg[4] = {gv.x, gv.y, gv.z, gv.w};
          uint32_t ov[4], av[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float g0 = bf16_lo(gg[e]), g1 = bf16_hi(gg[e]);
            const float s0 = bf16r(__fdividef(1.0f, 1.0f + __expf(-g0)));
            const float s1 = bf16r(__fdividef(1.0f, 1.0f + __expf(-g1)));
            const float a0 = bf16r(__uint_as_float(o[q * 8 + 2 * e]) * inv_l);
            const float a1 = bf16r(__uint_as_float(o[q * 8 + 2 * e + 1]) * inv_l);
            ov[e] = pack_bf16x2(a0 * s0, a1 * s1);
            av[e] = pack_bf16x2(a0, a1);
          }
          stg16(dst + q * 8, make_uint4(ov[0], ov[1], ov[2], ov[3]));
          if (TRAIN) stg16(osv + q * 8, make_uint4(av[0], av[1], av[2], av[3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// qkv: packed [M, ld] bf16 (see header). work: device array of n_work AttnWork records (built by the
// host planner). out: [M, ldo] bf16 = attention(q,k,v) * sigmoid(gate).
static int attn_fwd_launch(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                           float softmax_scale, void* out, int64_t ldo, void* o_save, float* lse, cudaStream_t stream) {
  if (!qkv || !work || !out) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width % 64 != 0 || gqa % 64 != 0 || ld % 8 != 0 || ldo % 8 != 0) return TTK_ERR_BAD_SHAPE;
  if (n_work <= 0) return TTK_OK;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tmQ, tmK, tmV;
  if (int e = make_tmap_bf16_2d(&tmQ, base, M, width, ld, AT_BM)) return e;
  if (int e = make_tmap_bf16_2d(&tmK, base + 2 * width, M, gqa, ld, AT_BN)) return e;
  if (int e = make_tmap_bf16_2d(&tmV, base + 2 * width + gqa, M, gqa, ld, AT_BN)) return e;
  AttnParams p{};
  p.work = static_cast<const AttnWork*>(work);
  p.gate = base + width;
  p.ld = ld;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.trace = g_trace;
  p.o_save = static_cast<__nv_bfloat16*>(o_save);
  p.lse = lse;
  p.M = M;
  static PerDeviceOnce once_plain, once_train;
  if (int e = set_smem_attr_once(once_plain, reinterpret_cast<const void*>(attn_fwd_kernel<false>), AT_SMEM)) return e;
  if (int e = set_smem_attr_once(once_train, reinterpret_cast<const void*>(attn_fwd_kernel<true>), AT_SMEM)) return e;
  if (o_save)
    attn_fwd_kernel<true><<<n_work, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  else
    attn_fwd_kernel<false><<<n_work, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  return launch_status();
}

int ttk_attn_varlen_fwd(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                        float softmax_scale, void* out, int64_t ldo, cudaStream_t stream) {
  return attn_fwd_launch(qkv, ld, M, width, gqa, work, n_work, softmax_scale, out, ldo, nullptr, nullptr, stream);
}

// Training forward: additionally saves what the backward kernels (attn_bwd.cu) need: o_save [M, ldo] = the attention
// output before the gate, lse fp32 [width/64][M] = log2-domain log-sum-exp of the scaled scores.
int ttk_attn_varlen_fwd_train(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                              float softmax_scale, void* out, int64_t ldo, void* o_save, float* lse,
                              cudaStream_t stream) {
  if (!o_save || !lse) return TTK_ERR_BAD_ARG;
  return attn_fwd_launch(qkv, ld, M, width, gqa, work, n_work, softmax_scale, out, ldo, o_save, lse, stream);
}

}  // extern "C"
