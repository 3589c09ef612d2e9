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
//   warp 0        TMA producer: Q tiles once, K and V tiles through 3-stage rings
//   warp 1        MMA issuer:   S_t = Q_t K^T (M128 N128 K64), PV_t = P_t V (M128 N64 K128, V is MN-major)
//   warp 2        TMEM allocator (512 columns: S0 S1 PV0 PV1)
//   warps 4-7     softmax for query tile 0 (thread == query row)
//   warps 8-11    softmax for query tile 1
// While one softmax group exponentiates, the tensor core works on the other tile (ping-pong).
// Online softmax keeps the running max / sum and the fp32 output row in registers; each P V product
// lands in a fresh TMEM buffer and is folded in one iteration later, off the critical path.
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
};

constexpr int AT_BM = 128;  // query rows per tile
constexpr int AT_BN = 128;  // keys per kv tile
constexpr int AT_D = 64;
constexpr int AT_KST = 3;  // K / V ring depth
constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;   // 16 KB
constexpr int AT_KV_BYTES = AT_BN * AT_D * 2;  // 16 KB
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;  // 32 KB (two 64-key swizzle atoms)
constexpr int AT_SMEM = 2 * AT_Q_BYTES + 2 * AT_KST * AT_KV_BYTES + 2 * AT_P_BYTES + 512 + 1024;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(384, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][128][64]
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;                   // [KST][128][64]
  uint8_t* sV = sK + AT_KST * AT_KV_BYTES;             // [KST][128][64]
  uint8_t* sP = sV + AT_KST * AT_KV_BYTES;             // [2][2 atoms][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_P_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [KST]
  uint64_t* k_empty = k_full + AT_KST;     // [KST]
  uint64_t* v_full = k_empty + AT_KST;     // [KST]
  uint64_t* v_empty = v_full + AT_KST;     // [KST]
  uint64_t* s_full = v_empty + AT_KST;     // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;          // [2]
  uint64_t* p_empty = p_full + 2;          // [2]
  uint64_t* pv_full = p_empty + 2;         // [2]
  uint64_t* pv_empty = pv_full + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_empty + 2);

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
      mbar_init(&s_empty[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&p_empty[t], 1);
      mbar_init(&pv_full[t], 1);
      mbar_init(&pv_empty[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S0 [0,128) S1 [128,256) PV0 [256,320) PV1 [320,384)

  // Register re-distribution: the producer / MMA / allocator warpgroup needs few registers; each softmax
  // thread keeps a 128-key score row plus its 64-wide fp32 output row live (12 warps x 168 = 4 x 40 + 8 x 232).
  // (setmaxnreg sits at the top of each role branch so that ptxas allocates registers per role.)

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
      constexpr uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_D, 0, 1);   // P (K-major) x V (MN-major)
      const bool act[2] = {act0, act1};
      auto issue_s = [&](int t, int st) {
        const uint32_t sa = smem_u32(sQ + t * AT_Q_BYTES);
        const uint32_t sb = smem_u32(sK + st * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16_ss(tmem_base + t * AT_BN, umma_smem_desc_sw128(sa + k * 32, 1024, 0),
                       umma_smem_desc_sw128(sb + k * 32, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int t, int st) {
        const uint32_t sa = smem_u32(sP + t * AT_P_BYTES);
        const uint32_t sb = smem_u32(sV + st * AT_KV_BYTES);
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)
          umma_bf16_ss(tmem_base + 2 * AT_BN + t * AT_D,
                       umma_smem_desc_sw128(sa + (k >> 2) * (AT_BM * 128) + (k & 3) * 32, 1024, 0),
                       umma_smem_desc_sw128(sb + k * 2048, 1024, 0), idesc_o, k != 0 ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < 2; ++t) {
        if (!act[t]) continue;
        issue_s(t, 0);
        umma_commit(&s_full[t]);
      }
      umma_commit(&k_empty[0]);  // K stage 0 is released once the S MMAs of kv tile 0 retire
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % AT_KST;
        const uint32_t ph = (j / AT_KST) & 1;
        const int st1 = (j + 1) % AT_KST;
        const uint32_t ph1 = ((j + 1) / AT_KST) & 1;
        const uint32_t par = j & 1;
        bool v_ready = false, k_ready = false;
        for (int t = 0; t < 2; ++t) {
          if (!act[t]) continue;
          // ---- PV_t(j) = P_t(j) V(j)
          mbar_wait(&p_full[t], par);
          if (!v_ready) {
            mbar_wait(&v_full[st], ph);
            v_ready = true;
          }
          if (j > 0) mbar_wait(&pv_empty[t], (j - 1) & 1);
          tc_fence_after();
          issue_pv(t, st);
          umma_commit(&pv_full[t]);
          umma_commit(&p_empty[t]);
          // ---- S_t(j+1) = Q_t K(j+1)^T
          if (j + 1 < n_kv) {
            if (!k_ready) {
              mbar_wait(&k_full[st1], ph1);
              k_ready = true;
            }
            mbar_wait(&s_empty[t], par);
            tc_fence_after();
            issue_s(t, st1);
            umma_commit(&s_full[t]);
          }
        }
        umma_commit(&v_empty[st]);
        if (j + 1 < n_kv) umma_commit(&k_empty[st1]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int t = (warp - 4) >> 2;  // query tile of this softmax group
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row within the tile == TMEM lane
    const bool active = t == 0 ? act0 : act1;
    if (active) {
      const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * AT_BN;
      const uint32_t t_pv = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 2 * AT_BN + t * AT_D;
      uint8_t* myP = sP + t * AT_P_BYTES;
      float o[AT_D];
#pragma unroll
      for (int i = 0; i < AT_D; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f;
      const float c = p.scale_log2;

      for (int j = 0; j < n_kv; ++j) {
        const uint32_t par = j & 1;
        // ---- S(j): the whole 128-key row into registers with one wait, then hand S back to the tensor core
        mbar_wait(&s_full[t], par);
        tc_fence_after();
        uint32_t sv[AT_BN];
#pragma unroll
        for (int c0 = 0; c0 < AT_BN; c0 += 32)
          tmem_ld_32x32b_x32(t_s + c0, *reinterpret_cast<uint32_t(*)[32]>(&sv[c0]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t]);  // S(j+1) = Q K(j+1)^T may now overwrite the accumulator

        const int kv_valid = w.kv_len - j * AT_BN;  // >= 1; < 128 only for the clip's last kv tile
        if (kv_valid < AT_BN) {
#pragma unroll
          for (int i = 0; i < AT_BN; ++i)
            if (i >= kv_valid) sv[i] = 0xff800000u;  // -inf: exp2 -> 0, never the max
        }
        float m_tile = fmaxf(__uint_as_float(sv[0]), __uint_as_float(sv[1]));
#pragma unroll
        for (int i = 2; i < AT_BN; i += 2)
          m_tile = fmax3(m_tile, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        const float m_new = fmaxf(m_run, m_tile);
        const float alpha = ex2_approx((m_run - m_new) * c);  // 0 on the first tile
        const float mc = m_new * c;

        // ---- fold in P V of the previous kv tile (its MMA ran while the other query tile was busy)
        if (j > 0) {
          mbar_wait(&pv_full[t], (j - 1) & 1);
          tc_fence_after();
          uint32_t v[AT_D];
          tmem_ld_32x32b_x32(t_pv, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32b_x32(t_pv + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pv_empty[t]);
#pragma unroll
          for (int i = 0; i < AT_D; ++i) o[i] = (o[i] + __uint_as_float(v[i])) * alpha;
        }

        // ---- P(j) = 2^(s*c - m*c) as bf16 into swizzled smem (A operand of P V); row sum in fp32
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int i = 0; i < AT_BN; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), c, -mc));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), c, -mc));
          l0 += p0;
          l1 += p1;
          sv[i >> 1] = pack_bf16x2(p0, p1);
        }
        if (j > 0) mbar_wait(&p_empty[t], (j - 1) & 1);  // P V(j-1) has consumed the P buffer
#pragma unroll
        for (int q = 0; q < AT_BN / 8; ++q) {
          uint8_t* atom = myP + (q >> 3) * (AT_BM * 128);
          *reinterpret_cast<uint4*>(atom + sw128_offset(r, q & 7)) =
              make_uint4(sv[4 * q], sv[4 * q + 1], sv[4 * q + 2], sv[4 * q + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        l_run = l_run * alpha + (l0 + l1);
        m_run = m_new;
      }
      // last P V
      {
        mbar_wait(&pv_full[t], (n_kv - 1) & 1);
        tc_fence_after();
        uint32_t v[AT_D];
        tmem_ld_32x32b_x32(t_pv, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32b_x32(t_pv + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o[i] += __uint_as_float(v[i]);
      }
      // epilogue: out = bf16(O / l) * bf16(sigmoid(gate))
      const int qv = w.q_valid[t];
      if (r < qv) {
        const int row = w.q_row0[t] + r;
        const int head = w.q_head[t];
        const float inv_l = 1.0f / l_run;
        const __nv_bfloat16* g = p.gate + static_cast<int64_t>(row) * p.ld + head * AT_D;
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + head * AT_D;
#pragma unroll
        for (int q = 0; q < AT_D / 8; ++q) {
          const uint4 gv = ldg16(g + q * 8);
          const uint32_t gg[4] = {gv.x, gv.y, gv.z, gv.w};
          uint32_t ov[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float g0 = bf16_lo(gg[e]), g1 = bf16_hi(gg[e]);
            const float s0 = bf16r(1.0f / (1.0f + __expf(-g0)));
            const float s1 = bf16r(1.0f / (1.0f + __expf(-g1)));
            const float a0 = bf16r(o[q * 8 + 2 * e] * inv_l);
            const float a1 = bf16r(o[q * 8 + 2 * e + 1] * inv_l);
            ov[e] = pack_bf16x2(a0 * s0, a1 * s1);
          }
          stg16(dst + q * 8, make_uint4(ov[0], ov[1], ov[2], ov[3]));
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
int ttk_attn_varlen_fwd(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                        float softmax_scale, void* out, int64_t ldo, cudaStream_t stream) {
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
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess)
      return TTK_ERR_CUDA;
    attr_done = true;
  }
  attn_fwd_kernel<<<n_work, 384, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  return launch_status();
}

}  // extern "C"
