// Backward of the variable-length grouped-query attention + sigmoid output gate (head_dim 64) on tcgen05 / TMEM / TMA.
//
// Replaces autograd's backward of  flash_attn_varlen_func(q, k, v, cu_seqlens, ...) * sigmoid(gate)
//   reference: model/base/transformer.py:100-103; flash-attn's backward (flash_attn_interface.py:1370-1443 ->
//   _flash_attn_varlen_backward) computes, per clip and head, with P = softmax(q k^T / 8), D_i = sum_d dO_id O_id:
//       dV = P^T dO        dP = dO V^T        dS = P o (dP - D)        dQ = dS K / 8        dK = dS^T Q / 8
//   followed by the backward of apply_rotary_emb (rope.py:19-27): the conjugate rotation of dQ, dK.
//
// Three launches:
//   ttk_attn_bwd_prep   row kernel: dO = d_out * sigmoid(gate), d_gate, D = rowsum(dO o O) per (row, head)
//   ttk_attn_bwd_dkv    one CTA per (128-key tile, kv head): the tile's K, V stay in shared memory, the (Q, dO) tiles of
//                       every query tile of the clip and every query head of the kv group stream past. The scores are
//                       formed TRANSPOSED (S^T = K Q^T, lanes = keys), so that P^T and dS^T can go back to tensor memory
//                       as the A operand of  dV += P^T dO  and  dK += dS^T Q  (TS-mode MMAs, B = the streamed tile,
//                       MN-major) -- the grouped heads accumulate into the same accumulators.
//   ttk_attn_bwd_dq     one CTA per (128-row query tile, query head): Q, dO stay, (K, V) tiles stream; dQ += dS K.
// Both tensor-core kernels are one template: stationary pair (X1, X2), streamed pair (Y1, Y2),
//   S = X1 Y1^T, dP = X2 Y2^T, [P, dS] = f(S, dP),  acc0 += P Y2 (dkv only),  acc1 += dS Y1.
// P is recomputed from the log-sum-exp the forward kernel saved (no running maximum): two extra GEMMs instead of
// atomics on dQ, and the results are deterministic.
//   warp 0  TMA producer   warp 1  MMA issuer   warp 2  TMEM allocator   warps 4-19  P / dS math (thread = lane row x 32 columns)
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

struct AttnBwdWork {
  int st_row0;    // first packed row of the stationary tile
  int st_valid;   // rows of the tile inside the clip
  int st_head;    // dkv: kv head; dq: query head
  int o_head0;    // dkv: first query head of the kv group; dq: kv head
  int n_heads;    // dkv: query heads per kv head; dq: 1
  int clip_row0;  // first packed row of the clip
  int clip_len;   // rows of the clip
  int pad;
};
static_assert(sizeof(AttnBwdWork) == 32, "AttnBwdWork is mirrored in titok_video_b200/plan.py");

struct AttnBwdParams {
  const AttnBwdWork* work;
  const float* lse;    // [hq][M] log2-domain log-sum-exp of the scaled scores (saved by the forward kernel)
  const float* delta;  // [hq][M] rowsum(dO o O)
  const float* rope;   // [M,60] (cos, sin)
  __nv_bfloat16* dqkv; // [M, ld]: [dq | dgate | dk | dv]
  int64_t ld;
  int M, width, gqa;
  float scale, scale_log2;
};

constexpr int AB_T = 128;            // tile rows (both sides)
constexpr int AB_H = 64;             // rows of the streamed tile per stream
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;  // 16 KB
constexpr int AB_STAGES = 3;
constexpr int AB_STAT_IT = 32;  // dkv: iterations whose per-column statistics (lse, D) are staged in shared memory at once
constexpr int AB_STAT_BYTES = AB_STAT_IT * 256 * 4;
constexpr int AB_SMEM = 2 * AB_TILE + AB_STAGES * 2 * AB_TILE + AB_STAT_BYTES + 256 + 1024;
constexpr int AB_THREADS = 128 + 16 * 32;  // 4 control warps + 16 math warps (four per scheduler: latency hiding)
// TMEM columns. Two STREAMS per CTA: stream s owns rows [64 s, 64 s + 64) of every streamed tile, i.e. a [128 x 64] slice
// of S and dP, its own P / dS buffers and its own barriers; both streams accumulate into the same acc0 / acc1. While the
// math warps of one stream are busy, the tensor core works for the other: the serial latencies of a stream (mbarrier
// round trips, tcgen05.ld / st, MMA issue) hide behind the other stream instead of adding up.
constexpr uint32_t AB_TM_S = 0;     // S0 [0,64)    S1 [64,128)
constexpr uint32_t AB_TM_DP = 128;  // dP0 [128,192) dP1 [192,256)
constexpr uint32_t AB_TM_P = 256;   // P0 [256,288) P1 [288,320)   (64 bf16 = 32 columns)
constexpr uint32_t AB_TM_DS = 320;  // dS0 [320,352) dS1 [352,384)
constexpr uint32_t AB_TM_A0 = 384, AB_TM_A1 = 448;

__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool DKV>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem;                 // stationary: dkv K | dq Q
  uint8_t* sX2 = smem + AB_TILE;       //             dkv V | dq dO
  uint8_t* sY = smem + 2 * AB_TILE;    // ring of [Y1 | Y2]: dkv (Q, dO) | dq (K, V)
  float* cols = reinterpret_cast<float*>(sY + AB_STAGES * 2 * AB_TILE);  // dkv: [AB_STAT_IT iterations][lse 128 | delta 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(cols) + AB_STAT_BYTES);
  uint64_t* x_full = bars;
  uint64_t* y_full = bars + 1;
  uint64_t* y_empty = y_full + AB_STAGES;
  uint64_t* sdp_full = y_empty + AB_STAGES;  // [2] S and dP of the stream's iteration i are in tensor memory
  uint64_t* sdp_empty = sdp_full + 2;        // [2] ... and have been read into registers
  uint64_t* pds_full = sdp_empty + 2;        // [2] P and dS of the stream's iteration i are in tensor memory
  uint64_t* pds_empty = pds_full + 2;        // [2] ... and the MMAs reading them have retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pds_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const AttnBwdWork w = p.work[blockIdx.x];
  const int n_tiles = (w.clip_len + AB_T - 1) / AB_T;
  const int n_it = n_tiles * w.n_heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sdp_full[s], 1);
      mbar_init(&sdp_empty[s], 8);
      mbar_init(&pds_full[s], 8);
      mbar_init(&pds_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // 640 threads start with 96 registers; the control warpgroup drops to 40, the math warps grow to 104
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      mbar_arrive_expect_tx(x_full, 2 * AB_TILE);
      if (DKV) {
        tma_load_2d(sX1, &tmK, x_full, w.st_head * AB_D, w.st_row0);
        tma_load_2d(sX2, &tmV, x_full, w.st_head * AB_D, w.st_row0);
      } else {
        tma_load_2d(sX1, &tmQ, x_full, w.st_head * AB_D, w.st_row0);
        tma_load_2d(sX2, &tmDO, x_full, w.st_head * AB_D, w.st_row0);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_STAGES;
        const uint32_t ph = (it / AB_STAGES) & 1;
        const int head = w.o_head0 + it / n_tiles;
        const int row = w.clip_row0 + (it % n_tiles) * AB_T;
        mbar_wait(&y_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&y_full[st], 2 * AB_TILE);
        uint8_t* y1 = sY + st * 2 * AB_TILE;
        if (DKV) {
          tma_load_2d(y1, &tmQ, &y_full[st], head * AB_D, row);
          tma_load_2d(y1 + AB_TILE, &tmDO, &y_full[st], head * AB_D, row);
        } else {
          tma_load_2d(y1, &tmK, &y_full[st], head * AB_D, row);
          tma_load_2d(y1 + AB_TILE, &tmV, &y_full[st], head * AB_D, row);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AB_T, AB_H, 0, 0);  // X (K-major) x Y rows of the stream (K-major)
      constexpr uint32_t idesc_a = umma_idesc_bf16(AB_T, AB_D, 0, 1);  // P / dS (tmem) x Y rows of the stream (MN-major)
      constexpr uint32_t HOFF = AB_H * 128;                            // byte offset of a stream's rows inside a tile
      auto issue_sdp = [&](int s, int st) {
        const uint32_t y1 = smem_u32(sY + st * 2 * AB_TILE) + s * HOFF;
        const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2);
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_bf16_ss(tmem_base + AB_TM_S + s * AB_H, umma_smem_desc_sw128(x1 + k * 32, 1024, 0),
                       umma_smem_desc_sw128(y1 + k * 32, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_D / 16; ++k)
          umma_bf16_ss(tmem_base + AB_TM_DP + s * AB_H, umma_smem_desc_sw128(x2 + k * 32, 1024, 0),
                       umma_smem_desc_sw128(y1 + AB_TILE + k * 32, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
      };
      auto issue_acc = [&](int s, int st, bool accumulate) {
        const uint32_t y1 = smem_u32(sY + st * 2 * AB_TILE) + s * HOFF;
        if (DKV) {
#pragma unroll
          for (int k = 0; k < AB_H / 16; ++k)  // acc0 (dV) += P^T dO
            umma_bf16_ts(tmem_base + AB_TM_A0, tmem_base + AB_TM_P + s * (AB_H / 2) + k * 8,
                         umma_smem_desc_sw128(y1 + AB_TILE + k * 2048, 1024, 0), idesc_a, (accumulate || k != 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < AB_H / 16; ++k)  // acc1 += dS Y1   (dkv: dK += dS^T Q; dq: dQ += dS K)
          umma_bf16_ts(tmem_base + AB_TM_A1, tmem_base + AB_TM_DS + s * (AB_H / 2) + k * 8,
                       umma_smem_desc_sw128(y1 + k * 2048, 1024, 0), idesc_a, (accumulate || k != 0) ? 1u : 0u);
      };
      mbar_wait(x_full, 0);
      mbar_wait(&y_full[0], 0);
      tc_fence_after();
      for (int s = 0; s < 2; ++s) {
        issue_sdp(s, 0);
        umma_commit(&sdp_full[s]);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_STAGES;
        const uint32_t par = it & 1;
        if (it + 1 < n_it) {
          // S / dP of the next tile as soon as the stream's math warps hold the current ones in registers
          const int st1 = (it + 1) % AB_STAGES;
          mbar_wait(&y_full[st1], ((it + 1) / AB_STAGES) & 1);
          for (int s = 0; s < 2; ++s) {
            mbar_wait(&sdp_empty[s], par);
            tc_fence_after();
            issue_sdp(s, st1);
            umma_commit(&sdp_full[s]);
          }
        }
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&pds_full[s], par);
          tc_fence_after();
          issue_acc(s, st, it > 0 || s > 0);
          umma_commit(&pds_empty[s]);
        }
        umma_commit(&y_empty[st]);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int cw = warp - 4;        // 0..15
    const int s = cw >> 3;          // stream: rows [64 s, +64) of every streamed tile == 64 columns of S / dP
    const int h = (cw >> 2) & 1;    // 32-column half of the stream's slice
    const int quarter = warp & 3;   // TMEM lane quarter
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale_log2;
    const int c0 = h * 32;                     // first column (within the stream's slice) of this thread
    float lse_r = 0.f, delta_r = 0.f;
    if (!DKV && r < w.st_valid) {
      lse_r = p.lse[static_cast<int64_t>(w.st_head) * p.M + w.st_row0 + r];
      delta_r = p.delta[static_cast<int64_t>(w.st_head) * p.M + w.st_row0 + r];
    }
    for (int it = 0; it < n_it; ++it) {
      const uint32_t par = it & 1;
      const int tile = it % n_tiles;
      const int y_valid = w.clip_len - tile * AB_T - s * AB_H;  // valid rows of the stream's slice (>= 64: all)
      // per-column (query row) log-sum-exp and D of the streamed tiles: staged for AB_STAT_IT iterations at a time by all
      // 512 math threads (one barrier pair per 32 iterations instead of one per iteration); +inf / 0 beyond the clip => P = dS = 0
      if (DKV && (it % AB_STAT_IT) == 0) {
        named_bar_sync(3, 512);  // nobody reads the previous chunk any more
        const int chunk = min(AB_STAT_IT, n_it - it) * 256;
        for (int j = threadIdx.x - 128; j < chunk; j += 512) {
          const int it2 = it + (j >> 8);
          const int cidx = j & 255, cc = cidx & 127;
          const int rowc = (it2 % n_tiles) * AB_T + cc;
          const bool ok = rowc < w.clip_len;
          const int64_t gi = static_cast<int64_t>(w.o_head0 + it2 / n_tiles) * p.M + w.clip_row0 + rowc;
          float v;
          if (cidx < 128) v = ok ? p.lse[gi] : __int_as_float(0x7f800000);
          else v = ok ? p.delta[gi] : 0.f;
          cols[j] = v;
        }
        named_bar_sync(3, 512);
      }
      const float* cl = cols + (it % AB_STAT_IT) * 256 + s * AB_H;  // lse at cl[c], D at cl[128 + c] for the stream's column c
      mbar_wait(&sdp_full[s], par);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + AB_TM_S + s * AB_H + c0, sv);
      tmem_ld_32x32b_x32(tmem_base + lane_off + AB_TM_DP + s * AB_H + c0, dv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[s]);  // S / dP of the next iteration may overwrite the accumulators
      // in place: sv[0..15] <- packed P, dv[0..15] <- packed dS
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float l[4], d[4];
        if (DKV) {
          const float4 lv = *reinterpret_cast<const float4*>(cl + c0 + i);
          const float4 dl = *reinterpret_cast<const float4*>(cl + 128 + c0 + i);
          l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
          d[0] = dl.x; d[1] = dl.y; d[2] = dl.z; d[3] = dl.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            l[e] = lse_r;
            d[e] = delta_r;
          }
        }
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[e] = ab_ex2(fmaf(__uint_as_float(sv[i + e]), c, -l[e]));
          if (!DKV && c0 + i + e >= y_valid) pv[e] = 0.f;
        }
        const uint32_t ds0 = pack_bf16x2(pv[0] * (__uint_as_float(dv[i]) - d[0]), pv[1] * (__uint_as_float(dv[i + 1]) - d[1]));
        const uint32_t ds1 = pack_bf16x2(pv[2] * (__uint_as_float(dv[i + 2]) - d[2]), pv[3] * (__uint_as_float(dv[i + 3]) - d[3]));
        sv[i >> 1] = pack_bf16x2(pv[0], pv[1]);
        sv[(i >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
        dv[i >> 1] = ds0;
        dv[(i >> 1) + 1] = ds1;
      }
      if (it > 0) {
        mbar_wait(&pds_empty[s], (it - 1) & 1);  // the MMAs of the previous iteration no longer read P / dS
        tc_fence_after();
      }
      if (DKV) tmem_st_32x32b_x16(tmem_base + lane_off + AB_TM_P + s * (AB_H / 2) + (c0 >> 1), *reinterpret_cast<uint32_t(*)[16]>(&sv[0]));
      tmem_st_32x32b_x16(tmem_base + lane_off + AB_TM_DS + s * (AB_H / 2) + (c0 >> 1), *reinterpret_cast<uint32_t(*)[16]>(&dv[0]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[s]);
    }

    // ---- epilogue: accumulators -> bf16 (RoPE conjugate rotation for dQ / dK) -> dqkv; warp group g = 2 s + h takes 16 of
    // the 64 head dims. Own stream's barrier first (a barrier this thread has followed phase by phase), then stream 1's,
    // whose last MMAs were issued after stream 0's: by then it is at most one phase behind, so the parity wait cannot alias.
    mbar_wait(&pds_empty[s], (n_it - 1) & 1);
    mbar_wait(&pds_empty[1], (n_it - 1) & 1);
    tc_fence_after();
    const int g16 = s * 2 + h;
    const int row = w.st_row0 + r;
    const bool row_ok = r < w.st_valid;
    auto store16 = [&](uint32_t tm_col, float mul, int col, bool rot) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + lane_off + tm_col + g16 * 16, v);
      tmem_ld_wait();
      if (!row_ok) return;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = bf16r(__uint_as_float(v[i]) * mul);
      if (rot) {
        const float2* cs = reinterpret_cast<const float2*>(p.rope) + static_cast<int64_t>(row) * 30 + g16 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (g16 * 8 + i < 30) {  // complex lanes 30, 31 (head dims 60..63) are not rotated (rope.py:22-24)
            const float2 t = __ldg(cs + i);
            const float a = f[2 * i], b = f[2 * i + 1];
            f[2 * i] = a * t.x + b * t.y;
            f[2 * i + 1] = b * t.x - a * t.y;
          }
        }
      }
      __nv_bfloat16* dst = p.dqkv + static_cast<int64_t>(row) * p.ld + col + g16 * 16;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        stg16(dst + q * 8, make_uint4(pack_bf16x2(f[q * 8], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                                      pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7])));
    };
    if (DKV) {
      store16(AB_TM_A0, 1.0f, 2 * p.width + p.gqa + w.st_head * AB_D, false);  // dV
      store16(AB_TM_A1, p.scale, 2 * p.width + w.st_head * AB_D, true);        // dK
    } else {
      store16(AB_TM_A1, p.scale, w.st_head * AB_D, true);  // dQ
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// prep: one warp per packed row. d_out, O [M, width]; gate = qkv[:, width:2*width].
//   s = bf16(sigmoid(gate)); dO = bf16(d_out * s); d_gate = bf16(bf16(d_out * O) * s * (1 - s));
//   delta[h][row] = sum over the head's 64 dims of dO * O (fp32)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ d_out, int64_t ldd,
                                                            const __nv_bfloat16* __restrict__ o, int64_t ldo_,
                                                            const __nv_bfloat16* __restrict__ qkv, int64_t ld,
                                                            __nv_bfloat16* __restrict__ dO, int64_t lddo,
                                                            __nv_bfloat16* __restrict__ dqkv, int64_t ldq,
                                                            float* __restrict__ delta, int M) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  constexpr int width = NV * 256;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 256 + lane * 8;
    const uint4 dv = ldg16(d_out + row * ldd + col);
    const uint4 ov = ldg16(o + row * ldo_ + col);
    const uint4 gv = ldg16(qkv + row * ld + width + col);
    const uint32_t dd[4] = {dv.x, dv.y, dv.z, dv.w}, oo[4] = {ov.x, ov.y, ov.z, ov.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t r_do[4], r_dg[4];
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float res_do[2], res_dg[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float d = hh ? bf16_hi(dd[e]) : bf16_lo(dd[e]);
        const float a = hh ? bf16_hi(oo[e]) : bf16_lo(oo[e]);
        const float g = hh ? bf16_hi(gg[e]) : bf16_lo(gg[e]);
        const float s = bf16r(__fdividef(1.0f, 1.0f + __expf(-g)));
        const float x = bf16r(d * s);
        res_do[hh] = x;
        res_dg[hh] = bf16r(d * a) * s * (1.0f - s);
        part = fmaf(x, a, part);
      }
      r_do[e] = pack_bf16x2(res_do[0], res_do[1]);
      r_dg[e] = pack_bf16x2(res_dg[0], res_dg[1]);
    }
    stg16(dO + row * lddo + col, make_uint4(r_do[0], r_do[1], r_do[2], r_do[3]));
    stg16(dqkv + row * ldq + width + col, make_uint4(r_dg[0], r_dg[1], r_dg[2], r_dg[3]));
    // 8 lanes share a head
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if ((lane & 7) == 0) delta[static_cast<int64_t>(i * 4 + (lane >> 3)) * M + row] = part;
  }
}

template <bool DKV>
static int launch_attn_bwd(const void* qkv, int64_t ld, const void* dO, int64_t lddo, int M, int width, int gqa,
                           const void* work, int n_work, const float* lse, const float* delta, const float* rope,
                           float softmax_scale, void* dqkv, int64_t ldq, cudaStream_t stream) {
  if (!qkv || !dO || !work || !lse || !delta || !rope || !dqkv) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width % 64 != 0 || gqa % 64 != 0 || ld % 8 != 0 || lddo % 8 != 0 || ldq % 8 != 0) return TTK_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(rope) & 7u) != 0) return TTK_ERR_ALIGNMENT;
  if (n_work <= 0) return TTK_OK;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tmQ, tmK, tmV, tmDO;
  if (int e = make_tmap_bf16_2d(&tmQ, base, M, width, ld, AB_T)) return e;
  if (int e = make_tmap_bf16_2d(&tmK, base + 2 * width, M, gqa, ld, AB_T)) return e;
  if (int e = make_tmap_bf16_2d(&tmV, base + 2 * width + gqa, M, gqa, ld, AB_T)) return e;
  if (int e = make_tmap_bf16_2d(&tmDO, dO, M, width, lddo, AB_T)) return e;
  AttnBwdParams p{};
  p.work = static_cast<const AttnBwdWork*>(work);
  p.lse = lse;
  p.delta = delta;
  p.rope = rope;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.ld = ldq;
  p.M = M;
  p.width = width;
  p.gqa = gqa;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  auto kern = attn_bwd_kernel<DKV>;
  static PerDeviceOnce once;  // one per template instantiation
  if (int e = set_smem_attr_once(once, reinterpret_cast<const void*>(kern), AB_SMEM)) return e;
  kern<<<n_work, AB_THREADS, AB_SMEM, stream>>>(tmQ, tmK, tmV, tmDO, p);
  return launch_status();
}

}  // namespace ttk

using namespace ttk;

extern "C" {

int ttk_attn_bwd_prep(const void* d_out, int64_t ldd, const void* o, int64_t ldo, const void* qkv, int64_t ld, int M,
                      int width, void* dO, int64_t lddo, void* dqkv, int64_t ldq, float* delta, cudaStream_t stream) {
  if (!d_out || !o || !qkv || !dO || !dqkv || !delta) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width <= 0 || width % 256 || width > 1024 || ldd % 8 || ldo % 8 || ld % 8 || lddo % 8 || ldq % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  const int grid = (M + 7) / 8;
#define PREP(NV)                                                                                                     \
  attn_bwd_prep_kernel<NV><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(d_out), ldd,                  \
                                                     static_cast<const __nv_bfloat16*>(o), ldo,                     \
                                                     static_cast<const __nv_bfloat16*>(qkv), ld,                    \
                                                     static_cast<__nv_bfloat16*>(dO), lddo,                         \
                                                     static_cast<__nv_bfloat16*>(dqkv), ldq, delta, M)
  switch (width / 256) {
    case 1: PREP(1); break;
    case 2: PREP(2); break;
    case 3: PREP(3); break;
    default: PREP(4); break;
  }
#undef PREP
  return launch_status();
}

int ttk_attn_bwd_dkv(const void* qkv, int64_t ld, const void* dO, int64_t lddo, int M, int width, int gqa,
                     const void* work, int n_work, const float* lse, const float* delta, const float* rope,
                     float softmax_scale, void* dqkv, int64_t ldq, cudaStream_t stream) {
  return launch_attn_bwd<true>(qkv, ld, dO, lddo, M, width, gqa, work, n_work, lse, delta, rope, softmax_scale, dqkv, ldq,
                               stream);
}

int ttk_attn_bwd_dq(const void* qkv, int64_t ld, const void* dO, int64_t lddo, int M, int width, int gqa,
                    const void* work, int n_work, const float* lse, const float* delta, const float* rope,
                    float softmax_scale, void* dqkv, int64_t ldq, cudaStream_t stream) {
  return launch_attn_bwd<false>(qkv, ld, dO, lddo, M, width, gqa, work, n_work, lse, delta, rope, softmax_scale, dqkv, ldq,
                                stream);
}

}  // extern "C"
