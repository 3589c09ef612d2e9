// Bandwidth-bound row kernels of the TiTok encoder / decoder: one warp per packed token row,
// 16-byte vector loads/stores, warp-shuffle reductions, fp32 math with bf16 rounding at the
// reference's op boundaries.
//
//   rmsnorm / resid_norm   flash-attn RMSNorm (fp32, eps 1e-5, weight only) and the residual / KEEL
//                          updates of ResidualAttentionBlock.forward (transformer.py:126-146)
//   enc_embed              TiTokEncoder.forward embed (blocks.py:95-97)
//   dec_embed              TiTokDecoder.forward embed (blocks.py:164-167)
//   enc_head_fsq           latent gather -> ln_post -> proj_out -> FSQ (blocks.py:101-103, fsq.py:123-135)
//   patchify / unpatchify  einops rearranges of model/base/utils.py:26-51 (feature order permuted to
//                          (c p0 p1 p2) so both sides move 16-byte runs; weights are permuted to match)
#include "common.cuh"
#include "fsq.cuh"
#include "host_util.cuh"

namespace ttk {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA
constexpr float RMS_EPS = 1e-5f;

// Each lane owns NV vectors of 8 consecutive bf16: element index (i*256 + lane*8 + e).
template <int NV>
struct RowVec {
  float v[NV * 8];
  __device__ __forceinline__ void load(const __nv_bfloat16* row, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint4 r = ldg16(row + i * 256 + lane * 8);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[i * 8 + 2 * e] = bf16_lo(rr[e]);
        v[i * 8 + 2 * e + 1] = bf16_hi(rr[e]);
      }
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* row, int lane) const {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      stg16(row + i * 256 + lane * 8,
            make_uint4(pack_bf16x2(v[i * 8], v[i * 8 + 1]), pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]),
                       pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]), pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7])));
  }
  __device__ __forceinline__ float sumsq() const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) s += v[i] * v[i];
    return warp_sum(s);
  }
  // v = bf16(v * rstd * w)   (the RMSNorm output is stored in the activation dtype)
  __device__ __forceinline__ void norm(const float* w, int lane, float rstd) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 w0 = *reinterpret_cast<const float4*>(w + i * 256 + lane * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w + i * 256 + lane * 8 + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i * 8 + e] = bf16r(v[i * 8 + e] * rstd * ww[e]);
    }
  }
};

__device__ __forceinline__ float rstd_of(float sumsq, int width) {
  return 1.0f / sqrtf(sumsq / static_cast<float>(width) + RMS_EPS);
}

// ------------------------------------------------------------------------------------------------
// y = RMSNorm(x) * w
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                 const float* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ y, int64_t ldy, int M) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  RowVec<NV> r;
  r.load(x + row * ldx, lane);
  r.norm(w, lane, rstd_of(r.sumsq(), NV * 256));
  r.store(y + row * ldy, lane);
}

// ------------------------------------------------------------------------------------------------
// mode 0: x' = x + y                         (layer 0, transformer.py:128-130)
// mode 1: x' = RMSNorm(alpha*x + y) * w_post (KEEL, transformer.py:141-145)
// xn = RMSNorm(x') * w_next (optional)       (pre-norm of the next sub-layer / ln_post)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
resid_norm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                  __nv_bfloat16* __restrict__ x_out, __nv_bfloat16* __restrict__ xn_out,
                  const float* __restrict__ w_post, const float* __restrict__ w_next, float alpha, int mode, int M,
                  int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  RowVec<NV> a, b;
  a.load(x + row * ld, lane);
  b.load(y + row * ld, lane);
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) {
    const float xa = (mode == 1) ? bf16r(a.v[i] * alpha) : a.v[i];
    a.v[i] = bf16r(xa + b.v[i]);
  }
  if (mode == 1) a.norm(w_post, lane, rstd_of(a.sumsq(), NV * 256));
  a.store(x_out + row * ld, lane);
  if (xn_out) {
    a.norm(w_next, lane, rstd_of(a.sumsq(), NV * 256));
    a.store(xn_out + row * ld, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder embed. src_row[r] >= 0: patch row, v = bf16(proj[src_row[r]] + mask_token) -> ln_pre_p
//                src_row[r] <  0: latent row, v = mask_token (constant row)            -> ln_pre_t
// x[r] = norm ; xn[r] = RMSNorm(x[r]) * w_next
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
enc_embed_kernel(const __nv_bfloat16* __restrict__ proj, int64_t ldp, const int32_t* __restrict__ src_row,
                 const float* __restrict__ mask_token, const float* __restrict__ w_t, const float* __restrict__ w_p,
                 const float* __restrict__ w_next, __nv_bfloat16* __restrict__ x_out,
                 __nv_bfloat16* __restrict__ xn_out, __nv_bfloat16* __restrict__ e0_out, int M, int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mt = bf16r(mask_token[0]);  // mask_token.to(dtype)
  const int src = src_row[row];
  RowVec<NV> a;
  if (src >= 0) {
    a.load(proj + src * ldp, lane);
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) a.v[i] = bf16r(a.v[i] + mt);
    if (e0_out) a.store(e0_out + row * ld, lane);  // training: the pre-norm row (input of rmsnorm_bwd)
    a.norm(w_p, lane, rstd_of(a.sumsq(), NV * 256));
  } else {
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) a.v[i] = mt;
    if (e0_out) a.store(e0_out + row * ld, lane);
    a.norm(w_t, lane, rstd_of(a.sumsq(), NV * 256));
  }
  a.store(x_out + row * ld, lane);
  if (xn_out) {
    a.norm(w_next, lane, rstd_of(a.sumsq(), NV * 256));
    a.store(xn_out + row * ld, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder embed. src_row[r] >= 0: latent row, v = bf16(bf16(codes[src] @ Win^T + b) + mask_token) -> ln_pre_t
//                src_row[r] <  0: patch row,  v = mask_token                                     -> ln_pre_p
// Win is [width, TS] bf16 (nn.Linear(token_size, width)), TS <= 8.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
dec_embed_kernel(const __nv_bfloat16* __restrict__ codes, int TS, const int32_t* __restrict__ src_row,
                 const __nv_bfloat16* __restrict__ w_in, const __nv_bfloat16* __restrict__ b_in,
                 const float* __restrict__ mask_token, const float* __restrict__ w_t, const float* __restrict__ w_p,
                 const float* __restrict__ w_next, __nv_bfloat16* __restrict__ x_out,
                 __nv_bfloat16* __restrict__ xn_out, __nv_bfloat16* __restrict__ e0_out, int M, int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mt = bf16r(mask_token[0]);
  const int src = src_row[row];
  RowVec<NV> a;
  if (src >= 0) {
    float cv[FSQ_MAX_D];
#pragma unroll
    for (int k = 0; k < FSQ_MAX_D; ++k) cv[k] = (k < TS) ? __bfloat162float(codes[static_cast<int64_t>(src) * TS + k]) : 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int n = i * 256 + lane * 8 + e;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < FSQ_MAX_D; ++k)
          if (k < TS) acc = fmaf(cv[k], __bfloat162float(w_in[n * TS + k]), acc);
        acc += __bfloat162float(b_in[n]);
        a.v[i * 8 + e] = bf16r(bf16r(acc) + mt);
      }
    }
    if (e0_out) a.store(e0_out + row * ld, lane);
    a.norm(w_t, lane, rstd_of(a.sumsq(), NV * 256));
  } else {
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) a.v[i] = mt;
    if (e0_out) a.store(e0_out + row * ld, lane);
    a.norm(w_p, lane, rstd_of(a.sumsq(), NV * 256));
  }
  a.store(x_out + row * ld, lane);
  if (xn_out) {
    a.norm(w_next, lane, rstd_of(a.sumsq(), NV * 256));
    a.store(xn_out + row * ld, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder head: for latent token t: tok = RMSNorm(x[latent_row[t]]) * w_post ; z = bf16(tok @ Wout^T + b)
// then FSQ on z.float(). Wout is [TS, width] bf16. Outputs z [T,TS] bf16, codes [T,TS] bf16, idx [T] int32.
// If pre_normed != 0, x already holds RMSNorm(x) * w_post (fused upstream).
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
enc_head_fsq_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const int32_t* __restrict__ latent_row,
                    const float* __restrict__ w_post, int pre_normed, const __nv_bfloat16* __restrict__ w_out,
                    const __nv_bfloat16* __restrict__ b_out, int TS, __nv_bfloat16* __restrict__ z_out,
                    __nv_bfloat16* __restrict__ codes_out, int32_t* __restrict__ idx_out, int T,
                    const __grid_constant__ FsqConsts c) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (t >= T) return;
  const int row = latent_row ? latent_row[t] : t;  // (null map: x holds the latent rows only, in token order)
  RowVec<NV> a;
  a.load(x + row * ld, lane);
  if (!pre_normed) a.norm(w_post, lane, rstd_of(a.sumsq(), NV * 256));
  float zk[FSQ_MAX_D];
#pragma unroll
  for (int k = 0; k < FSQ_MAX_D; ++k) {
    float acc = 0.f;
    if (k < TS) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint4 r = ldg16(w_out + static_cast<int64_t>(k) * (NV * 256) + i * 256 + lane * 8);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc = fmaf(a.v[i * 8 + 2 * e], bf16_lo(rr[e]), acc);
          acc = fmaf(a.v[i * 8 + 2 * e + 1], bf16_hi(rr[e]), acc);
        }
      }
      acc = warp_sum(acc);
    }
    zk[k] = acc;
  }
  if (lane == 0) {
    float idx = 0.f;
#pragma unroll
    for (int k = 0; k < FSQ_MAX_D; ++k) {
      if (k < TS) {
        const float z = bf16r(zk[k] + __bfloat162float(b_out[k]));
        z_out[static_cast<int64_t>(t) * TS + k] = __float2bfloat16_rn(z);
        const float code = fsq_quantize_dim(z, c, k, idx);
        codes_out[static_cast<int64_t>(t) * TS + k] = __float2bfloat16_rn(code);
      }
    }
    idx_out[t] = static_cast<int32_t>(idx);
  }
}

// ------------------------------------------------------------------------------------------------
// patchify: clips (flat bf16 buffer, each clip [3,T,H,W]) -> patches [G, 3*P0*P1*P2] with feature order
// (c, p0, p1, p2). geom[g] = {element offset of (c=0,t0,h0,w0), W, H*W, T*H*W}; P2 == 8, so a patch is
// C*P0*P1 runs of 16 bytes along W.
// One CTA moves PATCH_CHUNK consecutive patches through shared memory so that BOTH sides are coalesced: on the
// clip side consecutive lanes take the same run of consecutive patches (neighbours along W: one contiguous
// segment), on the patch side a warp streams whole 1.5 KB patch rows.
// ------------------------------------------------------------------------------------------------
constexpr int PATCH_CHUNK = 16;

__device__ __forceinline__ int64_t patch_run_offset(const int64_t* ge, int rr, int P0, int P1) {
  const int c = rr / (P0 * P1);
  const int p0 = (rr / P1) % P0;
  const int p1 = rr % P1;
  return ge[0] + c * ge[3] + p0 * ge[2] + p1 * ge[1];
}

// 8 uint8 pixels -> 8 normalised bf16 values with the rounding sequence of the reference's dataset code on bf16 tensors
// (dataset/video_dataset.py:118-119): y = bf16(bf16(bf16(u8) / 255) * 2 - 1).
__device__ __forceinline__ uint4 normalize_u8x8(uint32_t lo, uint32_t hi) {
  const uint32_t w[2] = {lo, hi};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float f[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float q = bf16r(static_cast<float>((w[k] >> (8 * b)) & 0xffu) / 255.0f);
      f[b] = q * 2.0f - 1.0f;  // q * 2 is exact in bf16; the subtraction is rounded once by the packed convert below
    }
    o[2 * k] = pack_bf16x2(f[0], f[1]);
    o[2 * k + 1] = pack_bf16x2(f[2], f[3]);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// U8: `clips` holds decoded uint8 frames (same geometry, 8-byte runs); they are normalised on the way through (the
// separate ttk_normalize_u8 pass -- 3 bytes of HBM traffic per pixel -- disappears from the tokenisation path).
template <bool U8>
__global__ void __launch_bounds__(256) patchify_kernel(const void* __restrict__ clips_v,
                                                       const int64_t* __restrict__ geom, int C, int P0, int P1,
                                                       __nv_bfloat16* __restrict__ patches, int64_t ldp, int64_t G) {
  const __nv_bfloat16* clips = static_cast<const __nv_bfloat16*>(clips_v);
  const uint8_t* clips8 = static_cast<const uint8_t*>(clips_v);
  extern __shared__ uint4 patch_smem[];  // [PATCH_CHUNK][runs + 1] 16-byte cells (+1: bank-conflict padding)
  __shared__ int64_t sgeom[PATCH_CHUNK * 4];  // the chunk's geometry, read once (not once per 16-byte run)
  const int runs = C * P0 * P1;
  const int pitch = runs + 1;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * PATCH_CHUNK;
  const int np = static_cast<int>(min(static_cast<int64_t>(PATCH_CHUNK), G - g0));
  if (threadIdx.x < np * 4) sgeom[threadIdx.x] = geom[g0 * 4 + threadIdx.x];
  __syncthreads();
  const int total = PATCH_CHUNK * runs;
#pragma unroll 3
  for (int i = threadIdx.x; i < total; i += 256) {
    const int pi = i % PATCH_CHUNK, rr = i / PATCH_CHUNK;
    if (pi < np) {
      const int64_t off = patch_run_offset(sgeom + pi * 4, rr, P0, P1);
      if (U8) {
        const uint2 v = *reinterpret_cast<const uint2*>(clips8 + off);
        patch_smem[pi * pitch + rr] = normalize_u8x8(v.x, v.y);
      } else {
        patch_smem[pi * pitch + rr] = ldg16_stream(clips + off);
      }
    }
  }
  __syncthreads();
  const int total_out = np * runs;
#pragma unroll 3
  for (int i = threadIdx.x; i < total_out; i += 256) {
    const int pi = i / runs, rr = i - pi * runs;
    stg16(patches + (g0 + pi) * ldp + rr * 8, patch_smem[pi * pitch + rr]);
  }
}

// unpatchify: rows of `proj` ([*, ldp], feature order (c,p0,p1,p2)) for the patch rows -> clips.
// patch_row[g] = row of proj holding patch g. Same staging, opposite direction.
__global__ void __launch_bounds__(256) unpatchify_kernel(const __nv_bfloat16* __restrict__ proj, int64_t ldp,
                                                         const int32_t* __restrict__ patch_row,
                                                         const int64_t* __restrict__ geom, int C, int P0, int P1,
                                                         __nv_bfloat16* __restrict__ clips, int64_t G) {
  extern __shared__ uint4 patch_smem[];
  __shared__ int64_t sgeom[PATCH_CHUNK * 4];
  __shared__ int srow[PATCH_CHUNK];
  const int runs = C * P0 * P1;
  const int pitch = runs + 1;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * PATCH_CHUNK;
  const int np = static_cast<int>(min(static_cast<int64_t>(PATCH_CHUNK), G - g0));
  if (threadIdx.x < np * 4) sgeom[threadIdx.x] = geom[g0 * 4 + threadIdx.x];
  if (threadIdx.x >= 128 && threadIdx.x < 128 + np) srow[threadIdx.x - 128] = patch_row[g0 + threadIdx.x - 128];
  __syncthreads();
  const int total_in = np * runs;
#pragma unroll 3
  for (int i = threadIdx.x; i < total_in; i += 256) {
    const int pi = i / runs, rr = i - pi * runs;
    patch_smem[pi * pitch + rr] = ldg16_stream(proj + static_cast<int64_t>(srow[pi]) * ldp + rr * 8);
  }
  __syncthreads();
  const int total = PATCH_CHUNK * runs;
#pragma unroll 3
  for (int i = threadIdx.x; i < total; i += 256) {
    const int pi = i % PATCH_CHUNK, rr = i / PATCH_CHUNK;
    if (pi < np) stg16(clips + patch_run_offset(sgeom + pi * 4, rr, P0, P1), patch_smem[pi * pitch + rr]);
  }
}

// Expands the per-clip descriptors of a packed batch into the per-row metadata every other kernel consumes
// (TiTokEncoder.forward metadata, blocks.py:72-89; RoPE.forward position ids, rope.py:57-71; patch geometry of
// utils.py:26-51). desc[b] = {row_start, tok_start, pat_start, token_count, n_patches, g1, g2, clip_offset, W, H*W, T*H*W}.
// One thread per packed row; the clip of a row is found by binary search over row_start.
__global__ void __launch_bounds__(256) build_plan_kernel(const int64_t* __restrict__ desc, int B, int64_t M,
                                                         int P0, int P1, int P2, int32_t* __restrict__ enc_src,
                                                         int32_t* __restrict__ dec_src, int32_t* __restrict__ latent_row,
                                                         int32_t* __restrict__ patch_row, int64_t* __restrict__ geom,
                                                         int32_t* __restrict__ rope_pos) {
  for (int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; row < M;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int lo = 0, hi = B - 1;
    while (lo < hi) {  // last clip whose row_start <= row
      const int mid = (lo + hi + 1) >> 1;
      if (desc[mid * 12] <= row) lo = mid; else hi = mid - 1;
    }
    const int64_t* d = desc + lo * 12;
    const int64_t local = row - d[0];
    const int64_t tc = d[3];
    if (local < tc) {
      const int64_t tok = d[1] + local;
      enc_src[row] = -1;
      dec_src[row] = static_cast<int32_t>(tok);
      latent_row[tok] = static_cast<int32_t>(row);
      rope_pos[row * 3 + 0] = rope_pos[row * 3 + 1] = rope_pos[row * 3 + 2] = static_cast<int32_t>(local);
    } else {
      const int64_t pl = local - tc;
      const int64_t pat = d[2] + pl;
      const int64_t g1 = d[5], g2 = d[6];
      const int64_t d0 = pl / (g1 * g2), rem = pl - d0 * g1 * g2, d1 = rem / g2, d2 = rem - d1 * g2;
      enc_src[row] = static_cast<int32_t>(pat);
      dec_src[row] = -1;
      patch_row[pat] = static_cast<int32_t>(row);
      rope_pos[row * 3 + 0] = static_cast<int32_t>(d0 + tc);
      rope_pos[row * 3 + 1] = static_cast<int32_t>(d1 + tc);
      rope_pos[row * 3 + 2] = static_cast<int32_t>(d2 + tc);
      geom[pat * 4 + 0] = d[7] + (d0 * P0) * d[9] + (d1 * P1) * d[8] + d2 * P2;
      geom[pat * 4 + 1] = d[8];
      geom[pat * 4 + 2] = d[9];
      geom[pat * 4 + 3] = d[10];
    }
  }
}

// The same expansion for a BUCKET: the launch extents (M_max rows, T_max tokens, G_max patches) are fixed -- they are
// baked into a captured CUDA graph -- while the real sizes of this step's batch are read from device memory:
//   hdr int64[8] = {n_clips, M, T, G, scratch element offset of a 768-element patch, 0, 0, 0}.
// Rows / tokens / patches past the real sizes get harmless defaults: padded rows are latent-type rows in the encoder and
// patch-type rows in the decoder (mask-token rows, nothing is gathered), padded tokens read packed row 0, padded patches
// read and write the scratch patch behind the clips. Every kernel of the launch sequence is row-wise (attention goes by
// its work list, whose padded records have q_valid = 0), so the padding never touches a real row.
__global__ void __launch_bounds__(256) build_plan_bucket_kernel(const int64_t* __restrict__ desc,
                                                                const int64_t* __restrict__ hdr, int64_t M_max, int64_t T_max,
                                                                int64_t G_max, int P0, int P1, int P2,
                                                                int32_t* __restrict__ enc_src, int32_t* __restrict__ dec_src,
                                                                int32_t* __restrict__ latent_row,
                                                                int32_t* __restrict__ patch_row, int64_t* __restrict__ geom,
                                                                int32_t* __restrict__ rope_pos) {
  const int B = static_cast<int>(hdr[0]);
  const int64_t M = hdr[1], T = hdr[2], G = hdr[3], scratch = hdr[4];
  const int64_t n = M_max > G_max ? (M_max > T_max ? M_max : T_max) : (G_max > T_max ? G_max : T_max);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (i >= T && i < T_max) latent_row[i] = 0;
    if (i >= G && i < G_max) {
      patch_row[i] = 0;
      geom[i * 4 + 0] = scratch;
      geom[i * 4 + 1] = P2;
      geom[i * 4 + 2] = static_cast<int64_t>(P1) * P2;
      geom[i * 4 + 3] = static_cast<int64_t>(P0) * P1 * P2;
    }
    if (i >= M_max) continue;
    const int64_t row = i;
    if (row >= M) {
      enc_src[row] = -1;
      dec_src[row] = -1;
      rope_pos[row * 3 + 0] = rope_pos[row * 3 + 1] = rope_pos[row * 3 + 2] = 0;
      continue;
    }
    int lo = 0, hi = B - 1;
    while (lo < hi) {  // last clip whose row_start <= row
      const int mid = (lo + hi + 1) >> 1;
      if (desc[mid * 12] <= row) lo = mid; else hi = mid - 1;
    }
    const int64_t* d = desc + lo * 12;
    const int64_t local = row - d[0];
    const int64_t tc = d[3];
    if (local < tc) {
      const int64_t tok = d[1] + local;
      enc_src[row] = -1;
      dec_src[row] = static_cast<int32_t>(tok);
      latent_row[tok] = static_cast<int32_t>(row);
      rope_pos[row * 3 + 0] = rope_pos[row * 3 + 1] = rope_pos[row * 3 + 2] = static_cast<int32_t>(local);
    } else {
      const int64_t pl = local - tc;
      const int64_t pat = d[2] + pl;
      const int64_t g1 = d[5], g2 = d[6];
      const int64_t d0 = pl / (g1 * g2), rem = pl - d0 * g1 * g2, d1 = rem / g2, d2 = rem - d1 * g2;
      enc_src[row] = static_cast<int32_t>(pat);
      dec_src[row] = -1;
      patch_row[pat] = static_cast<int32_t>(row);
      rope_pos[row * 3 + 0] = static_cast<int32_t>(d0 + tc);
      rope_pos[row * 3 + 1] = static_cast<int32_t>(d1 + tc);
      rope_pos[row * 3 + 2] = static_cast<int32_t>(d2 + tc);
      geom[pat * 4 + 0] = d[7] + (d0 * P0) * d[9] + (d1 * P1) * d[8] + d2 * P2;
      geom[pat * 4 + 1] = d[8];
      geom[pat * 4 + 2] = d[9];
      geom[pat * 4 + 3] = d[10];
    }
  }
}

// RoPE table of a packed batch: rope[row, (f*3 + a)*2 + {0,1}] = cs_table[pos[row, a], f, {cos, sin}].
// pos: int32 [M,3] integer position ids (RoPE.forward, rope.py:57-71); cs_table: fp32 [n_ids, 10, 2] evaluated once on
// the host in float64 exactly as rope.py:40-54 does. The [M,60] table never crosses PCIe. One thread per complex lane.
__global__ void __launch_bounds__(256) rope_table_gather_kernel(const int32_t* __restrict__ pos,
                                                                const float2* __restrict__ cs_table, int n_ids,
                                                                float2* __restrict__ rope, int64_t M) {
  const int64_t total = M * 30;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / 30;
    const int l = static_cast<int>(i - row * 30);
    const int f = l / 3, a = l - f * 3;
    int id = pos[row * 3 + a];
    id = id < 0 ? 0 : (id >= n_ids ? n_ids - 1 : id);
    rope[i] = __ldg(cs_table + id * 10 + f);
  }
}

// Per-clip reconstruction error: out[2*i] += sum |a - b|, out[2*i+1] += sum (a - b)^2 over clip i (fp64 accumulators
// in global memory, caller-zeroed). The L1 term is the reference's reconstruction loss (loss_module.py:118), the squared
// term feeds PSNR (eval_metrics.py). Bandwidth-bound: 16-byte streaming loads of both buffers, warp-shuffle reduction,
// one pair of fp64 atomics per CTA. blockIdx.y = clip, blockIdx.x strides over the clip's 8-element vectors.
__global__ void __launch_bounds__(256) clip_error_kernel(const __nv_bfloat16* __restrict__ a,
                                                         const __nv_bfloat16* __restrict__ b,
                                                         const int64_t* __restrict__ clip_offset,
                                                         const int64_t* __restrict__ clip_numel, double* __restrict__ out) {
  const int clip = blockIdx.y;
  const int64_t off = clip_offset[clip];
  const int64_t nv = clip_numel[clip] / 8;  // clip sizes are multiples of 8 elements (W % 8 == 0)
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 va = ldg16_stream(a + off + i * 8);
    const uint4 vb = ldg16_stream(b + off + i * 8);
    const uint32_t aa[4] = {va.x, va.y, va.z, va.w}, bb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = bf16_lo(aa[e]) - bf16_lo(bb[e]);
      const float d1 = bf16_hi(aa[e]) - bf16_hi(bb[e]);
      s1 += fabsf(d0) + fabsf(d1);
      s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
    }
  }
  __shared__ float p1[8], p2[8];
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    p1[threadIdx.x >> 5] = s1;
    p2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < 8; ++i) {
      t1 += p1[i];
      t2 += p2[i];
    }
    atomicAdd(out + 2 * clip, t1);
    atomicAdd(out + 2 * clip + 1, t2);
  }
}

// uint8 frames -> normalised bf16 clips, rounding like the reference's dataset code on bf16 tensors
// (dataset/video_dataset.py:118-119): y = bf16(bf16(bf16(u8) / 255) * 2 - 1). 16 pixels per thread (one 16-byte load, two
// 16-byte stores). The 256 possible results are exact images of that torch expression (tests/test_gpu_kernels.py).
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                           int64_t n) {
  const int64_t nv = n / 16;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 v = ldg16_stream(src + i * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float q = bf16r(static_cast<float>((w[k] >> (8 * b)) & 0xffu) / 255.0f);
        f[b] = q * 2.0f - 1.0f;  // q * 2 is exact in bf16; the subtraction is rounded once by the packed convert below
      }
      o[2 * k] = pack_bf16x2(f[0], f[1]);
      o[2 * k + 1] = pack_bf16x2(f[2], f[3]);
    }
    stg16(dst + i * 16, make_uint4(o[0], o[1], o[2], o[3]));
    stg16(dst + i * 16 + 8, make_uint4(o[4], o[5], o[6], o[7]));
  }
  // tail (n % 16 elements)
  if (blockIdx.x == 0 && threadIdx.x < n - nv * 16) {
    const int64_t i = nv * 16 + threadIdx.x;
    const float q = bf16r(static_cast<float>(src[i]) / 255.0f);
    dst[i] = __float2bfloat16_rn(q * 2.0f - 1.0f);
  }
}

int fsq_make_consts(FsqConsts& c, int D, const float* half_l, const float* offset, const float* shift,
                    const float* half_width, const int32_t* basis, const int32_t* levels);

}  // namespace ttk

using namespace ttk;

#define TTK_DISPATCH_NV(width, ...)            \
  switch ((width) / 256) {                     \
    case 1: { constexpr int NV = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int NV = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int NV = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int NV = 4; __VA_ARGS__; break; } \
    default: return TTK_ERR_BAD_SHAPE;         \
  }

static inline int row_grid(int M) { return (M + ROW_WARPS - 1) / ROW_WARPS; }
static inline bool width_ok(int width) { return width > 0 && width % 256 == 0 && width <= 1024; }

extern "C" {

int ttk_rmsnorm_fwd(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int M, int width,
                    cudaStream_t stream) {
  if (!x || !w || !y) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!width_ok(width) || ldx % 8 || ldy % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  TTK_DISPATCH_NV(width, rmsnorm_kernel<NV><<<row_grid(M), ROW_WARPS * 32, 0, stream>>>(
                             static_cast<const __nv_bfloat16*>(x), ldx, w, static_cast<__nv_bfloat16*>(y), ldy, M));
  return launch_status();
}

int ttk_resid_norm(const void* x, const void* y, void* x_out, void* xn_out, const float* w_post, const float* w_next,
                   float alpha, int mode, int M, int width, int64_t ld, cudaStream_t stream) {
  if (!x || !y || !x_out) return TTK_ERR_BAD_ARG;
  if ((mode == 1 && !w_post) || (xn_out && !w_next)) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!width_ok(width) || ld % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  TTK_DISPATCH_NV(width, resid_norm_kernel<NV><<<row_grid(M), ROW_WARPS * 32, 0, stream>>>(
                             static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(y),
                             static_cast<__nv_bfloat16*>(x_out), static_cast<__nv_bfloat16*>(xn_out), w_post, w_next,
                             alpha, mode, M, ld));
  return launch_status();
}

static int enc_embed_launch(const void* proj, int64_t ldp, const int32_t* src_row, const float* mask_token,
                            const float* w_t, const float* w_p, const float* w_next, void* x_out, void* xn_out,
                            void* e0_out, int M, int width, int64_t ld, cudaStream_t stream) {
  if (!proj || !src_row || !mask_token || !w_t || !w_p || !x_out) return TTK_ERR_BAD_ARG;
  if (xn_out && !w_next) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!width_ok(width) || ld % 8 || ldp % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  TTK_DISPATCH_NV(width, enc_embed_kernel<NV><<<row_grid(M), ROW_WARPS * 32, 0, stream>>>(
                             static_cast<const __nv_bfloat16*>(proj), ldp, src_row, mask_token, w_t, w_p, w_next,
                             static_cast<__nv_bfloat16*>(x_out), static_cast<__nv_bfloat16*>(xn_out),
                             static_cast<__nv_bfloat16*>(e0_out), M, ld));
  return launch_status();
}

int ttk_enc_embed(const void* proj, int64_t ldp, const int32_t* src_row, const float* mask_token, const float* w_t,
                  const float* w_p, const float* w_next, void* x_out, void* xn_out, int M, int width, int64_t ld,
                  cudaStream_t stream) {
  return enc_embed_launch(proj, ldp, src_row, mask_token, w_t, w_p, w_next, x_out, xn_out, nullptr, M, width, ld, stream);
}

// training variant: e0_out [M, ld] additionally receives the rows BEFORE ln_pre_t / ln_pre_p
int ttk_enc_embed_train(const void* proj, int64_t ldp, const int32_t* src_row, const float* mask_token, const float* w_t,
                        const float* w_p, const float* w_next, void* x_out, void* xn_out, void* e0_out, int M, int width,
                        int64_t ld, cudaStream_t stream) {
  if (!e0_out) return TTK_ERR_BAD_ARG;
  return enc_embed_launch(proj, ldp, src_row, mask_token, w_t, w_p, w_next, x_out, xn_out, e0_out, M, width, ld, stream);
}

static int dec_embed_launch(const void* codes, int token_size, const int32_t* src_row, const void* w_in, const void* b_in,
                            const float* mask_token, const float* w_t, const float* w_p, const float* w_next,
                            void* x_out, void* xn_out, void* e0_out, int M, int width, int64_t ld, cudaStream_t stream) {
  if (!codes || !src_row || !w_in || !b_in || !mask_token || !w_t || !w_p || !x_out) return TTK_ERR_BAD_ARG;
  if (xn_out && !w_next) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!width_ok(width) || ld % 8 || token_size < 1 || token_size > FSQ_MAX_D) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  TTK_DISPATCH_NV(width, dec_embed_kernel<NV><<<row_grid(M), ROW_WARPS * 32, 0, stream>>>(
                             static_cast<const __nv_bfloat16*>(codes), token_size, src_row,
                             static_cast<const __nv_bfloat16*>(w_in), static_cast<const __nv_bfloat16*>(b_in),
                             mask_token, w_t, w_p, w_next, static_cast<__nv_bfloat16*>(x_out),
                             static_cast<__nv_bfloat16*>(xn_out), static_cast<__nv_bfloat16*>(e0_out), M, ld));
  return launch_status();
}

int ttk_dec_embed(const void* codes, int token_size, const int32_t* src_row, const void* w_in, const void* b_in,
                  const float* mask_token, const float* w_t, const float* w_p, const float* w_next, void* x_out,
                  void* xn_out, int M, int width, int64_t ld, cudaStream_t stream) {
  return dec_embed_launch(codes, token_size, src_row, w_in, b_in, mask_token, w_t, w_p, w_next, x_out, xn_out, nullptr, M,
                          width, ld, stream);
}

int ttk_dec_embed_train(const void* codes, int token_size, const int32_t* src_row, const void* w_in, const void* b_in,
                        const float* mask_token, const float* w_t, const float* w_p, const float* w_next, void* x_out,
                        void* xn_out, void* e0_out, int M, int width, int64_t ld, cudaStream_t stream) {
  if (!e0_out) return TTK_ERR_BAD_ARG;
  return dec_embed_launch(codes, token_size, src_row, w_in, b_in, mask_token, w_t, w_p, w_next, x_out, xn_out, e0_out, M,
                          width, ld, stream);
}

int ttk_enc_head_fsq(const void* x, int64_t ld, const int32_t* latent_row, const float* w_post, int pre_normed,
                     const void* w_out, const void* b_out, int token_size, void* z_out, void* codes_out,
                     int32_t* idx_out, int T, int width, const float* half_l, const float* offset, const float* shift,
                     const float* half_width, const int32_t* basis, const int32_t* levels, cudaStream_t stream) {
  if (!x || !w_out || !b_out || !z_out || !codes_out || !idx_out) return TTK_ERR_BAD_ARG;
  if (!pre_normed && !w_post) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!width_ok(width) || ld % 8) return TTK_ERR_BAD_SHAPE;
  FsqConsts c;
  if (int e = fsq_make_consts(c, token_size, half_l, offset, shift, half_width, basis, levels)) return e;
  if (T <= 0) return TTK_OK;
  TTK_DISPATCH_NV(width, enc_head_fsq_kernel<NV><<<row_grid(T), ROW_WARPS * 32, 0, stream>>>(
                             static_cast<const __nv_bfloat16*>(x), ld, latent_row, w_post, pre_normed,
                             static_cast<const __nv_bfloat16*>(w_out), static_cast<const __nv_bfloat16*>(b_out),
                             token_size, static_cast<__nv_bfloat16*>(z_out), static_cast<__nv_bfloat16*>(codes_out),
                             idx_out, T, c));
  return launch_status();
}

// geom: device int64 [G,4] = {offset, W, H*W, T*H*W} per patch. Requires P2 == 8 (16-byte runs).
int ttk_patchify(const void* clips, const int64_t* geom, int C, int P0, int P1, int P2, void* patches, int64_t ldp,
                 int64_t G, cudaStream_t stream) {
  if (!clips || !geom || !patches) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (P2 != 8 || C < 1 || P0 < 1 || P1 < 1 || ldp % 8) return TTK_ERR_BAD_SHAPE;
  if (G <= 0) return TTK_OK;
  const size_t smem = static_cast<size_t>(PATCH_CHUNK) * (C * P0 * P1 + 1) * 16;
  if (smem > 48 * 1024) return TTK_ERR_BAD_SHAPE;
  const int64_t blocks = (G + PATCH_CHUNK - 1) / PATCH_CHUNK;
  if (blocks > 0x7fffffffLL) return TTK_ERR_BAD_SHAPE;
  patchify_kernel<false><<<static_cast<int>(blocks), 256, smem, stream>>>(clips, geom, C, P0, P1,
                                                                        static_cast<__nv_bfloat16*>(patches), ldp, G);
  return launch_status();
}

// ttk_patchify on decoded uint8 frames (the reference's dataset output before `chunk.to(dtype) / 255; chunk * 2 - 1`,
// dataset/video_dataset.py:118-119): gathers the patches and normalises them in one pass. clips: uint8, every clip
// starting on an 8-byte boundary; patches: bf16 [G, ldp], bit-identical to ttk_normalize_u8 followed by ttk_patchify.
int ttk_patchify_u8(const void* clips, const int64_t* geom, int C, int P0, int P1, int P2, void* patches, int64_t ldp,
                    int64_t G, cudaStream_t stream) {
  if (!clips || !geom || !patches) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (P2 != 8 || C < 1 || P0 < 1 || P1 < 1 || ldp % 8) return TTK_ERR_BAD_SHAPE;
  if (reinterpret_cast<uintptr_t>(clips) & 7u) return TTK_ERR_ALIGNMENT;
  if (G <= 0) return TTK_OK;
  const size_t smem = static_cast<size_t>(PATCH_CHUNK) * (C * P0 * P1 + 1) * 16;
  if (smem > 48 * 1024) return TTK_ERR_BAD_SHAPE;
  const int64_t blocks = (G + PATCH_CHUNK - 1) / PATCH_CHUNK;
  if (blocks > 0x7fffffffLL) return TTK_ERR_BAD_SHAPE;
  patchify_kernel<true><<<static_cast<int>(blocks), 256, smem, stream>>>(clips, geom, C, P0, P1,
                                                                       static_cast<__nv_bfloat16*>(patches), ldp, G);
  return launch_status();
}

int ttk_unpatchify(const void* proj, int64_t ldp, const int32_t* patch_row, const int64_t* geom, int C, int P0, int P1,
                   int P2, void* clips, int64_t G, cudaStream_t stream) {
  if (!proj || !patch_row || !geom || !clips) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (P2 != 8 || C < 1 || P0 < 1 || P1 < 1 || ldp % 8) return TTK_ERR_BAD_SHAPE;
  if (G <= 0) return TTK_OK;
  const size_t smem = static_cast<size_t>(PATCH_CHUNK) * (C * P0 * P1 + 1) * 16;
  if (smem > 48 * 1024) return TTK_ERR_BAD_SHAPE;
  const int64_t blocks = (G + PATCH_CHUNK - 1) / PATCH_CHUNK;
  if (blocks > 0x7fffffffLL) return TTK_ERR_BAD_SHAPE;
  unpatchify_kernel<<<static_cast<int>(blocks), 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(proj), ldp, patch_row, geom, C, P0,
                                              P1, static_cast<__nv_bfloat16*>(clips), G);
  return launch_status();
}

// out: device fp64 [2 * n_clips], caller-zeroed; clip_offset / clip_numel: device int64 [n_clips] (elements).
int ttk_clip_error(const void* a, const void* b, const int64_t* clip_offset, const int64_t* clip_numel, int n_clips,
                   int64_t max_clip_numel, double* out, cudaStream_t stream) {
  if (n_clips <= 0) return TTK_OK;
  if (!a || !b || !clip_offset || !clip_numel || !out) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) != 0) return TTK_ERR_ALIGNMENT;
  if (n_clips > 65535 || max_clip_numel <= 0) return TTK_ERR_BAD_SHAPE;
  int64_t bx = (max_clip_numel / 8 + 256 * 8 - 1) / (256 * 8);  // about 8 vectors per thread
  if (bx < 1) bx = 1;
  if (bx > 4LL * num_sms()) bx = 4LL * num_sms();
  clip_error_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n_clips)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), clip_offset, clip_numel, out);
  return launch_status();
}

// dst bf16 [n] = normalised src uint8 [n] (both 16-byte aligned).
int ttk_normalize_u8(const void* src, void* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return TTK_OK;
  if (!src || !dst) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) != 0) return TTK_ERR_ALIGNMENT;
  int64_t blocks = (n / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  normalize_u8_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const uint8_t*>(src),
                                                                     static_cast<__nv_bfloat16*>(dst), n);
  return launch_status();
}

// rope [M,60] fp32 <- gather of cs_table [n_ids,10,2] fp32 by pos [M,3] int32 (see rope_table_gather_kernel).
int ttk_rope_table_gather(const int32_t* pos, const float* cs_table, int n_ids, float* rope, int64_t M,
                          cudaStream_t stream) {
  if (M <= 0) return TTK_OK;
  if (!pos || !cs_table || !rope || n_ids <= 0) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (((reinterpret_cast<uintptr_t>(cs_table) | reinterpret_cast<uintptr_t>(rope)) & 7u) != 0) return TTK_ERR_ALIGNMENT;
  const int64_t blocks = (M * 30 + 255) / 256;
  const int grid = static_cast<int>(blocks < 16LL * num_sms() ? blocks : 16LL * num_sms());
  rope_table_gather_kernel<<<grid, 256, 0, stream>>>(pos, reinterpret_cast<const float2*>(cs_table), n_ids,
                                                     reinterpret_cast<float2*>(rope), M);
  return launch_status();
}

// Per-row metadata of a packed batch from per-clip descriptors (see build_plan_kernel). All pointers are device memory.
int ttk_build_plan(const int64_t* desc, int n_clips, int64_t M, int P0, int P1, int P2, int32_t* enc_src_row,
                   int32_t* dec_src_row, int32_t* latent_row, int32_t* patch_row, int64_t* geom, int32_t* rope_pos,
                   cudaStream_t stream) {
  if (M <= 0 || n_clips <= 0) return TTK_OK;
  if (!desc || !enc_src_row || !dec_src_row || !latent_row || !patch_row || !geom || !rope_pos) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  const int64_t blocks = (M + 255) / 256;
  const int grid = static_cast<int>(blocks < 16LL * num_sms() ? blocks : 16LL * num_sms());
  build_plan_kernel<<<grid, 256, 0, stream>>>(desc, n_clips, M, P0, P1, P2, enc_src_row, dec_src_row, latent_row, patch_row,
                                              geom, rope_pos);
  return launch_status();
}

// ttk_build_plan for a bucket of batch compositions (see build_plan_bucket_kernel): fixed extents by value, the real sizes
// of the step in device memory (hdr), so that ONE captured CUDA graph serves every composition of the bucket.
int ttk_build_plan_bucket(const int64_t* desc, const int64_t* hdr, int64_t M_max, int64_t T_max, int64_t G_max, int P0,
                          int P1, int P2, int32_t* enc_src_row, int32_t* dec_src_row, int32_t* latent_row,
                          int32_t* patch_row, int64_t* geom, int32_t* rope_pos, cudaStream_t stream) {
  if (M_max <= 0) return TTK_OK;
  if (!desc || !hdr || !enc_src_row || !dec_src_row || !latent_row || !patch_row || !geom || !rope_pos) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (T_max < 0 || G_max < 0 || P0 < 1 || P1 < 1 || P2 < 1) return TTK_ERR_BAD_SHAPE;
  const int64_t n = M_max > G_max ? (M_max > T_max ? M_max : T_max) : (G_max > T_max ? G_max : T_max);
  const int64_t blocks = (n + 255) / 256;
  const int grid = static_cast<int>(blocks < 16LL * num_sms() ? blocks : 16LL * num_sms());
  build_plan_bucket_kernel<<<grid, 256, 0, stream>>>(desc, hdr, M_max, T_max, G_max, P0, P1, P2, enc_src_row, dec_src_row,
                                                     latent_row, patch_row, geom, rope_pos);
  return launch_status();
}

}  // extern "C"
