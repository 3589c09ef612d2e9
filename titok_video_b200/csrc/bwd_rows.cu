// Bandwidth-bound row kernels of the training path (backward of the encoder / decoder stacks).
//
//   rmsnorm_bwd        backward of flash-attn's RMSNorm (fa:ops/triton/layer_norm.py:1093-1126; bwd kernel :390-520),
//                      optionally through the KEEL pre-sum u = alpha*x + y (transformer.py:141-145) and with the
//                      residual-path gradient added:  dx = add_scale * add + d(norm)/du
//   geglu_fwd / bwd    gelu(gate) * value of GEGLU.forward (transformer.py:47-52) on a stored w12 output
//   gather / scatter   row gathers by the packing maps (latent rows / patch rows, blocks.py:85-86,101,174)
//   colsum             bias gradients (column sums) and the mask_token gradient (sum of everything)
//   head_bwd           backward of the encoder head Linear(width -> token_size) on the latent rows (blocks.py:101-103)
//   dec_in_bwd         backward of the decoder's Linear(token_size -> width) on the latent rows (blocks.py:164-165)
// One warp per packed row, 16-byte vector access, fp32 math, bf16 gradients for bf16 activations (what autograd
// produces under the reference's bf16 autocast), fp32 gradients for parameters.
#include "common.cuh"
#include "host_util.cuh"

namespace ttk {

constexpr float RMS_EPS_B = 1e-5f;
constexpr int BW_WARPS = 8;

template <int NV>
struct Row {
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
};

// Raw (still packed) row: lets a warp issue the loads of its NEXT row before it works on the current one, so that two
// rows per warp are in flight (the row kernels are latency-bound at 32 resident warps per SM otherwise).
template <int NV>
struct RawRow {
  uint4 q[NV];
  __device__ __forceinline__ void load(const __nv_bfloat16* row, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) q[i] = ldg16(row + i * 256 + lane * 8);
  }
  __device__ __forceinline__ void unpack(Row<NV>& r) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t rr[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r.v[i * 8 + 2 * e] = bf16_lo(rr[e]);
        r.v[i * 8 + 2 * e + 1] = bf16_hi(rr[e]);
      }
    }
  }
};

template <int NV>
__device__ __forceinline__ void load_w(const float* w, int lane, float (&out)[NV * 8]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(w + i * 256 + lane * 8);
    const float4 b = *reinterpret_cast<const float4*>(w + i * 256 + lane * 8 + 4);
    out[i * 8 + 0] = a.x; out[i * 8 + 1] = a.y; out[i * 8 + 2] = a.z; out[i * 8 + 3] = a.w;
    out[i * 8 + 4] = b.x; out[i * 8 + 5] = b.y; out[i * 8 + 6] = b.z; out[i * 8 + 7] = b.w;
  }
}

// ------------------------------------------------------------------------------------------------
// RMSNorm backward.  u = (y ? bf16(bf16(alpha*x) + y) : x);  rstd = 1/sqrt(mean(u^2)+eps);  xhat = u*rstd
//   dxhat = dy * w;  du = rstd * (dxhat - xhat * mean(dxhat * xhat));  dx = du + add_scale * add
//   dw += sum_rows dy * xhat      (rows with sel[row] < 0 use w2 / dw2 instead: the two pre-norms of the embed)
// ------------------------------------------------------------------------------------------------
template <int NV, bool SEL>
__global__ void __launch_bounds__(BW_WARPS * 32)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y, float alpha,
                   const float* __restrict__ w, const float* __restrict__ w2, const int32_t* __restrict__ sel,
                   const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ add, float add_scale,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dw2, int M, int64_t ld) {
  constexpr int W = NV * 256;
  __shared__ float red[BW_WARPS][SEL ? 2 : 1][256];  // per-warp partial dw of one 256-column vector at a time
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  float wv[NV * 8], wv2[SEL ? NV * 8 : 1];
  float acc[NV * 8], acc2[SEL ? NV * 8 : 1];
  load_w<NV>(w, lane, wv);
  if constexpr (SEL) load_w<NV>(w2, lane, wv2);
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) acc[i] = 0.f;
  if constexpr (SEL) {
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) acc2[i] = 0.f;
  }
  const int stride = gridDim.x * BW_WARPS;
  int row = blockIdx.x * BW_WARPS + wid;
  RawRow<NV> nx, ny, ng, na;
  if (row < M) {
    nx.load(x + row * ld, lane);
    if (y) ny.load(y + row * ld, lane);
    ng.load(dy + row * ld, lane);
    if (add) na.load(add + row * ld, lane);
  }
  for (; row < M; row += stride) {
    Row<NV> u, g, a;
    nx.unpack(u);
    if (y) {
      Row<NV> b;
      ny.unpack(b);
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) u.v[i] = bf16r(bf16r(u.v[i] * alpha) + b.v[i]);
    }
    ng.unpack(g);
    if (add) na.unpack(a);
    if (row + stride < M) {  // next row's loads go out before this row's reductions
      const int64_t nr = row + stride;
      nx.load(x + nr * ld, lane);
      if (y) ny.load(y + nr * ld, lane);
      ng.load(dy + nr * ld, lane);
      if (add) na.load(add + nr * ld, lane);
    }
    bool second = false;
    if constexpr (SEL) second = sel[row] < 0;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) ss += u.v[i] * u.v[i];
    ss = warp_sum(ss);
    const float rstd = 1.0f / sqrtf(ss / static_cast<float>(W) + RMS_EPS_B);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      const float xh = u.v[i] * rstd;
      float wi = wv[i];
      if constexpr (SEL) wi = second ? wv2[i] : wi;
      const float dxh = g.v[i] * wi;
      dot = fmaf(dxh, xh, dot);
      if constexpr (SEL) {
        if (second) acc2[i] = fmaf(g.v[i], xh, acc2[i]);
        else acc[i] = fmaf(g.v[i], xh, acc[i]);
      } else {
        acc[i] = fmaf(g.v[i], xh, acc[i]);
      }
      u.v[i] = xh;
      g.v[i] = dxh;
    }
    dot = warp_sum(dot) / static_cast<float>(W);
    if (add) {
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) g.v[i] = fmaf(add_scale, a.v[i], rstd * (g.v[i] - u.v[i] * dot));
    } else {
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) g.v[i] = rstd * (g.v[i] - u.v[i] * dot);
    }
    g.store(dx + row * ld, lane);
  }
  // dw: reduce the 8 warps of the CTA through shared memory, one atomic per column per CTA
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[wid][0][lane * 8 + e] = acc[i * 8 + e];
      if constexpr (SEL) red[wid][1][lane * 8 + e] = acc2[i * 8 + e];
    }
    __syncthreads();
    {
      const int col = threadIdx.x;  // 256 threads == 256 columns
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < BW_WARPS; ++k) {
        s += red[k][0][col];
        if constexpr (SEL) s2 += red[k][1][col];
      }
      if (dw && s != 0.f) atomicAdd(dw + i * 256 + col, s);
      if constexpr (SEL) {
        if (dw2 && s2 != 0.f) atomicAdd(dw2 + i * 256 + col, s2);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// GEGLU on a stored w12 output h12 [M, 2*inner] = [value | gate] (transformer.py:50-52)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_exact(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float g) {
  const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752f));
  return cdf + g * 0.3989422804014327f * __expf(-0.5f * g * g);
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h12, int64_t ld12, int inner,
                                                        __nv_bfloat16* __restrict__ h, int64_t ldh, int64_t M) {
  const int vpr = inner / 8;
  const int64_t total = M * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vpr;
    const int col = static_cast<int>(i - row * vpr) * 8;
    const uint4 xv = ldg16(h12 + row * ld12 + col);
    const uint4 gv = ldg16(h12 + row * ld12 + inner + col);
    const uint32_t xx[4] = {xv.x, xv.y, xv.z, xv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack_bf16x2(bf16r(gelu_exact(bf16_lo(gg[e]))) * bf16_lo(xx[e]), bf16r(gelu_exact(bf16_hi(gg[e]))) * bf16_hi(xx[e]));
    stg16(h + row * ldh + col, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// dvalue = bf16(dh * bf16(gelu(gate)));  dgate = bf16(bf16(dh * value) * gelu'(gate))
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h12, int64_t ld12, int inner,
                                                        const __nv_bfloat16* __restrict__ dh, int64_t ldh,
                                                        __nv_bfloat16* __restrict__ dh12, int64_t ldd, int64_t M) {
  const int vpr = inner / 8;
  const int64_t total = M * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vpr;
    const int col = static_cast<int>(i - row * vpr) * 8;
    const uint4 xv = ldg16(h12 + row * ld12 + col);
    const uint4 gv = ldg16(h12 + row * ld12 + inner + col);
    const uint4 dv = ldg16(dh + row * ldh + col);
    const uint32_t xx[4] = {xv.x, xv.y, xv.z, xv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t ox[4], og[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g0 = bf16_lo(gg[e]), g1 = bf16_hi(gg[e]);
      const float d0 = bf16_lo(dd[e]), d1 = bf16_hi(dd[e]);
      ox[e] = pack_bf16x2(d0 * bf16r(gelu_exact(g0)), d1 * bf16r(gelu_exact(g1)));
      og[e] = pack_bf16x2(bf16r(d0 * bf16_lo(xx[e])) * gelu_grad(g0), bf16r(d1 * bf16_hi(xx[e])) * gelu_grad(g1));
    }
    stg16(dh12 + row * ldd + col, make_uint4(ox[0], ox[1], ox[2], ox[3]));
    stg16(dh12 + row * ldd + inner + col, make_uint4(og[0], og[1], og[2], og[3]));
  }
}

// ------------------------------------------------------------------------------------------------
// dst[i] = src[idx[i]]   /   dst[idx[i]] = src[i]     (rows of `width` bf16, width % 8 == 0)
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) move_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t lds,
                                                        const int32_t* __restrict__ idx, __nv_bfloat16* __restrict__ dst,
                                                        int64_t ldd, int64_t n, int width) {
  const int vpr = width / 8;
  const int64_t total = n * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vpr;
    const int col = static_cast<int>(i - row * vpr) * 8;
    const int64_t other = idx[row];
    if (SCATTER) stg16(dst + other * ldd + col, ldg16(src + row * lds + col));
    else stg16(dst + row * ldd + col, ldg16(src + other * lds + col));
  }
}

// ------------------------------------------------------------------------------------------------
// out[c] += sum_r x[r, c] (fp32), total[0] += sum of everything. Block = 32 column groups (8 columns, one 16-byte load
// each) x 8 row lanes; a CTA covers 128 rows x 256 columns (16 rows per thread in flight), reduces its row lanes
// through shared memory and issues one atomic per column.
// ------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int64_t M, int N,
                                                     float* __restrict__ out, float* __restrict__ total) {
  __shared__ float red[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.y * 32 + cx) * 8;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * CS_ROWS;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (col < N) {
#pragma unroll 4
    for (int i = 0; i < CS_ROWS / 8; ++i) {
      const int64_t r = r0 + ry + i * 8;
      if (r < M) {
        const uint4 v = ldg16_stream(x + r * ld + col);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += bf16_lo(w[e]);
          acc[2 * e + 1] += bf16_hi(w[e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ry][cx][e] = acc[e];
  __syncthreads();
  float tsum = 0.f;
  {
    // thread t sums column (t % 8) of column group (t / 8) over the 8 row lanes
    const int g = threadIdx.x >> 3, e = threadIdx.x & 7;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][g][e];
    const int c = (blockIdx.y * 32 + g) * 8 + e;
    if (c < N) {
      if (out) atomicAdd(out + c, s);
      tsum = s;
    }
  }
  if (total) {
    __shared__ float part[8];
    tsum = warp_sum(tsum);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = tsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += part[i];
      atomicAdd(total, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder head backward. z[t] = xn[latent_row[t]] @ Wout^T + b  (Wout bf16 [TS, W]).
//   dxn[latent_row[t]] = bf16(sum_k dz[t,k] * Wout[k,:])     (dxn is zero elsewhere: caller clears it)
//   dWout[k,:] += sum_t dz[t,k] * xn[latent_row[t],:]         db[k] += sum_t dz[t,k]
// blockIdx.y = k (the CTA accumulates dWout[k,:]); the k == 0 CTAs also write dxn.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(BW_WARPS * 32)
head_bwd_kernel(const __nv_bfloat16* __restrict__ dz, int TS, const __nv_bfloat16* __restrict__ xn, int64_t ld,
                const int32_t* __restrict__ latent_row, const __nv_bfloat16* __restrict__ w_out,
                __nv_bfloat16* __restrict__ dxn, float* __restrict__ dw, float* __restrict__ db, int T) {
  __shared__ float red[BW_WARPS][256];
  __shared__ float redb[BW_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = blockIdx.y;
  float acc[NV * 8];
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  for (int t = blockIdx.x * BW_WARPS + wid; t < T; t += gridDim.x * BW_WARPS) {
    const int row = latent_row[t];
    Row<NV> a;
    a.load(xn + row * ld, lane);
    const float dzk = __bfloat162float(dz[static_cast<int64_t>(t) * TS + k]);
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) acc[i] = fmaf(dzk, a.v[i], acc[i]);
    bsum += dzk;
    if (k == 0) {
      Row<NV> o;
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) o.v[i] = 0.f;
      for (int kk = 0; kk < TS; ++kk) {
        const float d = __bfloat162float(dz[static_cast<int64_t>(t) * TS + kk]);
        Row<NV> wr;
        wr.load(w_out + static_cast<int64_t>(kk) * (NV * 256), lane);
#pragma unroll
        for (int i = 0; i < NV * 8; ++i) o.v[i] = fmaf(d, wr.v[i], o.v[i]);
      }
      o.store(dxn + row * ld, lane);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[wid][lane * 8 + e] = acc[i * 8 + e];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < BW_WARPS; ++q) s += red[q][threadIdx.x];
    atomicAdd(dw + static_cast<int64_t>(k) * (NV * 256) + i * 256 + threadIdx.x, s);
    __syncthreads();
  }
  if (lane == 0) redb[wid] = bsum;  // every lane of a warp holds the same bsum
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int q = 0; q < BW_WARPS; ++q) s += redb[q];
    atomicAdd(db + k, s);
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder proj_in backward. e[latent_row[t]] = codes[t] @ Win^T + b + mask_token  (Win bf16 [W, TS]).
//   dcodes[t,k] = sum_n de[row_t, n] * Win[n,k]      dWin[n,k] += sum_t de[row_t,n] * codes[t,k]
//   db[n] += sum_t de[row_t,n]                        blockIdx.y = k
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(BW_WARPS * 32)
dec_in_bwd_kernel(const __nv_bfloat16* __restrict__ de, int64_t ld, const int32_t* __restrict__ latent_row,
                  const __nv_bfloat16* __restrict__ codes, int TS, const __nv_bfloat16* __restrict__ w_in,
                  float* __restrict__ dcodes, float* __restrict__ dw, float* __restrict__ db, int T) {
  __shared__ float red[BW_WARPS][2][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = blockIdx.y;
  float acc[NV * 8], accb[NV * 8], wk[NV * 8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[i * 8 + e] = 0.f;
      accb[i * 8 + e] = 0.f;
      wk[i * 8 + e] = __bfloat162float(w_in[static_cast<int64_t>(i * 256 + lane * 8 + e) * TS + k]);
    }
  for (int t = blockIdx.x * BW_WARPS + wid; t < T; t += gridDim.x * BW_WARPS) {
    const int row = latent_row[t];
    Row<NV> g;
    g.load(de + row * ld, lane);
    const float ck = __bfloat162float(codes[static_cast<int64_t>(t) * TS + k]);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      acc[i] = fmaf(g.v[i], ck, acc[i]);
      accb[i] += g.v[i];
      dot = fmaf(g.v[i], wk[i], dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) dcodes[static_cast<int64_t>(t) * TS + k] = dot;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[wid][0][lane * 8 + e] = acc[i * 8 + e];
      red[wid][1][lane * 8 + e] = accb[i * 8 + e];
    }
    __syncthreads();
    float s = 0.f, sb = 0.f;
#pragma unroll
    for (int q = 0; q < BW_WARPS; ++q) {
      s += red[q][0][threadIdx.x];
      sb += red[q][1][threadIdx.x];
    }
    const int n = i * 256 + threadIdx.x;
    atomicAdd(dw + static_cast<int64_t>(n) * TS + k, s);
    if (k == 0) atomicAdd(db + n, sb);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Weight refresh: the fp32 (or bf16) master parameters of a stack -> their kernel-layout copies (bf16 GEMM operands, fp32
// norm weights) in ONE launch. table: device int64 [n][4] = {src pointer, dst pointer, numel, kind}, kind = 2 * (src is
// bf16) + (dst is fp32). blockIdx.y = tensor, blockIdx.x strides over its elements.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) multi_cast_kernel(const int64_t* __restrict__ table) {
  const int64_t* e = table + static_cast<int64_t>(blockIdx.y) * 4;
  const int64_t n = e[2];
  const int kind = static_cast<int>(e[3]);
  const void* src = reinterpret_cast<const void*>(static_cast<uintptr_t>(e[0]));
  void* dst = reinterpret_cast<void*>(static_cast<uintptr_t>(e[1]));
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = (kind & 2) ? __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]) : static_cast<const float*>(src)[i];
    if (kind & 1) static_cast<float*>(dst)[i] = v;
    else static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
  }
}

static inline bool bw_width_ok(int width) { return width > 0 && width % 256 == 0 && width <= 1024; }

}  // namespace ttk

using namespace ttk;

#define TTK_BW_NV(width, ...)                              \
  switch ((width) / 256) {                                 \
    case 1: { constexpr int NV = 1; __VA_ARGS__; break; }  \
    case 2: { constexpr int NV = 2; __VA_ARGS__; break; }  \
    case 3: { constexpr int NV = 3; __VA_ARGS__; break; }  \
    case 4: { constexpr int NV = 4; __VA_ARGS__; break; }  \
    default: return TTK_ERR_BAD_SHAPE;                     \
  }

extern "C" {

int ttk_rmsnorm_bwd(const void* x, const void* y, float alpha, const float* w, const float* w2, const int32_t* sel,
                    const void* dy, const void* add, float add_scale, void* dx, float* dw, float* dw2, int M, int width,
                    int64_t ld, cudaStream_t stream) {
  if (!x || !w || !dy || !dx) return TTK_ERR_BAD_ARG;
  if (sel && !w2) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!bw_width_ok(width) || ld % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  int grid = (M + BW_WARPS - 1) / BW_WARPS;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(y);
  const __nv_bfloat16* dyb = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* ab = static_cast<const __nv_bfloat16*>(add);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx);
  if (sel) {
    TTK_BW_NV(width, rmsnorm_bwd_kernel<NV, true><<<grid, BW_WARPS * 32, 0, stream>>>(xb, yb, alpha, w, w2, sel, dyb, ab,
                                                                                      add_scale, dxb, dw, dw2, M, ld));
  } else {
    TTK_BW_NV(width, rmsnorm_bwd_kernel<NV, false><<<grid, BW_WARPS * 32, 0, stream>>>(xb, yb, alpha, w, w2, sel, dyb, ab,
                                                                                       add_scale, dxb, dw, dw2, M, ld));
  }
  return launch_status();
}

static int ew_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = 16LL * num_sms();
  return static_cast<int>(b < cap ? (b < 1 ? 1 : b) : cap);
}

int ttk_geglu_fwd(const void* h12, int64_t ld12, int inner, void* h, int64_t ldh, int64_t M, cudaStream_t stream) {
  if (!h12 || !h) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (inner <= 0 || inner % 8 || ld12 % 8 || ldh % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  geglu_fwd_kernel<<<ew_grid(M * (inner / 8)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(h12), ld12, inner,
                                                                 static_cast<__nv_bfloat16*>(h), ldh, M);
  return launch_status();
}

int ttk_geglu_bwd(const void* h12, int64_t ld12, int inner, const void* dh, int64_t ldh, void* dh12, int64_t ldd,
                  int64_t M, cudaStream_t stream) {
  if (!h12 || !dh || !dh12) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (inner <= 0 || inner % 8 || ld12 % 8 || ldh % 8 || ldd % 8) return TTK_ERR_BAD_SHAPE;
  if (M <= 0) return TTK_OK;
  geglu_bwd_kernel<<<ew_grid(M * (inner / 8)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(h12), ld12, inner,
                                                                 static_cast<const __nv_bfloat16*>(dh), ldh,
                                                                 static_cast<__nv_bfloat16*>(dh12), ldd, M);
  return launch_status();
}

int ttk_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int64_t n, int width,
                    cudaStream_t stream) {
  if (!src || !idx || !dst) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width <= 0 || width % 8 || lds % 8 || ldd % 8) return TTK_ERR_BAD_SHAPE;
  if (n <= 0) return TTK_OK;
  move_rows_kernel<false><<<ew_grid(n * (width / 8)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds, idx,
                                                                        static_cast<__nv_bfloat16*>(dst), ldd, n, width);
  return launch_status();
}

int ttk_scatter_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int64_t n, int width,
                     cudaStream_t stream) {
  if (!src || !idx || !dst) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (width <= 0 || width % 8 || lds % 8 || ldd % 8) return TTK_ERR_BAD_SHAPE;
  if (n <= 0) return TTK_OK;
  move_rows_kernel<true><<<ew_grid(n * (width / 8)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds, idx,
                                                                       static_cast<__nv_bfloat16*>(dst), ldd, n, width);
  return launch_status();
}

int ttk_colsum(const void* x, int64_t ld, int64_t M, int N, float* out, float* total, cudaStream_t stream) {
  if (!x || (!out && !total)) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (N <= 0 || N % 8 || ld % 8) return TTK_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return TTK_ERR_ALIGNMENT;
  if (M <= 0) return TTK_OK;
  const int64_t blocks = (M + CS_ROWS - 1) / CS_ROWS;
  if (blocks > 0x7fffffffLL) return TTK_ERR_BAD_SHAPE;
  colsum_kernel<<<dim3(static_cast<unsigned>(blocks), static_cast<unsigned>((N + 255) / 256)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ld, M, N, out, total);
  return launch_status();
}

int ttk_multi_cast(const int64_t* table, int n, int64_t max_numel, cudaStream_t stream) {
  if (n <= 0) return TTK_OK;
  if (!table || max_numel <= 0) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (n > 65535) return TTK_ERR_BAD_SHAPE;
  int64_t bx = (max_numel + 256 * 4 - 1) / (256 * 4);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  multi_cast_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n)), 256, 0, stream>>>(table);
  return launch_status();
}

int ttk_head_bwd(const void* dz, int token_size, const void* xn, int64_t ld, const int32_t* latent_row, const void* w_out,
                 void* dxn, float* dw, float* db, int T, int width, cudaStream_t stream) {
  if (!dz || !xn || !latent_row || !w_out || !dxn || !dw || !db) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!bw_width_ok(width) || ld % 8 || token_size < 1 || token_size > 8) return TTK_ERR_BAD_SHAPE;
  if (T <= 0) return TTK_OK;
  int gx = (T + BW_WARPS - 1) / BW_WARPS;
  if (gx > 2 * num_sms()) gx = 2 * num_sms();
  TTK_BW_NV(width, head_bwd_kernel<NV><<<dim3(gx, token_size), BW_WARPS * 32, 0, stream>>>(
                       static_cast<const __nv_bfloat16*>(dz), token_size, static_cast<const __nv_bfloat16*>(xn), ld,
                       latent_row, static_cast<const __nv_bfloat16*>(w_out), static_cast<__nv_bfloat16*>(dxn), dw, db, T));
  return launch_status();
}

int ttk_dec_in_bwd(const void* de, int64_t ld, const int32_t* latent_row, const void* codes, int token_size,
                   const void* w_in, float* dcodes, float* dw, float* db, int T, int width, cudaStream_t stream) {
  if (!de || !latent_row || !codes || !w_in || !dcodes || !dw || !db) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (!bw_width_ok(width) || ld % 8 || token_size < 1 || token_size > 8) return TTK_ERR_BAD_SHAPE;
  if (T <= 0) return TTK_OK;
  int gx = (T + BW_WARPS - 1) / BW_WARPS;
  if (gx > 2 * num_sms()) gx = 2 * num_sms();
  TTK_BW_NV(width, dec_in_bwd_kernel<NV><<<dim3(gx, token_size), BW_WARPS * 32, 0, stream>>>(
                       static_cast<const __nv_bfloat16*>(de), ld, latent_row, static_cast<const __nv_bfloat16*>(codes),
                       token_size, static_cast<const __nv_bfloat16*>(w_in), dcodes, dw, db, T));
  return launch_status();
}

}  // extern "C"
