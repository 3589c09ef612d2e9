// FSQ quantizer kernels (model/quantizer/fsq.py) and the codebook-usage histogram
// (train_utils/codebook_logging.py:20-30). All bandwidth-bound: 16-byte coalesced global access,
// tiles staged through shared memory so that D-element vectors (D = 5 is not a power of two) never
// cause misaligned or strided global transactions.
#include "common.cuh"
#include "fsq.cuh"
#include "host_util.cuh"

namespace ttk {

constexpr int FSQ_TILE = 1024;  // vectors per CTA
constexpr int FSQ_THREADS = 256;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __ushort_as_bfloat16(static_cast<unsigned short>(pack_bf16x2(v, 0.f) & 0xffffu));  // F2FP, not the XU-pipe F2F
}

// coalesced copy of `n` elements global <-> shared using 16-byte accesses where alignment allows
template <typename T>
__device__ __forceinline__ void tile_load(T* s, const T* g, int n) {
  constexpr int V = 16 / sizeof(T);
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int nv = n / V;
    for (int i = threadIdx.x; i < nv; i += blockDim.x)
      reinterpret_cast<uint4*>(s)[i] = ldg16_stream(reinterpret_cast<const uint4*>(g) + i);
    for (int i = nv * V + threadIdx.x; i < n; i += blockDim.x) s[i] = g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = g[i];
  }
}
template <typename T>
__device__ __forceinline__ void tile_store(T* g, const T* s, int n) {
  constexpr int V = 16 / sizeof(T);
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int nv = n / V;
    for (int i = threadIdx.x; i < nv; i += blockDim.x)
      reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(s)[i];
    for (int i = nv * V + threadIdx.x; i < n; i += blockDim.x) g[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = s[i];
  }
}

// FSQ.forward (fsq.py:123-135): codes in z's dtype, int32 indices.
template <typename T>
__global__ void __launch_bounds__(FSQ_THREADS) fsq_fwd_kernel(const T* __restrict__ z, T* __restrict__ codes,
                                                              int32_t* __restrict__ indices, int64_t n,
                                                              const __grid_constant__ FsqConsts c) {
  extern __shared__ uint4 fsq_smem[];
  T* s = reinterpret_cast<T*>(fsq_smem);
  const int D = c.D;
  for (int64_t tile = blockIdx.x; tile * FSQ_TILE < n; tile += gridDim.x) {
    const int64_t v0 = tile * FSQ_TILE;
    const int nv = static_cast<int>(min(static_cast<int64_t>(FSQ_TILE), n - v0));
    tile_load(s, z + v0 * D, nv * D);
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      float idx = 0.f;
      T* sv = s + v * D;
#pragma unroll
      for (int d = 0; d < FSQ_MAX_D; ++d) {
        if (d < D) sv[d] = from_f32<T>(fsq_quantize_dim(to_f32<T>(sv[d]), c, d, idx));
      }
      indices[v0 + v] = static_cast<int32_t>(idx);
    }
    __syncthreads();
    tile_store(codes + v0 * D, s, nv * D);
    __syncthreads();
  }
}

// Straight-through backward of FSQ.forward: d z = d codes * half_l * (1 - tanh^2(z + shift)) / half_width
template <typename T>
__global__ void __launch_bounds__(FSQ_THREADS) fsq_bwd_kernel(const T* __restrict__ z, const T* __restrict__ dcodes,
                                                              T* __restrict__ dz, int64_t n_elems,
                                                              const __grid_constant__ FsqConsts c) {
  const int D = c.D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_elems;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const float t = tanhf(__fadd_rn(to_f32<T>(z[i]), c.shift[d]));
    const float g = __fdiv_rn(to_f32<T>(dcodes[i]), c.half_width[d]);
    dz[i] = from_f32<T>(__fmul_rn(__fmul_rn(g, c.half_l[d]), __fsub_rn(1.0f, __fmul_rn(t, t))));
  }
}

// FSQ.indices_to_codes (fsq.py:96-103,111-121): ((idx // basis) % levels - half_width) / half_width
template <typename TI, typename TO>
__global__ void __launch_bounds__(FSQ_THREADS) fsq_i2c_kernel(const TI* __restrict__ idx, TO* __restrict__ codes,
                                                              int64_t n, const __grid_constant__ FsqConsts c) {
  extern __shared__ uint4 fsq_smem[];
  TO* s = reinterpret_cast<TO*>(fsq_smem);
  const int D = c.D;
  for (int64_t tile = blockIdx.x; tile * FSQ_TILE < n; tile += gridDim.x) {
    const int64_t v0 = tile * FSQ_TILE;
    const int nv = static_cast<int>(min(static_cast<int64_t>(FSQ_TILE), n - v0));
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      const long long id = static_cast<long long>(idx[v0 + v]);
#pragma unroll
      for (int d = 0; d < FSQ_MAX_D; ++d) {
        if (d < D) {
          // python floor-div / mod semantics for non-negative divisors
          long long q = id / c.ibasis[d];
          if ((id % c.ibasis[d] != 0) && (id < 0)) --q;
          long long m = q % c.levels[d];
          if (m < 0) m += c.levels[d];
          const float lv = static_cast<float>(m);
          s[v * D + d] = from_f32<TO>(fsq_div_hw(__fsub_rn(lv, c.half_width[d]), c.half_width[d], c.rcp_half_width[d]));
        }
      }
    }
    __syncthreads();
    tile_store(codes + v0 * D, s, nv * D);
    __syncthreads();
  }
}

// Codebook-usage histogram: counts[k] += #(indices == k). Shared-memory atomics per CTA, one flush of
// the non-zero bins with global atomics.
__global__ void __launch_bounds__(512) hist_smem_kernel(const int32_t* __restrict__ idx, int64_t n, int K,
                                                        unsigned int* __restrict__ counts) {
  extern __shared__ unsigned int hbins[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) hbins[k] = 0;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(idx) & 15u) == 0) ? n / 4 : 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const uint4 v = ldg16_stream(reinterpret_cast<const uint4*>(idx) + i);
    const unsigned int a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (a[e] < static_cast<unsigned int>(K)) atomicAdd(&hbins[a[e]], 1u);
  }
  for (int64_t i = n4 * 4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const unsigned int a = static_cast<unsigned int>(idx[i]);
    if (a < static_cast<unsigned int>(K)) atomicAdd(&hbins[a], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const unsigned int v = hbins[k];
    if (v) atomicAdd(&counts[k], v);
  }
}

__global__ void __launch_bounds__(512) hist_gmem_kernel(const int32_t* __restrict__ idx, int64_t n, int K,
                                                        unsigned int* __restrict__ counts) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const unsigned int a = static_cast<unsigned int>(idx[i]);
    if (a < static_cast<unsigned int>(K)) atomicAdd(&counts[a], 1u);
  }
}

// usage / entropy of a count vector (codebook_logging.py:26-29): out[0] = #non-zero bins,
// out[1] = entropy in nats of counts / sum(counts), out[2] = sum(counts).
__global__ void __launch_bounds__(1024) codebook_stats_kernel(const unsigned int* __restrict__ counts, int K,
                                                              double* __restrict__ out) {
  __shared__ double s_a[32], s_b[32];
  auto block_sum = [&](double v) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x < 32) {
      r = (threadIdx.x < (blockDim.x >> 5)) ? s_a[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (threadIdx.x == 0) s_b[0] = r;
    }
    __syncthreads();
    r = s_b[0];
    __syncthreads();
    return r;
  };
  double tot = 0, nz = 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    tot += counts[k];
    nz += counts[k] != 0;
  }
  tot = block_sum(tot);
  nz = block_sum(nz);
  double h = 0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double c = counts[k];
    if (c > 0) {
      const double pk = c / tot;
      h -= pk * log(pk);
    }
  }
  h = block_sum(h);
  if (threadIdx.x == 0) {
    out[0] = nz;
    out[1] = tot > 0 ? h : 0.0;
    out[2] = tot;
  }
}

static int fill_consts(FsqConsts& c, int D, const float* half_l, const float* offset, const float* shift,
                       const float* half_width, const int32_t* basis, const int32_t* levels) {
  if (D < 1 || D > FSQ_MAX_D) return TTK_ERR_BAD_SHAPE;
  if (!half_l || !offset || !shift || !half_width || !basis || !levels) return TTK_ERR_BAD_ARG;
  c.D = D;
  for (int d = 0; d < FSQ_MAX_D; ++d) {
    const bool ok = d < D;
    c.half_l[d] = ok ? half_l[d] : 0.f;
    c.offset[d] = ok ? offset[d] : 0.f;
    c.shift[d] = ok ? shift[d] : 0.f;
    c.half_width[d] = ok ? half_width[d] : 1.f;
    c.rcp_half_width[d] = 1.0f / c.half_width[d];
    c.basis[d] = ok ? static_cast<float>(basis[d]) : 0.f;
    c.ibasis[d] = ok ? basis[d] : 1;
    c.levels[d] = ok ? levels[d] : 1;
  }
  return TTK_OK;
}

int fsq_make_consts(FsqConsts& c, int D, const float* half_l, const float* offset, const float* shift,
                    const float* half_width, const int32_t* basis, const int32_t* levels) {
  return fill_consts(c, D, half_l, offset, shift, half_width, basis, levels);
}

}  // namespace ttk

using namespace ttk;

extern "C" {

// dtype: 0 = bf16, 1 = fp32. Host arrays half_l/offset/shift/half_width/basis/levels have D entries.
int ttk_fsq_fwd(const void* z, void* codes, int32_t* indices, int64_t n, int dtype, int D, const float* half_l,
                const float* offset, const float* shift, const float* half_width, const int32_t* basis,
                const int32_t* levels, cudaStream_t stream) {
  if (n <= 0) return TTK_OK;  // empty batch: nothing to do (empty tensors have null data pointers)
  if (!z || !codes || !indices) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  FsqConsts c;
  if (int e = fill_consts(c, D, half_l, offset, shift, half_width, basis, levels)) return e;
  if (n <= 0) return TTK_OK;
  const int64_t tiles = (n + FSQ_TILE - 1) / FSQ_TILE;
  const int grid = static_cast<int>(tiles < 8LL * num_sms() ? tiles : 8LL * num_sms());
  if (dtype == 0) {
    fsq_fwd_kernel<__nv_bfloat16><<<grid, FSQ_THREADS, FSQ_TILE * D * 2, stream>>>(
        static_cast<const __nv_bfloat16*>(z), static_cast<__nv_bfloat16*>(codes), indices, n, c);
  } else if (dtype == 1) {
    fsq_fwd_kernel<float><<<grid, FSQ_THREADS, FSQ_TILE * D * 4, stream>>>(static_cast<const float*>(z),
                                                                         static_cast<float*>(codes), indices, n, c);
  } else {
    return TTK_ERR_BAD_ARG;
  }
  return launch_status();
}

int ttk_fsq_bwd(const void* z, const void* dcodes, void* dz, int64_t n, int dtype, int D, const float* half_l,
                const float* offset, const float* shift, const float* half_width, const int32_t* basis,
                const int32_t* levels, cudaStream_t stream) {
  if (n <= 0) return TTK_OK;
  if (!z || !dcodes || !dz) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  FsqConsts c;
  if (int e = fill_consts(c, D, half_l, offset, shift, half_width, basis, levels)) return e;
  if (n <= 0) return TTK_OK;
  const int64_t ne = n * D;
  const int64_t blocks = (ne + FSQ_THREADS - 1) / FSQ_THREADS;
  const int grid = static_cast<int>(blocks < 16LL * num_sms() ? blocks : 16LL * num_sms());
  if (dtype == 0)
    fsq_bwd_kernel<__nv_bfloat16><<<grid, FSQ_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(z),
                                                                   static_cast<const __nv_bfloat16*>(dcodes),
                                                                   static_cast<__nv_bfloat16*>(dz), ne, c);
  else if (dtype == 1)
    fsq_bwd_kernel<float><<<grid, FSQ_THREADS, 0, stream>>>(static_cast<const float*>(z),
                                                           static_cast<const float*>(dcodes),
                                                           static_cast<float*>(dz), ne, c);
  else
    return TTK_ERR_BAD_ARG;
  return launch_status();
}

// idx_dtype: 0 = int32, 1 = int64. out_dtype: 0 = bf16, 1 = fp32.
int ttk_fsq_indices_to_codes(const void* idx, int idx_dtype, void* codes, int out_dtype, int64_t n, int D,
                             const float* half_width, const int32_t* basis, const int32_t* levels,
                             cudaStream_t stream) {
  if (n <= 0) return TTK_OK;
  if (!idx || !codes) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  FsqConsts c;
  float zeros[FSQ_MAX_D] = {0};
  if (int e = fill_consts(c, D, zeros, zeros, zeros, half_width, basis, levels)) return e;
  if (n <= 0) return TTK_OK;
  const int64_t tiles = (n + FSQ_TILE - 1) / FSQ_TILE;
  const int grid = static_cast<int>(tiles < 8LL * num_sms() ? tiles : 8LL * num_sms());
  const size_t smem = static_cast<size_t>(FSQ_TILE) * D * (out_dtype == 0 ? 2 : 4);
  if (idx_dtype == 0 && out_dtype == 0)
    fsq_i2c_kernel<int32_t, __nv_bfloat16><<<grid, FSQ_THREADS, smem, stream>>>(
        static_cast<const int32_t*>(idx), static_cast<__nv_bfloat16*>(codes), n, c);
  else if (idx_dtype == 0 && out_dtype == 1)
    fsq_i2c_kernel<int32_t, float><<<grid, FSQ_THREADS, smem, stream>>>(static_cast<const int32_t*>(idx),
                                                                       static_cast<float*>(codes), n, c);
  else if (idx_dtype == 1 && out_dtype == 0)
    fsq_i2c_kernel<int64_t, __nv_bfloat16><<<grid, FSQ_THREADS, smem, stream>>>(
        static_cast<const int64_t*>(idx), static_cast<__nv_bfloat16*>(codes), n, c);
  else if (idx_dtype == 1 && out_dtype == 1)
    fsq_i2c_kernel<int64_t, float><<<grid, FSQ_THREADS, smem, stream>>>(static_cast<const int64_t*>(idx),
                                                                       static_cast<float*>(codes), n, c);
  else
    return TTK_ERR_BAD_ARG;
  return launch_status();
}

// counts[K] (uint32, caller-zeroed or accumulated across calls) += histogram of idx[n]; values outside
// [0, K) are ignored.
int ttk_hist_u32(const int32_t* idx, int64_t n, int K, uint32_t* counts, cudaStream_t stream) {
  if (n <= 0) return TTK_OK;
  if (!idx || !counts || K <= 0) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  if (n <= 0) return TTK_OK;
  const int threads = 512;
  const int64_t per_cta = 64 * 1024;  // indices per CTA so the flush of K bins is amortised
  int64_t blocks = (n + per_cta - 1) / per_cta;
  if (blocks > 2LL * num_sms()) blocks = 2LL * num_sms();
  if (blocks < 1) blocks = 1;
  const size_t smem = static_cast<size_t>(K) * 4;
  if (smem <= 96 * 1024) {
    static PerDeviceOnce once;
    if (int e = set_smem_attr_once(once, reinterpret_cast<const void*>(hist_smem_kernel), 96 * 1024)) return e;
    hist_smem_kernel<<<static_cast<int>(blocks), threads, smem, stream>>>(idx, n, K, counts);
  } else {
    hist_gmem_kernel<<<static_cast<int>(blocks * 4), threads, 0, stream>>>(idx, n, K, counts);
  }
  return launch_status();
}

// out (device, 3 doubles): #non-zero bins, entropy (nats), total count.
int ttk_codebook_stats(const uint32_t* counts, int K, double* out, cudaStream_t stream) {
  if (!counts || !out || K <= 0) return TTK_ERR_BAD_ARG;
  if (int e = check_device_sm100()) return e;
  codebook_stats_kernel<<<1, 1024, 0, stream>>>(counts, K, out);
  return launch_status();
}

}  // extern "C"
