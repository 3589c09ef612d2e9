// Host-side helpers shared by the C-ABI launchers: TMA tensor-map encoding through the
// driver entry point (no link-time dependency on libcuda), device checks, launch error mapping.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "common.cuh"

namespace ttk {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

PFN_tensorMapEncodeTiled get_encode_fn();

// 2-D bf16 row-major tensor [rows, cols] with row pitch `ld` elements; box = [box_rows, 64 cols],
// SWIZZLE_128B, out-of-bounds elements read as zero.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols = 64);

// development aid: device buffer for clock64 pipeline stamps (ttk_debug_set_trace), null in production
extern long long* g_trace;

// Both are answered for the CURRENT device of the calling thread (cached per device ordinal).
int check_device_sm100();
int num_sms();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: set it once per (kernel, device).
// Usage in a launcher:  static PerDeviceOnce once;  if (int e = set_smem_attr_once(once, (const void*)kern, bytes)) return e;
constexpr int TTK_MAX_DEVICES = 64;
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
};
int set_smem_attr_once(PerDeviceOnce& once, const void* kernel, int bytes);

// Launch with the programmatic-stream-serialization attribute (TTK_PDL=0 switches it off): only for kernels that call
// pdl_wait() before their first global-memory access.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? TTK_OK : TTK_ERR_CUDA; }
inline int launch_status() { return cuda_status(cudaGetLastError()); }

}  // namespace ttk
