#include "host_util.cuh"

#include <mutex>

namespace ttk {

static PFN_tensorMapEncodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

PFN_tensorMapEncodeTiled get_encode_fn() {
  std::call_once(g_encode_once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      g_encode = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
    }
    (void)cudaGetLastError();
  });
  return g_encode;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  PFN_tensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return TTK_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0) return TTK_ERR_ALIGNMENT;
  if (box_rows == 0 || box_rows > 256 || box_cols * 2 > 128) return TTK_ERR_BAD_SHAPE;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TTK_OK : TTK_ERR_DRIVER;
}

long long* g_trace = nullptr;

static int g_sm_major = -1, g_sm_minor = -1, g_num_sms = 0;
static std::once_flag g_dev_once;
static void query_dev() {
  std::call_once(g_dev_once, []() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&g_sm_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&g_sm_minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
}

int check_device_sm100() {
  query_dev();
  return (g_sm_major == 10) ? TTK_OK : TTK_ERR_ARCH;
}
int num_sms() {
  query_dev();
  return g_num_sms > 0 ? g_num_sms : 148;
}

}  // namespace ttk
