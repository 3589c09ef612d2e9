#include "host_util.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

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

// A tensor map is a pure function of (address, extents, pitch, box): the workspaces, weights and activation slabs of a
// step keep their addresses, so the ~4 descriptors of each of a training step's ~215 launches are looked up instead of
// re-encoded through the driver (~0.7 us each: ~0.6 ms of host time per step at the reference's 3-clip batch, where
// the step is bound by the host's launch rate). Per-thread cache, bounded, no locking.
namespace {
struct TmapKey {
  uintptr_t base;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = k.base * 0x9e3779b97f4a7c15ull;
    h ^= (k.rows + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 0xc2b2ae3d27d4eb4full + (h << 6) + (h >> 2));
    h ^= (k.ld * 0x165667b19e3779f9ull + (h << 6) + (h >> 2));
    h ^= ((static_cast<uint64_t>(k.box_rows) << 32 | k.box_cols) + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};
constexpr size_t TMAP_CACHE_MAX = 8192;
}  // namespace

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  PFN_tensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return TTK_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0) return TTK_ERR_ALIGNMENT;
  if (box_rows == 0 || box_rows > 256 || box_cols * 2 > 128) return TTK_ERR_BAD_SHAPE;
  thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{reinterpret_cast<uintptr_t>(base), rows, cols, ld, box_rows, box_cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    std::memcpy(out, &it->second, sizeof(CUtensorMap));
    return TTK_OK;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return TTK_ERR_DRIVER;
  if (cache.size() >= TMAP_CACHE_MAX) cache.clear();
  cache.emplace(key, *out);
  return TTK_OK;
}

long long* g_trace = nullptr;

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = std::getenv("TTK_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Per-device state (the library is re-entrant across devices and threads: nothing is cached for "the first device seen").
struct DevInfo {
  std::atomic<int> ready{0};
  int major = -1, minor = -1, sms = 0;
};
static DevInfo g_dev[TTK_MAX_DEVICES];

static const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TTK_MAX_DEVICES) return nullptr;
  DevInfo& d = g_dev[dev];
  if (!d.ready.load(std::memory_order_acquire)) {
    int mj = -1, mn = -1, sms = 0;  // racing threads write identical values
    cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    d.major = mj;
    d.minor = mn;
    d.sms = sms;
    d.ready.store(1, std::memory_order_release);
  }
  return &d;
}

int check_device_sm100() {
  const DevInfo* d = dev_info();
  return (d && d->major == 10) ? TTK_OK : TTK_ERR_ARCH;
}
int num_sms() {
  const DevInfo* d = dev_info();
  return (d && d->sms > 0) ? d->sms : 148;
}

int set_smem_attr_once(PerDeviceOnce& once, const void* kernel, int bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return TTK_ERR_CUDA;
  if (dev < 0 || dev >= TTK_MAX_DEVICES) {  // beyond the cache: set it every time (cheap, idempotent)
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? TTK_OK : TTK_ERR_CUDA;
  }
  const unsigned long long bit = 1ull << dev;
  if (once.mask.load(std::memory_order_acquire) & bit) return TTK_OK;
  // function attributes are per device; two threads racing here set the same value twice, which is harmless
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return TTK_ERR_CUDA;
  once.mask.fetch_or(bit, std::memory_order_release);
  return TTK_OK;
}

}  // namespace ttk
