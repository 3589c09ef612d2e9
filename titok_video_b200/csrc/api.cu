// Library-level entry points of the C ABI (status strings, version).
#include "common.cuh"

extern "C" {

const char* ttk_strerror(int status) {
  switch (status) {
    case ttk::TTK_OK: return "ok";
    case ttk::TTK_ERR_BAD_ARG: return "bad argument (null pointer or invalid enum)";
    case ttk::TTK_ERR_BAD_SHAPE: return "unsupported shape";
    case ttk::TTK_ERR_ALIGNMENT: return "pointer or leading dimension is not 16-byte aligned";
    case ttk::TTK_ERR_ARCH: return "device is not compute capability 10.x (sm_100a kernels only, no fallback)";
    case ttk::TTK_ERR_CUDA: return "CUDA runtime error at launch";
    case ttk::TTK_ERR_DRIVER: return "cuTensorMapEncodeTiled failed or is unavailable";
    case ttk::TTK_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

int ttk_version(void) { return 100; }

}  // extern "C"
