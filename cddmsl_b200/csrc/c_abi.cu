// Library-wide pieces of the C ABI: version, error strings, launch counter, tuning knobs.
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"

namespace cddmsl {
unsigned long long g_launch_count = 0;
int tune_roi(const char* key, int value);
extern int g_head_tc;

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}
}  // namespace cddmsl

extern "C" int cddmsl_abi_version(void) { return 1; }

extern "C" uint64_t cddmsl_launch_count(void) { return (uint64_t)cddmsl::g_launch_count; }

extern "C" const char* cddmsl_error_string(int code) {
  switch (code) {
    case CDDMSL_OK: return "ok";
    case CDDMSL_EINVAL: return "cddmsl: invalid argument (shape, null pointer or unsupported size)";
    case CDDMSL_EWORKSPACE: return "cddmsl: workspace too small";
    case CDDMSL_EALIGN: return "cddmsl: pointer alignment requirement not met";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "cddmsl: unknown error";
  }
}

// Internal tuning hook used by bench sweeps (not part of the reference-facing surface).
extern "C" int cddmsl_tune(const char* key, int value) {
  if (!strcmp(key, "head_tc")) {
    cddmsl::g_head_tc = value;
    return 1;
  }
  return cddmsl::tune_roi(key, value);
}
