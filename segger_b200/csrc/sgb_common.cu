// Error reporting, version, device queries.
#include "sgb_api_internal.cuh"
#include <cstring>

namespace sgb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SGB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SGB_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace sgb

extern "C" int sgb_version(void) { return SGB_VERSION; }
extern "C" const char* sgb_last_error(void) { return sgb::g_err; }
