// Library-wide pieces of the C ABI: version, thread-local error message, device attribute cache.
#include <mutex>
#include <string.h>

#include "common.cuh"

namespace nsc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;   // immutable once written; a benign race writes the same value
  }
  return cached[dev];
}

}  // namespace nsc

extern "C" {

int nsc_version(void) { return 100; /* 0.1.0 */ }

const char* nsc_last_error(void) { return nsc::g_err; }

}  // extern "C"
