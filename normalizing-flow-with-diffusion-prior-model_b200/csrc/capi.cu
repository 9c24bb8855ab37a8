// Error plumbing and misc entry points of the C ABI.
#include <stdlib.h>

#include "common.cuh"

namespace nfdpm {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NFDPM_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
}  // namespace nfdpm

extern "C" int nfdpm_version(void) { return NFDPM_VERSION; }
extern "C" const char* nfdpm_last_error_string(void) { return nfdpm::err_buf(); }
extern "C" int nfdpm_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
