// Error plumbing and version of the C ABI (include/gpar_b200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace gpar {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -1000 - (int)e;
  }
  return 0;
}
}  // namespace gpar

extern "C" int gpar_abi_version(void) { return GPAR_ABI_VERSION; }
extern "C" const char* gpar_last_error(void) { return gpar::g_err; }
