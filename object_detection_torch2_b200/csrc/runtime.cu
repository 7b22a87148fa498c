// Error plumbing and build/device facts of libssdhead.
#include <stdarg.h>

#include "common.cuh"

namespace ssdh {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_status(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

}  // namespace ssdh

extern "C" int ssdh_version(void) { return SSDH_VERSION; }

extern "C" const char* ssdh_last_error(void) { return ssdh::g_error; }
