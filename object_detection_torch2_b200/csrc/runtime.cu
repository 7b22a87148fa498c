// Error plumbing and build/device facts of libssdhead.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <utility>

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

int ensure_dyn_smem(const void* kernel, int bytes, const char* what) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("%s: cudaGetDevice: %s", what, cudaGetErrorString(e)); return static_cast<int>(e); }
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(kernel, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return 0;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(%d bytes): %s", what, bytes, cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  done[key] = bytes;
  return 0;
}

}  // namespace ssdh

extern "C" int ssdh_version(void) { return SSDH_VERSION; }

extern "C" const char* ssdh_last_error(void) { return ssdh::g_error; }
