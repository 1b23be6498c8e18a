// ABI housekeeping: version, thread-local error string, device probes.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace rl {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rl

extern "C" int rl_abi_version(void) { return 1; }
extern "C" const char* rl_last_error(void) { return rl::g_err; }
extern "C" int rl_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    rl::set_error("rl_device_sm_count: %s", cudaGetErrorString(cudaGetLastError()));
    return RL_ERR_CUDA;
  }
  return n;
}
