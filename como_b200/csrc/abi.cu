// Error plumbing + version for the como_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace como {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return COMO_B200_ELAUNCH;
  }
  return COMO_B200_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

}  // namespace como

extern "C" int como_b200_abi_version(void) { return 3; }
extern "C" const char* como_b200_last_error(void) { return como::g_err; }

// Host evaluation of the SAME se3_exp_tau_phi the kernels call (common.cuh, __host__ __device__): lets a CPU-only test
// pin the restatement of lietorch's SE(3) exponential to the matrix exponential without a GPU.
extern "C" void como_b200_se3_exp(const double* tau_phi, double* T16) {
  como::se3_exp_tau_phi(tau_phi, tau_phi + 3, T16);
}
