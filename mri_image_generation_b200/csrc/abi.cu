// Error reporting and device checks for the C ABI (include/mri_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mri_b200.h"
#include "common.h"

namespace mri {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return -100 - (int)e;
}
int check_launch(const char* kernel) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return 0;
  (void)cudaGetLastError();
  return set_cuda_error(e, kernel);
}

}  // namespace mri

extern "C" int mri_abi_version(void) { return MRI_B200_ABI_VERSION; }
extern "C" const char* mri_last_error(void) { return mri::g_err; }
extern "C" int mri_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return mri::set_cuda_error(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return mri::set_cuda_error(e, "cudaDeviceGetAttribute");
  return major == 10 ? 1 : 0;
}
