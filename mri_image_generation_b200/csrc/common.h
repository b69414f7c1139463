// Error plumbing shared by all translation units of libmri_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace mri {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
// cudaPeekAtLastError after a launch (does not synchronize)
int check_launch(const char* kernel);
}  // namespace mri
