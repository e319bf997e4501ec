#include <cuda_runtime.h>
#include <stdio.h>
#include "host_util.h"

namespace vck {

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  char msg[512];
  snprintf(msg, sizeof msg, "%s: %s", what, cudaGetErrorString(e));
  return set_error(msg);
}

}  // namespace vck
