// CUDA-event timing of GEMM launches (bench instrumentation; see vc_gemm_profile in include/videocad_b200.h)
#pragma once
#include <cuda_runtime.h>
namespace vck {
bool gemm_profile_begin(cudaStream_t st, double flops, int* slot, const char* tag = nullptr);
void gemm_profile_end(cudaStream_t st, int slot);
}  // namespace vck
