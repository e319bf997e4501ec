// Kernel launch helper: every kernel of the library is launched with the programmatic-stream-serialization attribute
// (programmatic dependent launch) and begins with pdl_grid_sync().
//
// A training step is ~600 dependent kernels, most of them a few microseconds long; with ordinary stream order each one
// is scheduled only after its predecessor has drained and its memory flush has completed.  With PDL the next kernel's CTAs
// are dispatched as soon as every CTA of the running kernel has issued `griddepcontrol.launch_dependents` (first statement of
// each kernel) and SM resources allow; they then block in `griddepcontrol.wait` until the predecessor grid has completed and
// its writes are visible.  No kernel touches global memory before that wait, so the data flow is exactly that of stream
// order; only launch latency and per-CTA setup overlap with the predecessor's tail.  The edges survive CUDA-graph capture.
// VC_PDL=0 launches without the attribute (the device-side instructions are then no-ops).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <utility>

namespace vck {

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() {
  pdl_trigger();
  pdl_wait();
}
#endif

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VC_PDL");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}

template <typename K, typename... Args>
inline void launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);  // failures surface through check_launch (cudaGetLastError)
}

}  // namespace vck

#define VC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::vck::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
