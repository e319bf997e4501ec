// Host-side error plumbing shared by the CUDA translation units (and by the CPU emulation build).
#pragma once
#include <string>

namespace vck {

// stores msg as the calling thread's last error and returns a non-zero status
int set_error(const char* msg);
const char* last_error();
void count_launch();
long long launch_count();
void launch_count_reset();
// launches that took the CTA-pair (cta_group::2) GEMM kernel
void count_pair_launch();
long long pair_launch_count();
// GEMM profiling hooks (CUDA build: event pairs around each launch; emulation build: no-ops)
void gemm_profile_enable(int enable);
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches);
int gemm_profile_read_min(double min_flops, double* total_ms, double* total_flops, long long* launches);
int gemm_profile_dump(const char* path);  // per-launch CSV (tag,flops,ms)

#if defined(__CUDACC__)
// checks cudaGetLastError() after a launch; returns 0 or sets the error
int check_launch(const char* what);
#endif

}  // namespace vck
