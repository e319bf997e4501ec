#include "host_util.h"
#include <string.h>
#include <atomic>
#ifndef VC_CUDA_BUILD
#define VC_CUDA_BUILD 0
#endif

namespace vck {

static thread_local char g_err[1024] = {0};

int set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "unknown error", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return 1;
}

const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
void launch_count_reset() { g_launches.store(0); }
static std::atomic<long long> g_pair_launches{0};
void count_pair_launch() { g_pair_launches.fetch_add(1, std::memory_order_relaxed); }
long long pair_launch_count() { return g_pair_launches.load(); }

#if !VC_CUDA_BUILD
void gemm_profile_enable(int) {}
int gemm_profile_dump(const char*) { return 0; }
int gemm_profile_read_min(double, double* a, double* b, long long* c) { return gemm_profile_read(a, b, c); }
int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  if (total_ms) *total_ms = 0;
  if (total_flops) *total_flops = 0;
  if (launches) *launches = 0;
  return 0;
}
#endif

}  // namespace vck
