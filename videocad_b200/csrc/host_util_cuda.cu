#include <cuda_runtime.h>
#include <stdio.h>
#include <mutex>
#include <vector>
#include "host_util.h"
#include "gemm_profile.h"

namespace vck {

int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  char msg[512];
  snprintf(msg, sizeof msg, "%s: %s", what, cudaGetErrorString(e));
  return set_error(msg);
}

namespace {
struct Rec { cudaEvent_t a, b; double flops; char tag[64]; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;     // recorded pairs
std::vector<Rec> g_free;     // reusable event pairs
}  // namespace

void gemm_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = enable != 0;
  for (auto& r : g_recs) g_free.push_back(r);
  g_recs.clear();
}

bool gemm_profile_begin(cudaStream_t st, double flops, int* slot, const char* tag) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_on) return false;
  Rec r;
  if (!g_free.empty()) { r = g_free.back(); g_free.pop_back(); }
  else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
  r.flops = flops;
  snprintf(r.tag, sizeof r.tag, "%s", tag ? tag : "");
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  *slot = (int)g_recs.size() - 1;
  return true;
}
void gemm_profile_end(cudaStream_t st, int slot) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot >= 0 && slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, st);
}

int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  double ms = 0, fl = 0;
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float t = 0;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms += t;
    fl += r.flops;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (long long)g_recs.size();
  return 0;
}

int gemm_profile_dump(const char* path) {
  std::lock_guard<std::mutex> lk(g_mu);
  FILE* f = fopen(path, "w");
  if (!f) return set_error("gemm_profile_dump: cannot open file");
  fprintf(f, "tag,flops,ms\n");
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float t = 0;
    cudaEventElapsedTime(&t, r.a, r.b);
    fprintf(f, "%s,%.0f,%.6f\n", r.tag, r.flops, t);
  }
  fclose(f);
  return 0;
}

}  // namespace vck
