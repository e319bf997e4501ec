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

int gemm_profile_read_min(double min_flops, double* total_ms, double* total_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  double ms = 0, fl = 0;
  long long n = 0;
  for (auto& r : g_recs) {
    if (r.flops < min_flops) continue;
    cudaEventSynchronize(r.b);
    float t = 0;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms += t;
    fl += r.flops;
    ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = n;
  return 0;
}

int gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  return gemm_profile_read_min(0.0, total_ms, total_flops, launches);
}

namespace {
constexpr int kSide = 4, kEvents = 128;  // 0,1: sequence transformer; 2,3: image encoders
cudaStream_t g_side[kSide] = {nullptr, nullptr, nullptr, nullptr};
cudaEvent_t g_ev[kEvents];
bool g_side_init = false;
int g_ev_next = 0;
int side_init() {
  if (g_side_init) return 0;
  for (int i = 0; i < kSide; ++i)
    if (cudaStreamCreateWithFlags(&g_side[i], cudaStreamNonBlocking) != cudaSuccess) return set_error("cudaStreamCreate failed");
  for (int i = 0; i < kEvents; ++i)
    if (cudaEventCreateWithFlags(&g_ev[i], cudaEventDisableTiming) != cudaSuccess) return set_error("cudaEventCreate failed");
  g_side_init = true;
  return 0;
}
cudaEvent_t next_event() {
  cudaEvent_t e = g_ev[g_ev_next];
  g_ev_next = (g_ev_next + 1) % kEvents;
  return e;
}
}  // namespace

static bool g_side_enabled = true;
void side_streams_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_side_enabled = enable != 0;
}

int stream_fork(void* main_s, int i, void** side) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (i < 0 || i >= kSide) return set_error("stream_fork: bad index");
  if (!g_side_enabled) { *side = main_s; return 0; }  // serialised mode (per-kernel timing): everything on the caller's stream
  if (int rc = side_init()) return rc;
  cudaEvent_t e = next_event();
  if (cudaEventRecord(e, reinterpret_cast<cudaStream_t>(main_s)) != cudaSuccess) return set_error("stream_fork: event record failed");
  if (cudaStreamWaitEvent(g_side[i], e, 0) != cudaSuccess) return set_error("stream_fork: wait failed");
  *side = g_side[i];
  return 0;
}

int stream_join(void* main_s, int i) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (i < 0 || i >= kSide) return set_error("stream_join: bad index");
  if (!g_side_enabled) return 0;
  if (int rc = side_init()) return rc;
  cudaEvent_t e = next_event();
  if (cudaEventRecord(e, g_side[i]) != cudaSuccess) return set_error("stream_join: event record failed");
  if (cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(main_s), e, 0) != cudaSuccess) return set_error("stream_join: wait failed");
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
