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
// Auxiliary streams and fork/join events, one pool PER DEVICE (created on first use while that device is current; the callers
// launch in the device context of their tensors).  Indices 0,1: sequence transformer; 2,3: frame encoder; 4,5: CAD encoder
// (the two encoders run concurrently on different torch streams and must not share side streams: false dependencies).
constexpr int kSide = 6, kEvents = 256, kMaxDev = 64;
struct SidePool {
  cudaStream_t side[kSide];
  cudaEvent_t ev[kEvents];
  int ev_next;
  bool init;
};
SidePool g_pools[kMaxDev] = {};
SidePool* side_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { set_error("stream_fork: cannot query the current device"); return nullptr; }
  SidePool& P = g_pools[dev];
  if (!P.init) {
    for (int i = 0; i < kSide; ++i)
      if (cudaStreamCreateWithFlags(&P.side[i], cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return nullptr; }
    for (int i = 0; i < kEvents; ++i)
      if (cudaEventCreateWithFlags(&P.ev[i], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); return nullptr; }
    P.ev_next = 0;
    P.init = true;
  }
  return &P;
}
cudaEvent_t next_event(SidePool& P) {
  cudaEvent_t e = P.ev[P.ev_next];
  P.ev_next = (P.ev_next + 1) % kEvents;
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
  SidePool* P = side_pool();
  if (!P) return 1;
  cudaEvent_t e = next_event(*P);
  if (cudaEventRecord(e, reinterpret_cast<cudaStream_t>(main_s)) != cudaSuccess) return set_error("stream_fork: event record failed");
  if (cudaStreamWaitEvent(P->side[i], e, 0) != cudaSuccess) return set_error("stream_fork: wait failed");
  *side = P->side[i];
  return 0;
}

int stream_join(void* main_s, int i) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (i < 0 || i >= kSide) return set_error("stream_join: bad index");
  if (!g_side_enabled) return 0;
  SidePool* P = side_pool();
  if (!P) return 1;
  cudaEvent_t e = next_event(*P);
  if (cudaEventRecord(e, P->side[i]) != cudaSuccess) return set_error("stream_join: event record failed");
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
