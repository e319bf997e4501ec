// Fused training loss of the reference trainer (MultiClassesTrainer.compute_loss with use_mse=True,
// /root/reference/trainer.py:935-966; flexible_cross_entropy :853-916; restated op by op in videocad_b200/loss.py):
//
//   loss = 2 * CE_w(cmds, tgt[:,0]; ignore -1) + sum_i w[label(i)] * mean_{selected rows} per_row_i
//   per_row_i = -(1/count) * sum_{c in [t, hi]} log_softmax(params[r,i,:])[c],  hi = min(t + tol_i - 1, NV-1), count = hi - t + 1
//   selected  = target != -1 and argmax outside [t, hi];  empty selection -> 0;  NaN term -> dropped
//
// The torch restatement launches ~300 small kernels per step (forward + autograd backward); here: one row pass
// (one CTA per (row, parameter): max/argmax, log-sum-exp, window sum), one single-CTA finalize (fixed-order reductions, so
// the loss is run-to-run deterministic) and one gradient pass writing d loss / d logits directly.
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

constexpr int LT = 128;  // threads per CTA

struct LossWs {
  float *lse, *sel, *rowloss, *nwin_over_count;  // [R*NP]
  float *cnum, *cden, *clse;                      // [R]
  float* scal;                                    // coef[NP], coef_cmd, (pad to 16)
  float *pred, *cpred;                            // argmax of every parameter head [R*NP] and of the command head [R] (metrics)
};

__host__ __device__ inline LossWs carve(float* ws, int R, int NP) {
  LossWs w;
  const size_t rp = (size_t)R * NP;
  w.lse = ws; w.sel = ws + rp; w.rowloss = ws + 2 * rp; w.nwin_over_count = ws + 3 * rp;
  w.cnum = ws + 4 * rp; w.cden = w.cnum + R; w.clse = w.cden + R;
  w.scal = w.clse + R;
  w.pred = w.scal + 16; w.cpred = w.pred + rp;
  return w;
}

struct MaxIdx {
  float v; int i;
};
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {  // larger value, then lower index (first occurrence)
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < LT / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ void window_of(const LossCfg& c, int i, float target, bool& valid, long long& t, long long& hi, float& count) {
  const long long tg = (long long)target;  // .long(): truncation
  valid = tg != -1;
  // allowed classes = { clamp(target + o, 0, NV - 1) : 0 <= o < tolerance } (trainer.py:880-905): a contiguous window whose two
  // ends are clamped separately, so an out-of-range target still has a one-class window (never empty, count >= 1)
  const long long top = c.NV - 1;
  t = tg < 0 ? 0 : (tg > top ? top : tg);
  hi = tg + (c.tolerance[i] - 1);
  hi = hi < 0 ? 0 : (hi > top ? top : hi);
  count = (float)(hi - t + 1);
}

__global__ void __launch_bounds__(LT) loss_rows_kernel(const LossCfg c, const float* __restrict__ cmds, const float* __restrict__ params,
                                                       const float* __restrict__ targets, LossWs w) {
  pdl_grid_sync();
  __shared__ float red[LT / 32];
  __shared__ float redv[LT / 32];
  __shared__ int redi[LT / 32];
  const int r = blockIdx.x, i = blockIdx.y;
  const int ldt = 1 + c.NP;
  if (i == c.NP) {  // command head: weighted cross entropy, NC <= 16 classes
    if (threadIdx.x == 0) {
      const float* z = cmds + (size_t)r * c.NC;
      const long long tg = (long long)targets[(size_t)r * ldt];
      const bool valid = tg >= 0 && tg < c.NC;
      float m = -INFINITY;
      int am = 0;
      for (int k = 0; k < c.NC; ++k) if (z[k] > m) { m = z[k]; am = k; }  // first maximum, as torch.argmax
      w.cpred[r] = (float)am;
      float s = 0.f;
      for (int k = 0; k < c.NC; ++k) s += expf(z[k] - m);
      const float lse = m + logf(s);
      const float wt = valid ? c.cmd_w[tg] : 0.f;
      w.clse[r] = lse;
      w.cden[r] = wt;
      w.cnum[r] = valid ? wt * (lse - z[tg]) : 0.f;
    }
    return;
  }
  const float* z = params + ((size_t)r * c.NP + i) * c.NV;
  bool valid; long long t, hi; float count;
  window_of(c, i, targets[(size_t)r * ldt + 1 + i], valid, t, hi, count);
  MaxIdx best{-INFINITY, 0x7fffffff};
  for (int k = threadIdx.x; k < c.NV; k += LT) best = better(best, MaxIdx{z[k], k});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxIdx other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    best = better(best, other);
  }
  if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = best.v; redi[threadIdx.x >> 5] = best.i; }
  __syncthreads();
  best = MaxIdx{redv[0], redi[0]};
#pragma unroll
  for (int k = 1; k < LT / 32; ++k) best = better(best, MaxIdx{redv[k], redi[k]});
  float se = 0.f, sw = 0.f, nw = 0.f;
  for (int k = threadIdx.x; k < c.NV; k += LT) {
    const float v = z[k];
    se += expf(v - best.v);
    if (k >= t && k <= hi) { sw += v; nw += 1.f; }
  }
  se = block_sum(se, red);
  sw = block_sum(sw, red);
  nw = block_sum(nw, red);
  if (threadIdx.x == 0) {
    const float lse = best.v + logf(se);
    const bool in_window = best.i >= t && best.i <= hi;
    const float sel = (valid && !in_window) ? 1.f : 0.f;
    const float per_row = -(sw - nw * lse) / count;  // count >= 1 (window_of)
    const size_t o = (size_t)r * c.NP + i;
    w.lse[o] = lse;
    w.sel[o] = sel;
    w.rowloss[o] = per_row * sel;
    w.nwin_over_count[o] = nw / count;
    w.pred[o] = (float)best.i;
  }
}

// Metrics of MultiClassesTrainer.compute_loss (/root/reference/trainer.py:968-1061): integer counts from the argmax predictions
// of the row pass.  One CTA, shared-memory integer atomics (exact, order-independent).  Counter layout: see VC_METRIC_* in
// include/videocad_b200.h.
__global__ void __launch_bounds__(256) loss_metrics_kernel(const LossCfg c, const float* __restrict__ targets, LossWs w, int T, int topk,
                                                           const vc_metrics_cfg mc, unsigned long long* __restrict__ out) {
  pdl_grid_sync();
  __shared__ unsigned long long cnt[VC_METRIC_COUNT];
  for (int k = threadIdx.x; k < VC_METRIC_COUNT; k += 256) cnt[k] = 0ull;
  __syncthreads();
  const int ldt = 1 + c.NP;
  for (int r = threadIdx.x; r < c.R; r += 256) {
    const long long tc = (long long)targets[(size_t)r * ldt];
    const long long pc = (long long)w.cpred[r];
    const bool cmd_mask = tc != -1;
    const bool cmd_ok = cmd_mask && pc == tc;
    const bool in_topk = (r % T) < topk;
    if (cmd_mask) atomicAdd(&cnt[VC_METRIC_TOTAL], 1ull);
    if (cmd_ok) atomicAdd(&cnt[VC_METRIC_CORRECT], 1ull);
    if (tc >= 0 && tc < c.NC) {
      atomicAdd(&cnt[VC_METRIC_CMD_COUNTS + tc], 1ull);
      if (pc == tc) atomicAdd(&cnt[VC_METRIC_CMD_CORRECTS + tc], 1ull);
    }
    if (in_topk && cmd_mask) atomicAdd(&cnt[VC_METRIC_CMD_COUNTS_TOPK], 1ull);
    if (in_topk && cmd_ok) atomicAdd(&cnt[VC_METRIC_CMD_CORRECT_TOPK], 1ull);
    for (int i = 0; i < c.NP; ++i) {
      const long long tp = (long long)targets[(size_t)r * ldt + 1 + i];
      if (!cmd_mask || tp == -1) continue;  // param_mask
      atomicAdd(&cnt[VC_METRIC_PARAM_COUNTS + i], 1ull);
      atomicAdd(&cnt[VC_METRIC_TOTAL], 1ull);
      if (in_topk) atomicAdd(&cnt[VC_METRIC_PARAM_COUNTS_TOPK], 1ull);
      if (!cmd_ok) continue;                // params_mask additionally needs the command to be right
      const long long diff = (long long)w.pred[(size_t)r * c.NP + i] - tp;
      const bool ok = mc.above[i] ? (diff >= 0 && diff < mc.tolerance[i]) : ((diff < 0 ? -diff : diff) < mc.abs_tolerance);
      if (ok) {
        atomicAdd(&cnt[VC_METRIC_PARAM_CORRECTS + i], 1ull);
        atomicAdd(&cnt[VC_METRIC_CORRECT], 1ull);
        if (in_topk) atomicAdd(&cnt[VC_METRIC_PARAM_CORRECT_TOPK], 1ull);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < VC_METRIC_COUNT; k += 256) out[k] = cnt[k];
}

// one CTA: fixed-order reductions over the rows, the scalar loss, and the per-term gradient coefficients
__global__ void __launch_bounds__(256) loss_finalize_kernel(const LossCfg c, LossWs w, float* __restrict__ loss_out) {
  pdl_grid_sync();
  __shared__ float red[256];
  __shared__ float terms[VC_LOSS_MAX_PARAMS + 2];
  auto reduce = [&](const float* src, int stride, int off) -> float {
    float a = 0.f;
    for (int r = threadIdx.x; r < c.R; r += 256) a += src[(size_t)r * stride + off];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    const float out = red[0];
    __syncthreads();
    return out;
  };
  for (int i = 0; i < c.NP; ++i) {
    const float S = reduce(w.rowloss, c.NP, i);
    const float n = reduce(w.sel, c.NP, i);
    if (threadIdx.x == 0) {
      const float denom = fmaxf(n, 1.f);
      float term = S / denom;
      float coef = c.cmd_w[c.param_to_label[i]] / denom;
      if (isnan(term)) { term = 0.f; coef = 0.f; }
      terms[i] = term * c.cmd_w[c.param_to_label[i]];
      w.scal[i] = coef;
    }
  }
  const float num = reduce(w.cnum, 1, 0);
  const float den = reduce(w.cden, 1, 0);
  if (threadIdx.x == 0) {
    float loss = 2.f * (num / den);
    for (int i = 0; i < c.NP; ++i) loss += terms[i];
    w.scal[c.NP] = 2.f / den;
    loss_out[0] = loss;
  }
}

__global__ void __launch_bounds__(LT) loss_grad_kernel(const LossCfg c, const float* __restrict__ cmds, const float* __restrict__ params,
                                                       const float* __restrict__ targets, LossWs w, const float* __restrict__ upstream,
                                                       float* __restrict__ dcmds, float* __restrict__ dparams) {
  pdl_grid_sync();
  const int r = blockIdx.x, i = blockIdx.y;
  const int ldt = 1 + c.NP;
  const float up = upstream[0];
  if (i == c.NP) {
    if (threadIdx.x < c.NC) {
      const int k = threadIdx.x;
      const float* z = cmds + (size_t)r * c.NC;
      const long long tg = (long long)targets[(size_t)r * ldt];
      const float wt = w.cden[r];  // 0 for ignored rows
      const float g = wt * w.scal[c.NP] * up * (expf(z[k] - w.clse[r]) - (k == tg ? 1.f : 0.f));
      dcmds[(size_t)r * c.NC + k] = wt != 0.f ? g : 0.f;
    }
    return;
  }
  const size_t o = (size_t)r * c.NP + i;
  const float* z = params + o * c.NV;
  float* dz = dparams + o * c.NV;
  const float coef = w.scal[i] * w.sel[o] * up;
  if (coef == 0.f) {  // unselected row or dropped term
    for (int k = threadIdx.x; k < c.NV; k += LT) dz[k] = 0.f;
    return;
  }
  bool valid; long long t, hi; float count;
  window_of(c, i, targets[(size_t)r * ldt + 1 + i], valid, t, hi, count);
  const float lse = w.lse[o], a = w.nwin_over_count[o], inv = 1.f / count;
  for (int k = threadIdx.x; k < c.NV; k += LT)
    dz[k] = coef * (a * expf(z[k] - lse) - ((k >= t && k <= hi) ? inv : 0.f));
}

int check_cfg(const LossCfg& c) {
  if (c.R <= 0 || c.NV <= 0) return set_error("loss: empty problem");
  if (c.NP < 0 || c.NP > VC_LOSS_MAX_PARAMS || c.NC <= 0 || c.NC > VC_LOSS_MAX_CLASSES) return set_error("loss: NP <= 8 and NC <= 16 required");
  for (int i = 0; i < c.NP; ++i)
    if (c.param_to_label[i] < 0 || c.param_to_label[i] >= c.NC) return set_error("loss: param_to_label out of range");
  return 0;
}

}  // namespace

size_t loss_workspace_floats(int R, int NP) { return (size_t)5 * R * NP + (size_t)4 * R + 16; }

int loss_metrics(const LossCfg& cfg, const vc_metrics_cfg& mc, const float* targets, const float* ws, int T, int64_t* counts, stream_t s) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!targets || !ws || !counts) return set_error("loss_metrics: null argument");
  if (T <= 0 || cfg.R % T != 0) return set_error("loss_metrics: R must be a multiple of T");
  const LossWs w = carve(const_cast<float*>(ws), cfg.R, cfg.NP);
  VC_LAUNCH((loss_metrics_kernel), 1, 256, 0, reinterpret_cast<cudaStream_t>(s), cfg, targets, w, T, mc.topk, mc,
            reinterpret_cast<unsigned long long*>(counts));
  return check_launch("loss_metrics_kernel");
}

int loss_forward(const LossCfg& cfg, const float* cmds, const float* params, const float* targets, float* ws, float* loss_out, stream_t s) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!cmds || !params || !targets || !ws || !loss_out) return set_error("loss_forward: null argument");
  const LossWs w = carve(ws, cfg.R, cfg.NP);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  VC_LAUNCH((loss_rows_kernel), dim3(cfg.R, cfg.NP + 1), LT, 0, st, cfg, cmds, params, targets, w);
  if (int rc = check_launch("loss_rows_kernel")) return rc;
  VC_LAUNCH((loss_finalize_kernel), 1, 256, 0, st, cfg, w, loss_out);
  return check_launch("loss_finalize_kernel");
}

int loss_backward(const LossCfg& cfg, const float* cmds, const float* params, const float* targets, const float* ws,
                  const float* upstream, float* dcmds, float* dparams, stream_t s) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!cmds || !params || !targets || !ws || !upstream || !dcmds || !dparams) return set_error("loss_backward: null argument");
  const LossWs w = carve(const_cast<float*>(ws), cfg.R, cfg.NP);
  VC_LAUNCH((loss_grad_kernel), dim3(cfg.R, cfg.NP + 1), LT, 0, reinterpret_cast<cudaStream_t>(s), cfg, cmds, params, targets, w, upstream, dcmds,
                                                                                        dparams);
  return check_launch("loss_grad_kernel");
}

}  // namespace vck
