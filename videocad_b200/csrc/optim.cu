// Fused optimizer step of the reference trainer: clip_grad_norm_(max_norm) + Adam (trainer.py:493-494) over a few large
// flat tensors (videocad_b200 keeps each model segment's weights in one flat parameter).  torch's foreach Adam makes
// about ten passes over parameter-sized memory and clip_grad_norm_ two more; here: one read of the gradients for the norm,
// and one pass that reads g, p, m, v and writes g, p, m, v.  Purely HBM-bound: 9 x 4 bytes per parameter per step.
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

constexpr int kPartials = 1024;  // per-tensor partial sums of squares
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[kThreads / 32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// scal[0] = total norm, scal[1] = clip coefficient (clamped to 1)
__global__ void __launch_bounds__(kThreads) clip_finalize_kernel(const float* __restrict__ partials, int count, float max_norm,
                                                                  float* __restrict__ scal, float* __restrict__ total_norm_out) {
  pdl_grid_sync();
  __shared__ float red[kThreads];
  float a = 0.f;
  for (int i = threadIdx.x; i < count; i += kThreads) a += partials[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float total = sqrtf(red[0]);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.f);
    scal[0] = total;
    scal[1] = coef;
    if (total_norm_out) total_norm_out[0] = total;
  }
}

struct AdamScalars {
  float beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt;  // omb = 1 - beta, rounded from double as torch does
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float coef, const AdamScalars& a) {
  g *= coef;                                             // clip_grad_norm_: grads scaled in place
  m = m + a.omb1 * (g - m);                               // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + a.omb2 * g * g;                       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;     // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - a.step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(kThreads) adam_update_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                                float* __restrict__ v, long long n, const float* __restrict__ scal,
                                                                AdamScalars a) {
  pdl_grid_sync();
  const float coef = scal[1];
  const long long n4 = n >> 2;
  float4 *p4 = reinterpret_cast<float4*>(p), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
         *v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, coef, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, coef, a);
    p4[i] = pp; g4[i] = gg; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) adam_one(p[i], g[i], m[i], v[i], coef, a);
  }
}

}  // namespace

size_t clip_adam_scratch_floats() { return (size_t)VC_ADAM_MAX_TENSORS * kPartials + 8; }

int clip_adam_step(const vc_adam_tensor* t, int nt, double beta1, double beta2, double eps, double max_norm, int64_t step, float* scratch,
                   float* total_norm_out, stream_t s) {
  if (!t || nt <= 0 || nt > VC_ADAM_MAX_TENSORS) return set_error("clip_adam_step: 1..16 tensors");
  if (!scratch || step < 1) return set_error("clip_adam_step: scratch required, step >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  float* scal = scratch + (size_t)VC_ADAM_MAX_TENSORS * kPartials;
  int count = 0;
  for (int i = 0; i < nt; ++i) {
    if (!t[i].p || !t[i].g || !t[i].m || !t[i].v || t[i].n <= 0) return set_error("clip_adam_step: null tensor");
    if ((reinterpret_cast<uintptr_t>(t[i].p) | reinterpret_cast<uintptr_t>(t[i].g) | reinterpret_cast<uintptr_t>(t[i].m) |
         reinterpret_cast<uintptr_t>(t[i].v)) & 15)
      return set_error("clip_adam_step: tensors must be 16-byte aligned");
    long long blocks = (t[i].n / 4 + kThreads * 4 - 1) / (kThreads * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > kPartials) blocks = kPartials;
    VC_LAUNCH((sqnorm_partial_kernel), (unsigned)blocks, kThreads, 0, st, t[i].g, t[i].n, scratch + count);
    if (int rc = check_launch("sqnorm_partial_kernel")) return rc;
    count += (int)blocks;
  }
  VC_LAUNCH((clip_finalize_kernel), 1, kThreads, 0, st, scratch, count, (float)max_norm, scal, total_norm_out);
  if (int rc = check_launch("clip_finalize_kernel")) return rc;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  for (int i = 0; i < nt; ++i) {
    AdamScalars a;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.step_size = (float)((double)t[i].lr / bc1);
    a.bc2_sqrt = (float)sqrt(bc2);
    long long blocks = (t[i].n / 4 + kThreads * 2 - 1) / (kThreads * 2);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    VC_LAUNCH((adam_update_kernel), (unsigned)blocks, kThreads, 0, st, t[i].p, t[i].g, t[i].m, t[i].v, t[i].n, scal, a);
    if (int rc = check_launch("adam_update_kernel")) return rc;
  }
  return 0;
}

}  // namespace vck
