// Kernels of ONE position of the key/value-cached rollout (sequential_inference with action feedback,
// /root/reference/model/autoregressive_transformer.py:222-275) for a handful of sequences (B <= 16 per GPU: BASELINE config C4 is
// 8 per GPU).  A decode step pushes one token per sequence through the 8 post-norm decoder layers
// (torch.nn.TransformerDecoderLayer as built at autoregressive_transformer.py:54-62) and the two heads (:217-218), then picks
// the next action (argmax + action mask + normalise, :91-118, :256-266).
//
// With B rows the step is bound by reading every weight ONCE (293 MB of fp32 at H = 1024 + the cached keys/values: HBM) and,
// before that, by the length of its dependency chain: 8 layers x 8 dependent phases.  Design:
//   * dec_gemv_kernel  out[B, N] = act(x W^T + b) + residual for one Linear.  Each CTA owns `cols` output columns: their fp32
//     weight rows are ONE contiguous slab, fetched by a single TMA bulk copy (cp.async.bulk -> mbarrier) that is issued BEFORE
//     griddepcontrol.wait -- weights never change during a rollout, so the slab streams from HBM while the previous kernel
//     of the chain is still running (programmatic dependent launch).  Each thread owns one float4 slot of the K dimension for all
//     B rows (x lives in registers); a halving butterfly reduces the B partial sums of a column with log-many shuffles.
//     The input rows are built on load: plain rows, LayerNorm of the previous sub-layer's sum (two-pass statistics, as
//     ln_fwd_kernel), tanh(embed_action(a) + E[t]) for the first layer, or the merge of the attention kernel's split partials
//     -- so LayerNorm, the token embedding and the softmax merge cost no launches of their own.
//   * dec_attn_kernel  one query row per (sequence, head) against the cached keys/values: the keys are split over `nsplit` CTAs
//     and 8 warps each (online softmax per warp, merged in shared memory); partial (max, sum, acc) go to global and are merged by
//     the consuming out-projection GEMV.
//   * dec_select_kernel  final LayerNorm + command head + argmax of both heads + action mask + normalisation: the next action is
//     written on the device and the position counter advances there, so the 186-step loop never returns to the host and ONE
//     captured CUDA graph serves every position.
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

constexpr int DG_THREADS = 256;
constexpr int DG_WARPS = DG_THREADS / 32;
constexpr float DEC_LN_EPS = 1e-5f;

inline cudaStream_t cs(stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// global -> shared bulk copy (TMA, no tensor map): completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Sum v[0..MM) over the 32 lanes of a warp with MM - 1 + log2(32 / MM) shuffles instead of 5 * MM: each halving step sends one
// half of the values to the partner lane and keeps the other.  Returns the total of row (lane >> log2(32 / MM)) -- lanes that
// share those top bits hold the same value.
template <int MM>
__device__ __forceinline__ float reduce_rows(float (&v)[MM], int lane) {
  static_assert(MM == 8 || MM == 16, "8 or 16 rows");
  if constexpr (MM == 16) {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  constexpr int o8 = MM == 16 ? 8 : 16, o4 = o8 / 2, o2 = o4 / 2;
  {
    const bool up = (lane & o8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o8);
    }
  }
  {
    const bool up = (lane & o4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o4);
    }
  }
  float r;
  {
    const bool up = (lane & o2) != 0;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    r = keep + __shfl_xor_sync(0xffffffffu, send, o2);
  }
#pragma unroll
  for (int o = o2 / 2; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// per-row totals over the whole CTA: every thread contributes v[m]; afterwards out[m] (shared) holds the sum for m < MM
template <int MM>
__device__ __forceinline__ void block_row_sums(float (&v)[MM], float* red /* [DG_WARPS][MM] */, float* out /* [MM] */, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kShift = MM == 8 ? 2 : 1;
  const float r = reduce_rows<MM>(v, lane);
  if ((lane & ((1 << kShift) - 1)) == 0) red[warp * MM + (lane >> kShift)] = r;
  __syncthreads();
  if (tid < MM) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DG_WARPS; ++w) s += red[w * MM + tid];
    out[tid] = s;
  }
  __syncthreads();
}

template <int MM>
__global__ void __launch_bounds__(DG_THREADS) dec_gemv_kernel(const DecGemv a) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t dg_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(dg_smem);
  float* wsm = reinterpret_cast<float*>(dg_smem + 16);               // [cols_per_cta][K] weight slab
  float* part = wsm + (size_t)a.cols_per_cta * a.K;                    // [cols_per_cta][DG_WARPS][MM] per-warp partial sums
  float* red = part + (size_t)a.cols_per_cta * DG_WARPS * MM;          // [DG_WARPS][MM]
  float* stat = red + DG_WARPS * MM;                                   // [MM]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * a.cols_per_cta;
  const int cols = min(a.cols_per_cta, a.N - n0);
  const int K = a.K, M = a.M;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    // the CTA's weight rows [n0, n0 + cols) x K are contiguous in the row-major [N, K] weight: one bulk copy.  Issued before the
    // dependency wait: weights are constants of the rollout.
    const uint32_t bytes = (uint32_t)cols * (uint32_t)K * 4u;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(wsm, a.W + (size_t)n0 * K, bytes, bar);
  }
  __syncthreads();  // the initialised barrier is visible to every thread that will wait on it
  pdl_wait();
  const int t = a.t_ptr ? *a.t_ptr : 0;
  // ---------------------------------------------------------------- input rows: this thread's float4 slot of K, all rows
  const int k0 = tid * 4;
  const bool active = k0 < K;
  float4 x[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) x[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.in_mode == VC_DEC_IN_PLAIN || a.in_mode == VC_DEC_IN_LN) {
    if (active) {
#pragma unroll
      for (int m = 0; m < MM; ++m)
        if (m < M) x[m] = *reinterpret_cast<const float4*>(a.x + (size_t)m * a.ldx + k0);
    }
    if (a.in_mode == VC_DEC_IN_LN) {
      // LayerNorm over K with two-pass statistics (biased variance, eps inside the square root), as ln_fwd_kernel
      float s[MM];
#pragma unroll
      for (int m = 0; m < MM; ++m) s[m] = x[m].x + x[m].y + x[m].z + x[m].w;
      block_row_sums<MM>(s, red, stat, tid);
      float mean[MM];
#pragma unroll
      for (int m = 0; m < MM; ++m) mean[m] = stat[m] / (float)K;
      __syncthreads();  // stat is rewritten below
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const float d0 = x[m].x - mean[m], d1 = x[m].y - mean[m], d2 = x[m].z - mean[m], d3 = x[m].w - mean[m];
        s[m] = active ? d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 : 0.f;
      }
      block_row_sums<MM>(s, red, stat, tid);
      if (active) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + k0));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(a.beta + k0));
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          const float rstd = rsqrtf(stat[m] / (float)K + DEC_LN_EPS);
          x[m].x = (x[m].x - mean[m]) * rstd * g.x + bt.x;
          x[m].y = (x[m].y - mean[m]) * rstd * g.y + bt.y;
          x[m].z = (x[m].z - mean[m]) * rstd * g.z + bt.z;
          x[m].w = (x[m].w - mean[m]) * rstd * g.w + bt.w;
        }
      }
    }
  } else if (a.in_mode == VC_DEC_IN_EMBED) {
    // token of this position: tanh(embed_action(a_t) + E[t])  (autoregressive_transformer.py:110-113, 176-178)
    if (active) {
      float bias4[4], e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) bias4[j] = a.emb_b[k0 + j];
      if (a.emb_E) {
        const float4 e = *reinterpret_cast<const float4*>(a.emb_E + (size_t)t * K + k0);
        e4[0] = e.x; e4[1] = e.y; e4[2] = e.z; e4[3] = e.w;
      }
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        if (m < M) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float acc = bias4[j];
            for (int i = 0; i < a.act_dim; ++i) acc += a.actions[(size_t)m * a.act_dim + i] * a.emb_W[(size_t)(k0 + j) * a.act_dim + i];
            v[j] = tanhf(acc + e4[j]);
          }
          x[m] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  } else {  // VC_DEC_IN_ATTN: merge the attention kernel's split partials (online-softmax combine), K = nh * dh
    if (active) {
      const int h = k0 / a.dh, d = k0 - h * a.dh;
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        if (m < M) {
          const size_t base = ((size_t)m * a.nh + h) * a.nsplit;
          float mx = -INFINITY;
          for (int s = 0; s < a.nsplit; ++s) mx = fmaxf(mx, a.part_ml[(base + s) * 2]);
          float den = 0.f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s = 0; s < a.nsplit; ++s) {
            const float ms = a.part_ml[(base + s) * 2];
            if (ms == -INFINITY) continue;  // a split that saw no key
            const float w = __expf(ms - mx);
            den += w * a.part_ml[(base + s) * 2 + 1];
            const float4 o = *reinterpret_cast<const float4*>(a.part_o + (base + s) * a.dh + d);
            acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
          }
          const float inv = 1.0f / den;
          x[m] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        }
      }
    }
  }
  if (a.x_out != nullptr && blockIdx.x == 0 && active) {  // the rows just built are the residual operand of a later kernel
#pragma unroll
    for (int m = 0; m < MM; ++m)
      if (m < M) *reinterpret_cast<float4*>(a.x_out + (size_t)m * K + k0) = x[m];
  }
  // ---------------------------------------------------------------- dot products against the weight slab
  mbar_wait(bar, 0);
  constexpr int kShift = MM == 8 ? 2 : 1;
  for (int c = 0; c < cols; ++c) {
    float p[MM];
    if (active) {
      const float4 w4 = *reinterpret_cast<const float4*>(wsm + (size_t)c * K + k0);
#pragma unroll
      for (int m = 0; m < MM; ++m) p[m] = fmaf(w4.x, x[m].x, fmaf(w4.y, x[m].y, fmaf(w4.z, x[m].z, w4.w * x[m].w)));
    } else {
#pragma unroll
      for (int m = 0; m < MM; ++m) p[m] = 0.f;
    }
    const float r = reduce_rows<MM>(p, lane);
    if ((lane & ((1 << kShift) - 1)) == 0) part[((size_t)c * DG_WARPS + warp) * MM + (lane >> kShift)] = r;
  }
  __syncthreads();
  // ---------------------------------------------------------------- epilogue: consecutive threads <-> consecutive columns
  for (int idx = tid; idx < cols * M; idx += DG_THREADS) {
    const int m = idx / cols, c = idx - m * cols;
    const int n = n0 + c;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < DG_WARPS; ++w) v += part[((size_t)c * DG_WARPS + w) * MM + m];
    if (a.bias) v += __ldg(a.bias + n);
    v = apply_act(v, a.act);
    if (a.residual) v += a.residual[(size_t)m * a.ld_res + n];
    a.out[(size_t)m * a.out_row_stride + (size_t)t * a.out_t_stride + n] = v;
  }
}

// ------------------------------------------------------------------------------------------------- attention of one query row
template <int DPL>
__global__ void __launch_bounds__(DG_THREADS) dec_attn_kernel(const DecAttn a) {
  pdl_grid_sync();
  __shared__ float sm_m[DG_WARPS], sm_l[DG_WARPS];
  __shared__ float sm_acc[DG_WARPS][32 * DPL];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int t = *a.t_ptr;
  const int j_lo = a.window > 0 ? max(0, t - a.window + 1) : 0;
  const int nkeys = t - j_lo + 1;
  const int d0 = h * a.dh + lane * DPL;
  float q[DPL];
  {
    const float* qp = a.q + (size_t)b * a.q_bstride + (size_t)t * a.q_tstride + d0;
#pragma unroll
    for (int i = 0; i < DPL; ++i) q[i] = qp[i] * a.scale;
  }
  float mrun = -INFINITY, lrun = 0.f, acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  const float* kb = a.k + (size_t)b * a.kv_bstride + d0;
  const float* vb = a.v + (size_t)b * a.kv_bstride + d0;
  const int stride = a.nsplit * DG_WARPS;
  for (int i = s * DG_WARPS + warp; i < nkeys; i += 2 * stride) {
    // two keys per iteration: both keys' and values' loads are in flight before the first reduction
    const int i1 = i + stride;
    const bool has1 = i1 < nkeys;
    const float* k0p = kb + (size_t)(j_lo + i) * a.kv_rstride;
    const float* v0p = vb + (size_t)(j_lo + i) * a.kv_rstride;
    const float* k1p = kb + (size_t)(j_lo + (has1 ? i1 : i)) * a.kv_rstride;
    const float* v1p = vb + (size_t)(j_lo + (has1 ? i1 : i)) * a.kv_rstride;
    float k0[DPL], k1[DPL], v0[DPL], v1[DPL];
#pragma unroll
    for (int e = 0; e < DPL; ++e) { k0[e] = k0p[e]; k1[e] = k1p[e]; }
#pragma unroll
    for (int e = 0; e < DPL; ++e) { v0[e] = v0p[e]; v1[e] = v1p[e]; }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int e = 0; e < DPL; ++e) { s0 = fmaf(q[e], k0[e], s0); s1 = fmaf(q[e], k1[e], s1); }
    s0 = warp_sum(s0);
    s1 = has1 ? warp_sum(s1) : -INFINITY;  // has1 is warp-uniform
    const float mnew = fmaxf(mrun, fmaxf(s0, s1));
    const float corr = __expf(mrun - mnew), p0 = __expf(s0 - mnew), p1 = has1 ? __expf(s1 - mnew) : 0.f;
    lrun = lrun * corr + p0 + p1;
#pragma unroll
    for (int e = 0; e < DPL; ++e) acc[e] = acc[e] * corr + p0 * v0[e] + p1 * v1[e];
    mrun = mnew;
  }
  if (lane == 0) { sm_m[warp] = mrun; sm_l[warp] = lrun; }
#pragma unroll
  for (int e = 0; e < DPL; ++e) sm_acc[warp][lane * DPL + e] = acc[e];
  __syncthreads();
  const size_t pbase = ((size_t)b * a.nh + h) * a.nsplit + s;
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < DG_WARPS; ++w) mx = fmaxf(mx, sm_m[w]);
  if (tid < a.dh) {
    float o = 0.f;
    if (mx != -INFINITY) {
#pragma unroll
      for (int w = 0; w < DG_WARPS; ++w)
        if (sm_m[w] != -INFINITY) o += __expf(sm_m[w] - mx) * sm_acc[w][tid];
    }
    a.part_o[pbase * a.dh + tid] = o;
  }
  if (tid == 0) {
    float l = 0.f;
    if (mx != -INFINITY) {
#pragma unroll
      for (int w = 0; w < DG_WARPS; ++w)
        if (sm_m[w] != -INFINITY) l += __expf(sm_m[w] - mx) * sm_l[w];
    }
    a.part_ml[pbase * 2] = mx;
    a.part_ml[pbase * 2 + 1] = l;
  }
}

// ------------------------------------------------------------------------------------------------- heads' argmax + feedback
// command -> which of the 6 parameters it carries (autoregressive_transformer.py:83-89)
__constant__ int kActionMask[5][6] = {{1, 1, 0, 0, 0, 0}, {0, 0, 1, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 0}};

__global__ void __launch_bounds__(DG_THREADS) dec_select_kernel(const DecSelect a) {
  pdl_grid_sync();
  extern __shared__ float ds_x[];  // [H] normalised row
  __shared__ float red[DG_WARPS];
  __shared__ float bcast;
  __shared__ float logit_c[8];
  __shared__ int arg_p[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, H = a.H;
  const int t = *a.t_ptr;
  // final LayerNorm of this sequence's row (two-pass statistics)
  const float* y = a.y + (size_t)b * H;
  float s = 0.f;
  for (int k = tid; k < H; k += DG_THREADS) s += y[k];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < DG_WARPS; ++w) v += red[w]; bcast = v / (float)H; }
  __syncthreads();
  const float mean = bcast;
  float qv = 0.f;
  for (int k = tid; k < H; k += DG_THREADS) { const float d = y[k] - mean; qv += d * d; }
  qv = warp_sum(qv);
  __syncthreads();
  if (lane == 0) red[warp] = qv;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < DG_WARPS; ++w) v += red[w]; bcast = rsqrtf(v / (float)H + DEC_LN_EPS); }
  __syncthreads();
  const float rstd = bcast;
  for (int k = tid; k < H; k += DG_THREADS) ds_x[k] = (y[k] - mean) * rstd * a.gamma[k] + a.beta[k];
  __syncthreads();
  // command head: warp w <-> class w
  if (warp < a.NC) {
    const float* wr = a.Wc + (size_t)warp * H;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(wr[k], ds_x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + a.bc[warp];
      logit_c[warp] = v;
      a.cmds_all[((size_t)b * a.T + t) * a.NC + warp] = v;
    }
  }
  // parameter heads: warp i <-> parameter i, argmax over its NV logits (written by the head GEMV); first index on ties
  if (warp < a.NPAR) {
    const float* z = a.params_all + ((size_t)b * a.T + t) * ((size_t)a.NPAR * a.NV) + (size_t)warp * a.NV;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < a.NV; k += 32) {
      const float v = z[k];
      if (v > bv) { bv = v; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) arg_p[warp] = bi;
  }
  __syncthreads();
  if (tid == 0 && a.action_next != nullptr) {
    int cmd = 0;
    float cv = logit_c[0];
    for (int c = 1; c < a.NC; ++c)
      if (logit_c[c] > cv) { cv = logit_c[c]; cmd = c; }
    // apply_action_mask (autoregressive_transformer.py:91-108): parameters the command does not carry become -1; parameter 3
    // survives only if 200 <= parameter 2 < 250.  Then normalize_actions (:115-118): command / 4, parameters / 1000.
    float par[8];
    for (int i = 0; i < a.NPAR; ++i) par[i] = (cmd < 5 && i < 6 && kActionMask[cmd][i]) ? (float)arg_p[i] : -1.0f;
    if (a.NPAR > 3 && !(par[2] >= 200.f && par[2] < 250.f)) par[3] = -1.0f;
    float* out = a.action_next + (size_t)b * (1 + a.NPAR);
    out[0] = (float)cmd / 4.0f;
    for (int i = 0; i < a.NPAR; ++i) out[1 + i] = par[i] / 1000.0f;
  }
  // the last CTA to finish advances the position counter: every CTA read `t` before it arrives here
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(a.done_ctr, 1u);
    if (done == gridDim.x - 1) {
      *a.done_ctr = 0u;
      *a.t_ptr = t + 1;
      __threadfence();
    }
  }
}

}  // namespace

int dec_gemv(const DecGemv& g, stream_t s) {
  if (g.M <= 0 || g.M > 16) return set_error("dec_gemv: 1..16 rows");
  if (g.K <= 0 || g.K % 128 != 0 || g.K > 4 * DG_THREADS) return set_error("dec_gemv: K must be a multiple of 128 and <= 1024");
  if (g.N <= 0 || !g.W || !g.out) return set_error("dec_gemv: null weight / output or empty problem");
  if (g.in_mode == VC_DEC_IN_ATTN && (g.nh * g.dh != g.K || g.dh % 4 != 0 || !g.part_o || !g.part_ml || g.nsplit < 1))
    return set_error("dec_gemv: bad attention-partial input");
  if ((g.in_mode == VC_DEC_IN_PLAIN || g.in_mode == VC_DEC_IN_LN) && (!g.x || g.ldx % 4 != 0)) return set_error("dec_gemv: bad input rows");
  if (g.in_mode == VC_DEC_IN_LN && (!g.gamma || !g.beta)) return set_error("dec_gemv: LayerNorm input needs gamma / beta");
  if (g.in_mode == VC_DEC_IN_EMBED && (!g.actions || !g.emb_W || !g.emb_b)) return set_error("dec_gemv: bad embedding input");
  const int MM = g.M <= 8 ? 8 : 16;
  DecGemv a = g;
  // columns per CTA: one wave of CTAs over the 148 SMs, multiple of 4, weight slab <= 192 KB of shared memory
  int cols = (g.N + 147) / 148;
  cols = (cols + 3) / 4 * 4;
  const int cap = (192 * 1024) / (g.K * 4) / 4 * 4;
  if (cols > cap) cols = cap;
  if (cols < 4) cols = 4;
  a.cols_per_cta = cols;
  const int grid = (g.N + cols - 1) / cols;
  const size_t smem = 16 + (size_t)cols * g.K * 4 + (size_t)cols * DG_WARPS * MM * 4 + (size_t)DG_WARPS * MM * 4 + MM * 4 + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dec_gemv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dec_gemv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = true;
  }
  if (MM == 8) VC_LAUNCH((dec_gemv_kernel<8>), grid, DG_THREADS, smem, cs(s), a);
  else VC_LAUNCH((dec_gemv_kernel<16>), grid, DG_THREADS, smem, cs(s), a);
  return check_launch("dec_gemv_kernel");
}

int dec_attn(const DecAttn& a, int B, stream_t s) {
  if (B <= 0 || a.nh <= 0 || a.nsplit < 1 || !a.q || !a.k || !a.v || !a.part_o || !a.part_ml || !a.t_ptr) return set_error("dec_attn: bad arguments");
  if (a.dh != 64 && a.dh != 128 && a.dh != 256) return set_error("dec_attn: head dim must be 64, 128 or 256");
  const dim3 grid(a.nsplit, a.nh, B);
  if (a.dh == 64) VC_LAUNCH((dec_attn_kernel<2>), grid, DG_THREADS, 0, cs(s), a);
  else if (a.dh == 128) VC_LAUNCH((dec_attn_kernel<4>), grid, DG_THREADS, 0, cs(s), a);
  else VC_LAUNCH((dec_attn_kernel<8>), grid, DG_THREADS, 0, cs(s), a);
  return check_launch("dec_attn_kernel");
}

int dec_select(const DecSelect& a, stream_t s) {
  if (a.B <= 0 || a.H <= 0 || a.NC < 1 || a.NC > 8 || a.NPAR < 0 || a.NPAR > 8) return set_error("dec_select: unsupported head sizes");
  if (!a.y || !a.gamma || !a.beta || !a.Wc || !a.bc || !a.cmds_all || !a.params_all || !a.t_ptr || !a.done_ctr) return set_error("dec_select: null argument");
  VC_LAUNCH((dec_select_kernel), a.B, DG_THREADS, (size_t)a.H * 4, cs(s), a);
  return check_launch("dec_select_kernel");
}

}  // namespace vck
