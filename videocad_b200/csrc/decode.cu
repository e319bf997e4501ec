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
//     ln_fwd_kernel) or tanh(embed_action(a) + E[t]) for the first layer -- so LayerNorm and the token embedding cost no
//     launches of their own.
//   * dec_attn_kernel  one query row per (sequence, head) against the cached keys/values: the keys are split over `nsplit` CTAs
//     and 8 warps each (online softmax per warp, merged in shared memory); each CTA leaves its partial (max, sum, acc) in global
//     memory and the LAST one to arrive (atomic ticket) merges them into the attention output: no second launch.
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
// L2 evict-first: the 293 MB of weight rows are read once per step and must not push the step's re-used data (cached keys / values,
// LayerNorm parameters, activations) out of the 126 MB L2
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
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
  if (warp == 0) {
    // The CTA's weight rows [n0, n0 + cols) x K are one contiguous slab of the row-major [N, K] weight.  It is fetched as one TMA
    // bulk copy PER ROW, issued by the lanes of warp 0: a single 96 KB copy is processed as one serial request stream (measured:
    // ~6 B/clk per SM, profiles/r02c_rollout_launches.csv), many 2-4 KB copies are in flight together.  Issued before the
    // dependency wait: weights are constants of the rollout.
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
      mbar_expect_tx(bar, (uint32_t)cols * (uint32_t)K * 4u);
    }
    __syncwarp();
    for (int c = lane; c < cols; c += 32) bulk_g2s(wsm + (size_t)c * K, a.W + (size_t)(n0 + c) * K, (uint32_t)K * 4u, bar);
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
  }
  if (a.x_out != nullptr && blockIdx.x == 0 && active) {  // the rows just built are the residual operand of a later kernel
#pragma unroll
    for (int m = 0; m < MM; ++m)
      if (m < M) *reinterpret_cast<float4*>(a.x_out + (size_t)m * K + k0) = x[m];
  }
  // ---------------------------------------------------------------- epilogue operands do not depend on the dot products: request them now
  const int e_m = tid / cols, e_c = tid - e_m * cols;
  const bool e_on = tid < cols * M;
  float e_bias = 0.f, e_res = 0.f;
  if (e_on) {
    if (a.bias) e_bias = __ldg(a.bias + n0 + e_c);
    if (a.residual) e_res = a.residual[(size_t)e_m * a.ld_res + n0 + e_c];
  }
  // ---------------------------------------------------------------- dot products against the weight slab
  mbar_wait(bar, 0);
  constexpr int kShift = MM == 8 ? 2 : 1;
  int c = 0;
  for (; c + 4 <= cols; c += 4) {  // four columns per iteration: four independent load / FMA / shuffle chains
    float p0[MM], p1[MM], p2[MM], p3[MM];
    if (active) {
      const float4 w0 = *reinterpret_cast<const float4*>(wsm + (size_t)(c + 0) * K + k0);
      const float4 w1 = *reinterpret_cast<const float4*>(wsm + (size_t)(c + 1) * K + k0);
      const float4 w2 = *reinterpret_cast<const float4*>(wsm + (size_t)(c + 2) * K + k0);
      const float4 w3 = *reinterpret_cast<const float4*>(wsm + (size_t)(c + 3) * K + k0);
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        p0[m] = fmaf(w0.x, x[m].x, fmaf(w0.y, x[m].y, fmaf(w0.z, x[m].z, w0.w * x[m].w)));
        p1[m] = fmaf(w1.x, x[m].x, fmaf(w1.y, x[m].y, fmaf(w1.z, x[m].z, w1.w * x[m].w)));
        p2[m] = fmaf(w2.x, x[m].x, fmaf(w2.y, x[m].y, fmaf(w2.z, x[m].z, w2.w * x[m].w)));
        p3[m] = fmaf(w3.x, x[m].x, fmaf(w3.y, x[m].y, fmaf(w3.z, x[m].z, w3.w * x[m].w)));
      }
    } else {
#pragma unroll
      for (int m = 0; m < MM; ++m) p0[m] = p1[m] = p2[m] = p3[m] = 0.f;
    }
    const float r0 = reduce_rows<MM>(p0, lane), r1 = reduce_rows<MM>(p1, lane), r2 = reduce_rows<MM>(p2, lane), r3 = reduce_rows<MM>(p3, lane);
    if ((lane & ((1 << kShift) - 1)) == 0) {
      float* dst = part + ((size_t)c * DG_WARPS + warp) * MM + (lane >> kShift);
      dst[0] = r0; dst[DG_WARPS * MM] = r1; dst[2 * DG_WARPS * MM] = r2; dst[3 * DG_WARPS * MM] = r3;
    }
  }
  for (; c < cols; ++c) {
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
    const int m = idx / cols, cc = idx - m * cols;
    const int n = n0 + cc;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < DG_WARPS; ++w) v += part[((size_t)cc * DG_WARPS + w) * MM + m];
    const bool first = idx == tid;  // operands of the first round were requested before the dot products
    if (a.bias) v += first ? e_bias : __ldg(a.bias + n);
    v = apply_act(v, a.act);
    if (a.residual) v += first ? e_res : a.residual[(size_t)m * a.ld_res + n];
    a.out[(size_t)m * a.out_row_stride + (size_t)t * a.out_t_stride + n] = v;
  }
}

// ------------------------------------------------------------------------------------------------- attention of one query row
template <int DPL>
struct VecLd;
template <>
struct VecLd<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
};
template <>
struct VecLd<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct VecLd<8> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    const float4 t = *reinterpret_cast<const float4*>(p), u = *reinterpret_cast<const float4*>(p + 4);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; v[4] = u.x; v[5] = u.y; v[6] = u.z; v[7] = u.w;
  }
};

constexpr int DA_KEYS = 4;  // keys per warp iteration: all their key / value rows are requested before the first reduction

template <int DPL>
__global__ void __launch_bounds__(DG_THREADS) dec_attn_kernel(const DecAttn a) {
  pdl_grid_sync();
  __shared__ float sm_m[DG_WARPS], sm_l[DG_WARPS];
  __shared__ float sm_acc[DG_WARPS][32 * DPL];
  __shared__ unsigned int sm_ticket;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int t = *a.t_ptr;
  const int j_lo = a.window > 0 ? max(0, t - a.window + 1) : 0;
  const int nkeys = t - j_lo + 1;
  const int d0 = h * a.dh + lane * DPL;
  float q[DPL];
  VecLd<DPL>::ld(a.q + (size_t)b * a.q_bstride + (size_t)t * a.q_tstride + d0, q);
#pragma unroll
  for (int i = 0; i < DPL; ++i) q[i] *= a.scale;
  float mrun = -INFINITY, lrun = 0.f, acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  const float* kb = a.k + (size_t)b * a.kv_bstride + d0;
  const float* vb = a.v + (size_t)b * a.kv_bstride + d0;
  const int stride = a.nsplit * DG_WARPS;
  for (int i = s * DG_WARPS + warp; i < nkeys; i += DA_KEYS * stride) {
    float kk[DA_KEYS][DPL], vv[DA_KEYS][DPL], sc[DA_KEYS];
#pragma unroll
    for (int u = 0; u < DA_KEYS; ++u) {
      const int iu = i + u * stride;
      const size_t row = (size_t)(j_lo + (iu < nkeys ? iu : i)) * a.kv_rstride;  // past the end: re-read a valid row, weight 0
      VecLd<DPL>::ld(kb + row, kk[u]);
      VecLd<DPL>::ld(vb + row, vv[u]);
    }
    float mnew = mrun;
#pragma unroll
    for (int u = 0; u < DA_KEYS; ++u) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < DPL; ++e) d = fmaf(q[e], kk[u][e], d);
      d = warp_sum(d);
      sc[u] = (i + u * stride < nkeys) ? d : -INFINITY;  // warp-uniform
      mnew = fmaxf(mnew, sc[u]);
    }
    const float corr = __expf(mrun - mnew);  // mnew is finite: key i exists
    lrun *= corr;
#pragma unroll
    for (int e = 0; e < DPL; ++e) acc[e] *= corr;
#pragma unroll
    for (int u = 0; u < DA_KEYS; ++u) {
      const float p = __expf(sc[u] - mnew);  // exp(-inf) = 0 for the keys past the end
      lrun += p;
#pragma unroll
      for (int e = 0; e < DPL; ++e) acc[e] = fmaf(p, vv[u][e], acc[e]);
    }
    mrun = mnew;
  }
  if (lane == 0) { sm_m[warp] = mrun; sm_l[warp] = lrun; }
#pragma unroll
  for (int e = 0; e < DPL; ++e) sm_acc[warp][lane * DPL + e] = acc[e];
  __syncthreads();
  // merge the 8 warps of this CTA
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < DG_WARPS; ++w) mx = fmaxf(mx, sm_m[w]);
  float o = 0.f, l = 0.f;
  if (mx != -INFINITY) {
#pragma unroll
    for (int w = 0; w < DG_WARPS; ++w) {
      if (sm_m[w] != -INFINITY) {
        const float e = __expf(sm_m[w] - mx);
        l += e * sm_l[w];
        if (tid < a.dh) o += e * sm_acc[w][tid];
      }
    }
  }
  float* outp = a.out + (size_t)b * a.ld_out + h * a.dh;
  if (a.nsplit == 1) {  // a single part: finished
    if (tid < a.dh) outp[tid] = o / l;
    return;
  }
  const size_t pbase = ((size_t)b * a.nh + h) * a.nsplit;
  if (tid < a.dh) a.part_o[(pbase + s) * a.dh + tid] = o;
  if (tid == 0) { a.part_ml[(pbase + s) * 2] = mx; a.part_ml[(pbase + s) * 2 + 1] = l; }
  // the last part of this (sequence, head) to arrive merges all of them (no second launch, no extra pass in the consumer)
  __threadfence();
  __syncthreads();
  if (tid == 0) sm_ticket = atomicAdd(a.counters + (size_t)b * a.nh + h, 1u);
  __syncthreads();
  if (sm_ticket != (unsigned int)(a.nsplit - 1)) return;
  __threadfence();
  if (tid == 0) a.counters[(size_t)b * a.nh + h] = 0u;  // ready for the next launch
  float gm = -INFINITY;
  for (int s2 = 0; s2 < a.nsplit; ++s2) gm = fmaxf(gm, __ldcg(a.part_ml + (pbase + s2) * 2));
  if (tid < a.dh) {
    float den = 0.f, num = 0.f;
    for (int s2 = 0; s2 < a.nsplit; ++s2) {
      const float ms = __ldcg(a.part_ml + (pbase + s2) * 2);
      if (ms == -INFINITY) continue;  // a part that saw no key
      const float w = __expf(ms - gm);
      den += w * __ldcg(a.part_ml + (pbase + s2) * 2 + 1);
      num += w * __ldcg(a.part_o + (pbase + s2) * a.dh + tid);
    }
    outp[tid] = num / den;
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
  if (B <= 0 || a.nh <= 0 || a.nsplit < 1 || !a.q || !a.k || !a.v || !a.out || !a.t_ptr) return set_error("dec_attn: bad arguments");
  if (a.nsplit > 1 && (!a.part_o || !a.part_ml || !a.counters)) return set_error("dec_attn: split keys need the partial buffers and counters");
  if (a.q_bstride % 4 != 0 || a.q_tstride % 4 != 0 || a.kv_bstride % 4 != 0 || a.kv_rstride % 4 != 0) return set_error("dec_attn: strides must be multiples of 4");
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
