// Decoder attention for short sequences (Tq == Tk <= 32), forward and backward, exact fp32 SIMT.
//
// Serves nn.MultiheadAttention inside torch.nn.TransformerDecoderLayer at the training shapes of the reference
// (/root/reference/model/autoregressive_transformer.py:54-62, 180-197: T = 8 .. 32 model frames, d = H / nhead = 64 .. 256,
// causal tgt_mask and banded memory_mask) -- masks are index predicates, nothing is materialised.
//
// Why a second kernel next to attention.cu: at T = 8 the generic kernel (tiled for T up to 186) spends its time in five
// block-wide phases separated by __syncthreads, each behind its own global loads: 17-18 us per launch for 36 dot products per
// head (profiles/r01j launch list), i.e. a quarter of the decoder's forward.  Here one CTA per (sample, head) loads q, k, v
// [, dO] once (all loads in flight together), then each warp owns whole query rows: scores, softmax, dropout and P~V happen
// inside the warp with no block barrier.  The backward also delivers dq/dk/dv directly as split-bf16 GEMM operands and the
// in_proj bias gradient (column sums), which removes the separate conversion pass of the generic path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"
#include "attention_small.h"

namespace vck {

namespace {

constexpr int AS_THREADS = 128;
constexpr int AS_WARPS = AS_THREADS / 32;
constexpr int AS_TMAX = 32;   // one key per lane
constexpr int AS_JB = 4;      // keys per batch of interleaved warp reductions

struct SmallP {
  const float *q, *k, *v; long long ldq, ldk, ldv;
  int T, nh, d, mask, window;
  float scale;
  Drop drop; uint32_t thresh; float dscale;
};

struct SmallBwdOut {
  float *dq, *dk, *dv; long long lddq, lddk, lddv;                      // fp32 outputs (optional)
  __nv_bfloat16 *dq_hi, *dq_lo, *dk_hi, *dk_lo, *dv_hi, *dv_lo; long long ld_split;  // split outputs (optional)
  float *dbq, *dbk, *dbv;                                               // column sums, accumulated (optional)
};

__device__ __forceinline__ void key_range(const SmallP& p, int i, int& jlo, int& jhi) {
  jlo = 0; jhi = p.T - 1;
  if (p.mask != VC_MASK_NONE) jhi = i;
  if (p.mask == VC_MASK_WINDOW) jlo = max(0, i - p.window + 1);
}
__device__ __forceinline__ void query_range(const SmallP& p, int j, int& ilo, int& ihi) {
  ilo = 0; ihi = p.T - 1;
  if (p.mask != VC_MASK_NONE) ilo = j;
  if (p.mask == VC_MASK_WINDOW) ihi = min(p.T - 1, j + p.window - 1);
}
__device__ __forceinline__ float keep_factor(const SmallP& p, const DropKey& key, unsigned long long idx) {
  if (p.drop.p <= 0.f) return 1.0f;
  return (dropout_element(key, idx) >= p.thresh) ? p.dscale : 0.0f;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int t = 0; t < N; ++t) v[t] += __shfl_xor_sync(0xffffffffu, v[t], off);
  }
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, const float4& a) {
  const float v[4] = {a.x, a.y, a.z, a.w};
  uint2 h, l;
  split4(v, h, l);
  *reinterpret_cast<uint2*>(hi + off) = h;
  if (lo) *reinterpret_cast<uint2*>(lo + off) = l;
}

// NV = ceil(d / 128): lane owns columns lane*4 + 128*u .. +3 (u < NV) of a head
template <int NV>
__global__ void __launch_bounds__(AS_THREADS)
attn_small_fwd_kernel(const SmallP p, __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, long long ldo,
                      float* __restrict__ lse_out) {
  pdl_grid_sync();
  extern __shared__ float4 as_smem4[];
  float* smem = reinterpret_cast<float*>(as_smem4);
  const int T = p.T, d = p.d;
  float* Qs = smem;
  float* Ks = Qs + T * d;
  float* Vs = Ks + T * d;
  float* Ss = Vs + T * d;  // [T][32]
  const int h = blockIdx.x % p.nh, b = blockIdx.x / p.nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = d >> 2;
  for (int idx = threadIdx.x; idx < T * d4; idx += AS_THREADS) {
    const int r = idx / d4, c = (idx - r * d4) * 4;
    const long long row = (long long)b * T + r;
    const float4 q4 = *reinterpret_cast<const float4*>(p.q + row * p.ldq + (long long)h * d + c);
    const float4 k4 = *reinterpret_cast<const float4*>(p.k + row * p.ldk + (long long)h * d + c);
    const float4 v4 = *reinterpret_cast<const float4*>(p.v + row * p.ldv + (long long)h * d + c);
    *reinterpret_cast<float4*>(Qs + r * d + c) = q4;
    *reinterpret_cast<float4*>(Ks + r * d + c) = k4;
    *reinterpret_cast<float4*>(Vs + r * d + c) = v4;
  }
  const DropKey seed = drop_key_of(p.drop);
  __syncthreads();

  for (int i = warp; i < T; i += AS_WARPS) {
    int jlo, jhi;
    key_range(p, i, jlo, jhi);
    float4 qf[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane * 4 + 128 * u;
      qf[u] = c < d ? *reinterpret_cast<const float4*>(Qs + i * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float sv = -INFINITY;  // lane j holds the score of key j
    for (int j0 = jlo; j0 <= jhi; j0 += AS_JB) {
      float part[AS_JB];
#pragma unroll
      for (int t = 0; t < AS_JB; ++t) {
        part[t] = 0.f;
        const int j = j0 + t;
        if (j <= jhi) {
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            const int c = lane * 4 + 128 * u;
            if (c < d) part[t] += dot4(qf[u], *reinterpret_cast<const float4*>(Ks + j * d + c));
          }
        }
      }
      warp_sum_n(part);
#pragma unroll
      for (int t = 0; t < AS_JB; ++t)
        if (j0 + t <= jhi && lane == j0 + t) sv = part[t] * p.scale;
    }
    const float m = warp_max(sv);
    const float e = sv == -INFINITY ? 0.f : __expf(sv - m);
    const float lse = m + __logf(warp_sum(e));
    float pj = sv == -INFINITY ? 0.f : __expf(sv - lse);
    const unsigned long long base = (((unsigned long long)b * p.nh + h) * T + i) * (unsigned long long)T;
    if (pj != 0.f) pj *= keep_factor(p, seed, base + lane);
    Ss[i * 32 + lane] = pj;
    if (lane == 0) lse_out[((long long)b * p.nh + h) * T + i] = lse;
    __syncwarp();
    float4 acc[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = jlo; j <= jhi; ++j) {
      const float pr = Ss[i * 32 + j];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < d) fma4(acc[u], pr, *reinterpret_cast<const float4*>(Vs + j * d + c));
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane * 4 + 128 * u;
      if (c < d) store_split4(o_hi, o_lo, ((long long)b * T + i) * ldo + (long long)h * d + c, acc[u]);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(AS_THREADS)
attn_small_bwd_kernel(const SmallP p, const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo,
                      long long ldo, const float* __restrict__ lse, const float* __restrict__ dout, long long lddo,
                      const SmallBwdOut out) {
  pdl_grid_sync();
  extern __shared__ float4 as_smem4[];
  float* smem = reinterpret_cast<float*>(as_smem4);
  const int T = p.T, d = p.d;
  float* Qs = smem;
  float* Ks = Qs + T * d;    // reused for dK rows after phase A
  float* Vs = Ks + T * d;    // reused for dV rows after phase A
  float* dOs = Vs + T * d;
  float* dQs = dOs + T * d;
  float* Ps = dQs + T * d;   // [T][32]  p~ (dropped probabilities)
  float* dSs = Ps + T * 32;  // [T][32]  scale * dS
  const int h = blockIdx.x % p.nh, b = blockIdx.x / p.nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = d >> 2;
  for (int idx = threadIdx.x; idx < T * d4; idx += AS_THREADS) {
    const int r = idx / d4, c = (idx - r * d4) * 4;
    const long long row = (long long)b * T + r;
    const float4 q4 = *reinterpret_cast<const float4*>(p.q + row * p.ldq + (long long)h * d + c);
    const float4 k4 = *reinterpret_cast<const float4*>(p.k + row * p.ldk + (long long)h * d + c);
    const float4 v4 = *reinterpret_cast<const float4*>(p.v + row * p.ldv + (long long)h * d + c);
    const float4 g4 = *reinterpret_cast<const float4*>(dout + row * lddo + (long long)h * d + c);
    *reinterpret_cast<float4*>(Qs + r * d + c) = q4;
    *reinterpret_cast<float4*>(Ks + r * d + c) = k4;
    *reinterpret_cast<float4*>(Vs + r * d + c) = v4;
    *reinterpret_cast<float4*>(dOs + r * d + c) = g4;
  }
  const DropKey seed = drop_key_of(p.drop);
  __syncthreads();

  // ---- phase A (warp <-> query row i): P~, scale*dS and dQ_i = sum_j dS_ij K_j
  for (int i = warp; i < T; i += AS_WARPS) {
    int jlo, jhi;
    key_range(p, i, jlo, jhi);
    float4 qf[NV], gf[NV];
    float dl = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane * 4 + 128 * u;
      qf[u] = gf[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        qf[u] = *reinterpret_cast<const float4*>(Qs + i * d + c);
        gf[u] = *reinterpret_cast<const float4*>(dOs + i * d + c);
        const long long oo = ((long long)b * T + i) * ldo + (long long)h * d + c;
        const uint2 oh = *reinterpret_cast<const uint2*>(o_hi + oo);
        float4 ov = make_float4(__uint_as_float(oh.x << 16), __uint_as_float(oh.x & 0xffff0000u), __uint_as_float(oh.y << 16),
                                __uint_as_float(oh.y & 0xffff0000u));
        if (o_lo) {
          const uint2 ol = *reinterpret_cast<const uint2*>(o_lo + oo);
          ov.x += __uint_as_float(ol.x << 16); ov.y += __uint_as_float(ol.x & 0xffff0000u);
          ov.z += __uint_as_float(ol.y << 16); ov.w += __uint_as_float(ol.y & 0xffff0000u);
        }
        dl += dot4(gf[u], ov);
      }
    }
    const float delta = warp_sum(dl);
    const float lse_i = lse[((long long)b * p.nh + h) * T + i];
    float sv = 0.f, dpv = 0.f;
    bool mine = false;
    for (int j0 = jlo; j0 <= jhi; j0 += AS_JB) {
      float part[2 * AS_JB];
#pragma unroll
      for (int t = 0; t < AS_JB; ++t) {
        part[t] = part[AS_JB + t] = 0.f;
        const int j = j0 + t;
        if (j <= jhi) {
#pragma unroll
          for (int u = 0; u < NV; ++u) {
            const int c = lane * 4 + 128 * u;
            if (c < d) {
              part[t] += dot4(qf[u], *reinterpret_cast<const float4*>(Ks + j * d + c));
              part[AS_JB + t] += dot4(gf[u], *reinterpret_cast<const float4*>(Vs + j * d + c));
            }
          }
        }
      }
      warp_sum_n(part);
#pragma unroll
      for (int t = 0; t < AS_JB; ++t)
        if (j0 + t <= jhi && lane == j0 + t) { sv = part[t]; dpv = part[AS_JB + t]; mine = true; }
    }
    float pt = 0.f, ds = 0.f;
    if (mine) {
      const float pr = __expf(sv * p.scale - lse_i);
      const unsigned long long base = (((unsigned long long)b * p.nh + h) * T + i) * (unsigned long long)T;
      const float m = keep_factor(p, seed, base + lane);
      pt = pr * m;
      ds = pr * (dpv * m - delta) * p.scale;
    }
    Ps[i * 32 + lane] = pt;
    dSs[i * 32 + lane] = ds;
    __syncwarp();
    float4 acc[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = jlo; j <= jhi; ++j) {
      const float w = dSs[i * 32 + j];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < d) fma4(acc[u], w, *reinterpret_cast<const float4*>(Ks + j * d + c));
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane * 4 + 128 * u;
      if (c < d) {
        const long long row = (long long)b * T + i;
        if (out.dq) *reinterpret_cast<float4*>(out.dq + row * out.lddq + (long long)h * d + c) = acc[u];
        if (out.dq_hi) store_split4(out.dq_hi, out.dq_lo, row * out.ld_split + (long long)h * d + c, acc[u]);
        *reinterpret_cast<float4*>(dQs + i * d + c) = acc[u];
      }
    }
  }
  __syncthreads();  // Ps / dSs complete; K and V rows are dead from here on

  // ---- phase B (warp <-> key row j): dV_j = sum_i P~_ij dO_i, dK_j = sum_i dS_ij Q_i
  for (int j = warp; j < T; j += AS_WARPS) {
    int ilo, ihi;
    query_range(p, j, ilo, ihi);
    float4 av[NV], ak[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) av[u] = ak[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = ilo; i <= ihi; ++i) {
      const float pw = Ps[i * 32 + j], sw = dSs[i * 32 + j];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < d) {
          fma4(av[u], pw, *reinterpret_cast<const float4*>(dOs + i * d + c));
          fma4(ak[u], sw, *reinterpret_cast<const float4*>(Qs + i * d + c));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane * 4 + 128 * u;
      if (c < d) {
        const long long row = (long long)b * T + j;
        if (out.dk) *reinterpret_cast<float4*>(out.dk + row * out.lddk + (long long)h * d + c) = ak[u];
        if (out.dv) *reinterpret_cast<float4*>(out.dv + row * out.lddv + (long long)h * d + c) = av[u];
        if (out.dk_hi) store_split4(out.dk_hi, out.dk_lo, row * out.ld_split + (long long)h * d + c, ak[u]);
        if (out.dv_hi) store_split4(out.dv_hi, out.dv_lo, row * out.ld_split + (long long)h * d + c, av[u]);
        *reinterpret_cast<float4*>(Ks + j * d + c) = ak[u];
        *reinterpret_cast<float4*>(Vs + j * d + c) = av[u];
      }
    }
  }
  if (out.dbq == nullptr && out.dbk == nullptr && out.dbv == nullptr) return;  // uniform
  __syncthreads();
  // ---- bias gradients: column sums of dq / dk / dv over this sample's rows
  for (int c = threadIdx.x; c < d; c += AS_THREADS) {
    float sq = 0.f, sk = 0.f, sv = 0.f;
    for (int r = 0; r < T; ++r) {
      sq += dQs[r * d + c];
      sk += Ks[r * d + c];
      sv += Vs[r * d + c];
    }
    if (out.dbq) atomicAdd(out.dbq + (long long)h * d + c, sq);
    if (out.dbk) atomicAdd(out.dbk + (long long)h * d + c, sk);
    if (out.dbv) atomicAdd(out.dbv + (long long)h * d + c, sv);
  }
}

SmallP make_small(const AttnDesc& a) {
  SmallP p;
  p.q = a.q; p.k = a.k; p.v = a.v; p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv;
  p.T = a.Tq; p.nh = a.nh; p.d = a.d; p.mask = a.mask; p.window = a.window;
  p.scale = a.scale;
  p.drop = a.drop;
  p.thresh = dropout_threshold(a.drop.p);
  p.dscale = a.drop.p > 0.f ? 1.0f / (1.0f - a.drop.p) : 1.0f;
  return p;
}

template <class K>
int ensure_smem(K kernel, size_t smem, size_t& configured) {
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = 200 * 1024;
  }
  return 0;
}

}  // namespace

static int g_small_enabled = -1;  // -1: not decided yet (VC_ATTN_SMALL=0 in the environment disables)
void attention_small_enable(int on) { g_small_enabled = on != 0 ? 1 : 0; }
static bool small_enabled() {
  if (g_small_enabled < 0) {
    const char* e = getenv("VC_ATTN_SMALL");
    g_small_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_small_enabled != 0;
}

bool attention_small_eligible(const AttnDesc& a) {
  return small_enabled() && a.q_hi == nullptr && a.q != nullptr && a.Tq == a.Tk && a.Tq >= 1 && a.Tq <= AS_TMAX && a.d % 4 == 0 &&
         a.d >= 4 && a.d <= 256 && a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 &&
         (a.mask == VC_MASK_NONE || a.mask == VC_MASK_CAUSAL || (a.mask == VC_MASK_WINDOW && a.window >= 1));
}

int attention_small_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s) {
  if (ldo % 4 != 0) return set_error("attention_small_fwd: ldo must be a multiple of 4");
  if (!o_hi || !lse) return set_error("attention_small_fwd: null output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  const size_t smem = sizeof(float) * ((size_t)3 * a.Tq * a.d + (size_t)a.Tq * 32);
  static size_t cfg1 = 0, cfg2 = 0;
  if (a.d <= 128) {
    if (int rc = ensure_smem(attn_small_fwd_kernel<1>, smem, cfg1)) return rc;
    VC_LAUNCH((attn_small_fwd_kernel<1>), a.B * a.nh, AS_THREADS, smem, st, make_small(a), reinterpret_cast<__nv_bfloat16*>(o_hi),
              reinterpret_cast<__nv_bfloat16*>(o_lo), (long long)ldo, lse);
  } else {
    if (int rc = ensure_smem(attn_small_fwd_kernel<2>, smem, cfg2)) return rc;
    VC_LAUNCH((attn_small_fwd_kernel<2>), a.B * a.nh, AS_THREADS, smem, st, make_small(a), reinterpret_cast<__nv_bfloat16*>(o_hi),
              reinterpret_cast<__nv_bfloat16*>(o_lo), (long long)ldo, lse);
  }
  return check_launch("attn_small_fwd_kernel");
}

int attention_small_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
                        int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, bf16_t* dq_hi,
                        bf16_t* dq_lo, bf16_t* dk_hi, bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, float* dbq,
                        float* dbk, float* dbv, stream_t s) {
  if (ldo % 4 != 0 || lddo % 4 != 0) return set_error("attention_small_bwd: ldo/lddo must be multiples of 4");
  if (!o_hi || !lse || !dout) return set_error("attention_small_bwd: null input");
  if ((dq && lddq % 4 != 0) || (dk && lddk % 4 != 0) || (dv && lddv % 4 != 0) || (dq_hi && ld_split % 4 != 0))
    return set_error("attention_small_bwd: output strides must be multiples of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  SmallBwdOut out;
  out.dq = dq; out.dk = dk; out.dv = dv; out.lddq = lddq; out.lddk = lddk; out.lddv = lddv;
  out.dq_hi = reinterpret_cast<__nv_bfloat16*>(dq_hi); out.dq_lo = reinterpret_cast<__nv_bfloat16*>(dq_lo);
  out.dk_hi = reinterpret_cast<__nv_bfloat16*>(dk_hi); out.dk_lo = reinterpret_cast<__nv_bfloat16*>(dk_lo);
  out.dv_hi = reinterpret_cast<__nv_bfloat16*>(dv_hi); out.dv_lo = reinterpret_cast<__nv_bfloat16*>(dv_lo);
  out.ld_split = ld_split;
  out.dbq = dbq; out.dbk = dbk; out.dbv = dbv;
  const size_t smem = sizeof(float) * ((size_t)5 * a.Tq * a.d + (size_t)2 * a.Tq * 32);
  static size_t cfg1 = 0, cfg2 = 0;
  if (a.d <= 128) {
    if (int rc = ensure_smem(attn_small_bwd_kernel<1>, smem, cfg1)) return rc;
    VC_LAUNCH((attn_small_bwd_kernel<1>), a.B * a.nh, AS_THREADS, smem, st, make_small(a), reinterpret_cast<const __nv_bfloat16*>(o_hi),
              reinterpret_cast<const __nv_bfloat16*>(o_lo), (long long)ldo, lse, dout, (long long)lddo, out);
  } else {
    if (int rc = ensure_smem(attn_small_bwd_kernel<2>, smem, cfg2)) return rc;
    VC_LAUNCH((attn_small_bwd_kernel<2>), a.B * a.nh, AS_THREADS, smem, st, make_small(a), reinterpret_cast<const __nv_bfloat16*>(o_hi),
              reinterpret_cast<const __nv_bfloat16*>(o_lo), (long long)ldo, lse, dout, (long long)lddo, out);
  }
  return check_launch("attn_small_bwd_kernel");
}

}  // namespace vck
