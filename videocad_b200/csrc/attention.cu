// Multi-head attention core, forward and backward, fp32 SIMT (exact fp32 arithmetic, probabilities never leave the
// SM).  Serves both the ViT blocks (vit_pytorch Attention: n = 50 tokens, 16 heads x 64, no mask) and the decoder
// (nn.MultiheadAttention inside TransformerDecoderLayer: T <= 186, d = H / nhead, causal tgt_mask and banded
// memory_mask built at /root/reference/model/autoregressive_transformer.py:180-188 -- here the masks are index
// predicates, nothing is materialised).
//
//   forward : one CTA per (batch, head, query block): S = scale q k^T (+mask) in smem, row softmax, dropout,
//             O = P~ V streamed over key tiles; writes O as split-bf16 (operand of the out-proj GEMM) and the
//             row log-sum-exp for the backward.
//   backward: one CTA per (batch, head): recomputes P from (q, k, lse) per (key tile, query block), accumulates
//             dK/dV in shared memory and dQ through global memory (single owner, no atomics).
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"
#include "attention_vit.h"
#include "attention_small.h"

namespace vck {

namespace {

constexpr int ATT_THREADS = 128;

template <int DMAX> struct AttCfg;
template <> struct AttCfg<64>  { static constexpr int QB_F = 64, KT_F = 64, QB_B = 32, KT_B = 64; };
template <> struct AttCfg<128> { static constexpr int QB_F = 32, KT_F = 32, QB_B = 16, KT_B = 32; };
template <> struct AttCfg<256> { static constexpr int QB_F = 16, KT_F = 16, QB_B = 16, KT_B = 16; };

struct AttnP {
  const float *q, *k, *v; long long ldq, ldk, ldv;
  long long bsq, bsk, bsv;  // batch strides (elements)
  int B, Tq, Tk, nh, d, mask, window;
  float scale;
  Drop drop; uint32_t thresh; float dscale;
};

__device__ __forceinline__ bool masked(int mask, int window, int i, int j) {
  if (mask == VC_MASK_CAUSAL) return j > i;
  if (mask == VC_MASK_WINDOW) return (j > i) || (j <= i - window);
  return false;
}
// true if key tile [k0, k0+nk) is entirely masked for query rows [q0, q0+nq)
__device__ __forceinline__ bool tile_masked(int mask, int window, int q0, int nq, int k0, int nk) {
  if (mask == VC_MASK_NONE) return false;
  if (k0 > q0 + nq - 1) return true;
  if (mask == VC_MASK_WINDOW && (k0 + nk - 1) <= q0 - window) return true;
  return false;
}
__device__ __forceinline__ float drop_factor(const AttnP& p, const DropKey& key, unsigned long long idx) {
  if (p.drop.p <= 0.f) return 1.0f;
  return (dropout_element(key, idx) >= p.thresh) ? p.dscale : 0.0f;
}

// cooperative load of `rows` rows x d floats (global row stride ld) into smem with row stride LDS
template <int LDS>
__device__ __forceinline__ void load_rows(float* dst, const float* src, long long ld, int rows, int d) {
  const int d4 = d >> 2;
  for (int idx = threadIdx.x; idx < rows * d4; idx += ATT_THREADS) {
    const int r = idx / d4, c = (idx % d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * ld + c);
    float* o = dst + r * LDS + c;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
}

template <int DMAX>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnP p, __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, long long ldo,
                float* __restrict__ lse_out) {
  pdl_grid_sync();
  constexpr int QB = AttCfg<DMAX>::QB_F, KT = AttCfg<DMAX>::KT_F, LDS = DMAX + 1;
  constexpr int CPT = DMAX > ATT_THREADS ? DMAX / ATT_THREADS : 1;  // columns per thread
  constexpr int RG = DMAX < ATT_THREADS ? ATT_THREADS / DMAX : 1;   // row groups
  constexpr int RPT = QB / RG;                                      // rows per thread
  extern __shared__ float smem[];
  float* Qs = smem;              // [QB][LDS]
  float* KVs = Qs + QB * LDS;    // [KT][LDS]
  float* Ss = KVs + KT * LDS;    // [QB][Tk]

  const int qblocks = (p.Tq + QB - 1) / QB;
  const int qb = blockIdx.x % qblocks;
  const int bh = blockIdx.x / qblocks;
  const int h = bh % p.nh, b = bh / p.nh;
  const int q0 = qb * QB, nq = min(QB, p.Tq - q0);
  const int d = p.d, Tk = p.Tk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DropKey dkey = drop_key_of(p.drop);

  load_rows<LDS>(Qs, p.q + (long long)b * p.bsq + (long long)q0 * p.ldq + (long long)h * d, p.ldq, nq, d);

  // ---- scores
  for (int k0 = 0; k0 < Tk; k0 += KT) {
    const int nk = min(KT, Tk - k0);
    const bool skip = tile_masked(p.mask, p.window, q0, nq, k0, nk);
    __syncthreads();
    if (!skip) load_rows<LDS>(KVs, p.k + (long long)b * p.bsk + (long long)k0 * p.ldk + (long long)h * d, p.ldk, nk, d);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nq * KT; idx += ATT_THREADS) {
      const int i = idx / KT, j = idx % KT;
      if (j >= nk) continue;
      float s = -INFINITY;
      if (!skip && !masked(p.mask, p.window, q0 + i, k0 + j)) {
        const float* qr = Qs + i * LDS;
        const float* kr = KVs + j * LDS;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < d; ++c) acc = fmaf(qr[c], kr[c], acc);
        s = acc * p.scale;
      }
      Ss[i * Tk + k0 + j] = s;
    }
  }
  __syncthreads();

  // ---- row softmax (+ dropout), one warp per row
  for (int i = warp; i < nq; i += ATT_THREADS / 32) {
    float* row = Ss + i * Tk;
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) sum += __expf(row[j] - m);
    sum = warp_sum(sum);
    const float lse = m + __logf(sum);
    const unsigned long long base = (((unsigned long long)b * p.nh + h) * p.Tq + (q0 + i)) * (unsigned long long)Tk;
    for (int j = lane; j < Tk; j += 32) {
      const float pj = __expf(row[j] - lse);
      row[j] = pj * drop_factor(p, dkey, base + j);
    }
    if (lane == 0) lse_out[((long long)b * p.nh + h) * p.Tq + q0 + i] = lse;
  }

  // ---- O = P~ V
  const int cg = threadIdx.x % (DMAX < ATT_THREADS ? DMAX : ATT_THREADS);
  const int rg = threadIdx.x / (DMAX < ATT_THREADS ? DMAX : ATT_THREADS);
  float acc[RPT][CPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < Tk; k0 += KT) {
    const int nk = min(KT, Tk - k0);
    if (tile_masked(p.mask, p.window, q0, nq, k0, nk)) continue;  // uniform across the CTA
    __syncthreads();
    load_rows<LDS>(KVs, p.v + (long long)b * p.bsv + (long long)k0 * p.ldv + (long long)h * d, p.ldv, nk, d);
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const int c = cg + cc * ATT_THREADS;
      if (c < d) {
        for (int j = 0; j < nk; ++j) {
          const float vv = KVs[j * LDS + c];
#pragma unroll
          for (int r = 0; r < RPT; ++r) acc[r][cc] = fmaf(Ss[(rg * RPT + r) * Tk + k0 + j], vv, acc[r][cc]);
        }
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const int c = cg + cc * ATT_THREADS;
    if (c < d) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int i = rg * RPT + r;
        if (i < nq) {
          __nv_bfloat16 hi, lo;
          split_bf16(acc[r][cc], hi, lo);
          const long long off = ((long long)b * p.Tq + q0 + i) * ldo + (long long)h * d + c;
          o_hi[off] = hi;
          if (o_lo) o_lo[off] = lo;
        }
      }
    }
  }
}

template <int DMAX>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const AttnP p, const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo,
                long long ldo, const float* __restrict__ lse, const float* __restrict__ dout, long long lddo,
                float* __restrict__ dq, long long lddq, float* __restrict__ dk, long long lddk, float* __restrict__ dv,
                long long lddv) {
  pdl_grid_sync();
  constexpr int QB = AttCfg<DMAX>::QB_B, KT = AttCfg<DMAX>::KT_B, LDS = DMAX + 1;
  constexpr int CPT = DMAX > ATT_THREADS ? DMAX / ATT_THREADS : 1;
  constexpr int CW = DMAX < ATT_THREADS ? DMAX : ATT_THREADS;  // threads along columns
  constexpr int RG = ATT_THREADS / CW;
  extern __shared__ float smem[];
  float* Ks = smem;                // [KT][LDS]
  float* Vs = Ks + KT * LDS;
  float* dKs = Vs + KT * LDS;
  float* dVs = dKs + KT * LDS;
  float* Qs = dVs + KT * LDS;      // [QB][LDS]
  float* dOs = Qs + QB * LDS;
  float* Ps = dOs + QB * LDS;      // [QB][KT]   p~ (dropped probabilities)
  float* dSs = Ps + QB * KT;       // [QB][KT]   scale * dS
  float* delta = dSs + QB * KT;    // [Tq]

  const int h = blockIdx.x % p.nh, b = blockIdx.x / p.nh;
  const int d = p.d, Tk = p.Tk, Tq = p.Tq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = threadIdx.x % CW, rg = threadIdx.x / CW;
  const DropKey dkey = drop_key_of(p.drop);

  // ---- prologue: delta_i = dO_i . O_i ; dq slice := 0
  for (int i = warp; i < Tq; i += ATT_THREADS / 32) {
    const long long ro = ((long long)b * Tq + i);
    float s = 0.f;
    for (int c = lane; c < d; c += 32) {
      const long long oo = ro * ldo + (long long)h * d + c;
      const float ov = __bfloat162float(o_hi[oo]) + (o_lo ? __bfloat162float(o_lo[oo]) : 0.f);
      s += dout[ro * lddo + (long long)h * d + c] * ov;
    }
    s = warp_sum(s);
    if (lane == 0) delta[i] = s;
    for (int c = lane; c < d; c += 32) dq[ro * lddq + (long long)h * d + c] = 0.f;
  }
  __syncthreads();

  for (int k0 = 0; k0 < Tk; k0 += KT) {
    const int nk = min(KT, Tk - k0);
    __syncthreads();
    load_rows<LDS>(Ks, p.k + ((long long)b * Tk + k0) * p.ldk + (long long)h * d, p.ldk, nk, d);
    load_rows<LDS>(Vs, p.v + ((long long)b * Tk + k0) * p.ldv + (long long)h * d, p.ldv, nk, d);
    for (int idx = threadIdx.x; idx < KT * LDS; idx += ATT_THREADS) {
      dKs[idx] = 0.f;
      dVs[idx] = 0.f;
    }
    for (int q0 = 0; q0 < Tq; q0 += QB) {
      const int nq = min(QB, Tq - q0);
      if (tile_masked(p.mask, p.window, q0, nq, k0, nk)) continue;  // uniform
      __syncthreads();
      load_rows<LDS>(Qs, p.q + ((long long)b * Tq + q0) * p.ldq + (long long)h * d, p.ldq, nq, d);
      load_rows<LDS>(dOs, dout + ((long long)b * Tq + q0) * lddo + (long long)h * d, lddo, nq, d);
      __syncthreads();
      // P~ and scale*dS for this (query block, key tile)
      for (int idx = threadIdx.x; idx < QB * KT; idx += ATT_THREADS) {
        const int i = idx / KT, j = idx % KT;
        float pt = 0.f, ds = 0.f;
        if (i < nq && j < nk && !masked(p.mask, p.window, q0 + i, k0 + j)) {
          const float* qr = Qs + i * LDS;
          const float* kr = Ks + j * LDS;
          const float* dor = dOs + i * LDS;
          const float* vr = Vs + j * LDS;
          float s = 0.f, dp = 0.f;
#pragma unroll 8
          for (int c = 0; c < d; ++c) {
            s = fmaf(qr[c], kr[c], s);
            dp = fmaf(dor[c], vr[c], dp);
          }
          const long long gi = ((long long)b * p.nh + h) * Tq + q0 + i;
          const float pr = __expf(s * p.scale - lse[gi]);
          const float m = drop_factor(p, dkey, (unsigned long long)gi * (unsigned long long)Tk + (k0 + j));
          pt = pr * m;
          ds = pr * (dp * m - delta[q0 + i]) * p.scale;
        }
        Ps[idx] = pt;
        dSs[idx] = ds;
      }
      __syncthreads();
      // dV += P~^T dO ; dK += dS^T Q   (thread <-> column(s) c, key rows j = rg, rg+RG, ...)
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const int c = cg + cc * ATT_THREADS;
        if (c < d) {
          for (int j = rg; j < nk; j += RG) {
            float av = 0.f, ak = 0.f;
            for (int i = 0; i < nq; ++i) {
              av = fmaf(Ps[i * KT + j], dOs[i * LDS + c], av);
              ak = fmaf(dSs[i * KT + j], Qs[i * LDS + c], ak);
            }
            dVs[j * LDS + c] += av;
            dKs[j * LDS + c] += ak;
          }
          // dQ += dS K  (query rows i = rg, rg+RG, ...), single owner of this dq slice -> plain read-modify-write
          for (int i = rg; i < nq; i += RG) {
            float aq = 0.f;
            for (int j = 0; j < nk; ++j) aq = fmaf(dSs[i * KT + j], Ks[j * LDS + c], aq);
            float* dst = dq + ((long long)b * Tq + q0 + i) * lddq + (long long)h * d + c;
            *dst += aq;
          }
        }
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nk * d; idx += ATT_THREADS) {
      const int j = idx / d, c = idx % d;
      dk[((long long)b * Tk + k0 + j) * lddk + (long long)h * d + c] = dKs[j * LDS + c];
      dv[((long long)b * Tk + k0 + j) * lddv + (long long)h * d + c] = dVs[j * LDS + c];
    }
  }
}

inline AttnP make_params(const AttnDesc& a) {
  AttnP p;
  p.q = a.q; p.k = a.k; p.v = a.v; p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv;
  p.bsq = a.bsq ? a.bsq : (long long)a.Tq * a.ldq;
  p.bsk = a.bsk ? a.bsk : (long long)a.Tk * a.ldk;
  p.bsv = a.bsv ? a.bsv : (long long)a.Tk * a.ldv;
  p.B = a.B; p.Tq = a.Tq; p.Tk = a.Tk; p.nh = a.nh; p.d = a.d; p.mask = a.mask; p.window = a.window;
  p.scale = a.scale;
  p.drop = a.drop;
  p.thresh = dropout_threshold(a.drop.p);
  p.dscale = a.drop.p > 0.f ? 1.0f / (1.0f - a.drop.p) : 1.0f;
  return p;
}

int validate(const AttnDesc& a, const char* who) {
  if (a.q_hi != nullptr && !(vit_attention_eligible(a)))
    return set_error("attention: split-bf16 inputs are supported by the tensor-core ViT kernel only (n <= 64, d = 64, no mask)");
  if (a.d % 4 != 0 || a.d > 256 || a.d <= 0) return set_error("attention: head dim must be a multiple of 4 and <= 256");
  if (a.ldq % 4 != 0 || a.ldk % 4 != 0 || a.ldv % 4 != 0) return set_error("attention: q/k/v strides must be multiples of 4");
  if (a.mask != VC_MASK_NONE && a.Tq != a.Tk) return set_error("attention: masked attention needs Tq == Tk");
  if (a.mask == VC_MASK_WINDOW && a.window < 1) return set_error("attention: window must be >= 1");
  (void)who;
  return 0;
}

template <int DMAX>
int launch_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, cudaStream_t st) {
  constexpr int QB = AttCfg<DMAX>::QB_F, KT = AttCfg<DMAX>::KT_F, LDS = DMAX + 1;
  const size_t smem = sizeof(float) * ((size_t)QB * LDS + (size_t)KT * LDS + (size_t)QB * a.Tk);
  if (smem > 200 * 1024) return set_error("attention_fwd: sequence too long for the shared-memory score buffer");
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = 200 * 1024;
  }
  const int qblocks = (a.Tq + QB - 1) / QB;
  VC_LAUNCH((attn_fwd_kernel<DMAX>), a.B * a.nh * qblocks, ATT_THREADS, smem, st, make_params(a), reinterpret_cast<__nv_bfloat16*>(o_hi),
                                                                        reinterpret_cast<__nv_bfloat16*>(o_lo), ldo, lse);
  return check_launch("attn_fwd_kernel");
}

template <int DMAX>
int launch_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
               int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, cudaStream_t st) {
  constexpr int QB = AttCfg<DMAX>::QB_B, KT = AttCfg<DMAX>::KT_B, LDS = DMAX + 1;
  const size_t smem = sizeof(float) * ((size_t)4 * KT * LDS + (size_t)2 * QB * LDS + (size_t)2 * QB * KT + (size_t)a.Tq);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = true;
  }
  if (smem > 200 * 1024) return set_error("attention_bwd: sequence too long");
  VC_LAUNCH((attn_bwd_kernel<DMAX>), a.B * a.nh, ATT_THREADS, smem, st, 
      make_params(a), reinterpret_cast<const __nv_bfloat16*>(o_hi), reinterpret_cast<const __nv_bfloat16*>(o_lo), ldo, lse,
      dout, lddo, dq, lddq, dk, lddk, dv, lddv);
  return check_launch("attn_bwd_kernel");
}

}  // namespace


static inline bool dense_batches(const AttnDesc& a) { return a.bsq == 0 && a.bsk == 0 && a.bsv == 0; }

int attention_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s) {
  if (int rc = validate(a, "attention_fwd")) return rc;
  if (a.B <= 0 || a.Tq <= 0) return 0;
  if (!dense_batches(a) && (a.q_hi != nullptr || a.bsq % 4 != 0 || a.bsk % 4 != 0 || a.bsv % 4 != 0))
    return set_error("attention_fwd: batch strides need fp32 inputs and multiples of 4");
  if (dense_batches(a) && vit_attention_eligible(a)) return vit_attention_fwd(a, o_hi, o_lo, ldo, lse, s);
  if (dense_batches(a) && attention_small_eligible(a) && ldo % 4 == 0) return attention_small_fwd(a, o_hi, o_lo, ldo, lse, s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (a.d <= 64) return launch_fwd<64>(a, o_hi, o_lo, ldo, lse, st);
  if (a.d <= 128) return launch_fwd<128>(a, o_hi, o_lo, ldo, lse, st);
  return launch_fwd<256>(a, o_hi, o_lo, ldo, lse, st);
}

int attention_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                  const float* dout, int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                  int64_t lddv, stream_t s) {
  if (int rc = validate(a, "attention_bwd")) return rc;
  if (!dense_batches(a)) return set_error("attention_bwd: batch strides are supported by the forward only");
  if (lddo % 4 != 0) return set_error("attention_bwd: dout stride must be a multiple of 4");
  if (a.B <= 0 || a.Tq <= 0) return 0;
  if (vit_attention_eligible(a))
    return vit_attention_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, s);
  if (attention_small_eligible(a) && ldo % 4 == 0 && lddq % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0)
    return attention_small_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, 0, nullptr, nullptr, nullptr, s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (a.d <= 64) return launch_bwd<64>(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, st);
  if (a.d <= 128) return launch_bwd<128>(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, st);
  return launch_bwd<256>(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, st);
}

int attention_bwd_split(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                        const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, float* scratch, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi,
                        bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, stream_t s) {
  if (int rc = validate(a, "attention_bwd_split")) return rc;
  if (!dense_batches(a)) return set_error("attention_bwd_split: batch strides are supported by the forward only");
  if (a.B <= 0 || a.Tq <= 0) return 0;
  if (vit_attention_eligible(a))
    return vit_attention_bwd_split(a, o_hi, o_lo, ldo, lse, dout, dout_hi, dout_lo, lddo, dq_hi, dq_lo, dk_hi, dk_lo, dv_hi, dv_lo,
                                   ld_split, s);
  if (!dout) return set_error("attention_bwd_split: the generic kernel needs an fp32 upstream gradient");
  if (!scratch) return set_error("attention_bwd_split: scratch required");
  const int64_t W = (int64_t)a.nh * a.d;
  const int64_t Rq = (int64_t)a.B * a.Tq, Rk = (int64_t)a.B * a.Tk;
  float* dq = scratch;
  float* dk = dq + Rq * W;
  float* dv = dk + Rk * W;
  if (int rc = attention_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, W, dk, W, dv, W, s)) return rc;
  if (int rc = split_f32(dq, W, Rq, W, dq_hi, dq_lo, ld_split, s)) return rc;
  if (int rc = split_f32(dk, W, Rk, W, dk_hi, dk_lo, ld_split, s)) return rc;
  return split_f32(dv, W, Rk, W, dv_hi, dv_lo, ld_split, s);
}

// attention_bwd_split for attention behind a packed in_proj WITH bias (nn.MultiheadAttention): also accumulates the bias
// gradient, i.e. the column sums of dq / dk / dv, into dbq / dbk / dbv ([nh*d] each, caller-zeroed, any may be null).
int attention_bwd_split_bias(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                             const float* dout, int64_t lddo, float* scratch, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi,
                             bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, float* dbq, float* dbk, float* dbv,
                             stream_t s) {
  if (int rc = validate(a, "attention_bwd_split_bias")) return rc;
  if (!dense_batches(a)) return set_error("attention_bwd_split_bias: batch strides are supported by the forward only");
  if (a.B <= 0 || a.Tq <= 0) return 0;
  if (!dout) return set_error("attention_bwd_split_bias: an fp32 upstream gradient is required");
  if (attention_small_eligible(a) && ldo % 4 == 0 && lddo % 4 == 0 && ld_split % 4 == 0)
    return attention_small_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, nullptr, 0, nullptr, 0, nullptr, 0, dq_hi, dq_lo, dk_hi, dk_lo, dv_hi,
                               dv_lo, ld_split, dbq, dbk, dbv, s);
  if (!scratch) return set_error("attention_bwd_split_bias: scratch required");
  const int64_t W = (int64_t)a.nh * a.d;
  const int64_t Rq = (int64_t)a.B * a.Tq, Rk = (int64_t)a.B * a.Tk;
  float* dq = scratch;
  float* dk = dq + Rq * W;
  float* dv = dk + Rk * W;
  if (int rc = attention_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, W, dk, W, dv, W, s)) return rc;
  if (int rc = act_dropout_bwd(dq, W, Rq, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dq_hi, dq_lo, ld_split, dbq, s)) return rc;
  if (int rc = act_dropout_bwd(dk, W, Rk, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dk_hi, dk_lo, ld_split, dbk, s)) return rc;
  return act_dropout_bwd(dv, W, Rk, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dv_hi, dv_lo, ld_split, dbv, s);
}

}  // namespace vck
