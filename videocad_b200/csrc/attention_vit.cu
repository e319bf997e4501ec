// ViT attention core on tensor cores (warp-level mma.sync m16n8k16, split-bf16 3-pass = fp32-grade), forward and
// backward, for the shape the reference's ViT uses: n <= 64 tokens per image, 16 heads x 64, no mask
// (vit_pytorch Attention as configured at /root/reference/model/trajectory_model.py:54-67).
//
// The tile is far below the 128-row tcgen05 atom (n = 50 keys/queries per head), so this kernel uses the
// warp-level tensor path: one CTA per (image, head), 4 warps, each warp owns 16 query rows (forward, backward
// phase 1) or 16 key rows (backward phase 2).  K, V (and Q, dO in the backward) are converted to split-bf16 on the way
// into shared memory; scores/probabilities live in registers (forward) or make one trip through shared memory
// (backward, to transpose P~ and dS).  Dropout masks are regenerated from the counter-based generator (dropout_rng.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

constexpr int NT = 64;     // padded tokens
constexpr int HD = 64;     // head dim
constexpr int LDH = 72;    // smem row stride in bf16 elements (144 B: conflict-free ldmatrix)
constexpr int VA_THREADS = 128;

struct VitBwdOut {  // fp32 outputs (dq/dk/dv) or, when s_hi[0] != nullptr, split-bf16 outputs
  float* d[3]; long long ldd[3];
  __nv_bfloat16* s_hi[3]; __nv_bfloat16* s_lo[3]; long long lds;
};

struct VitAttnP {
  const float *q, *k, *v; long long ldq, ldk, ldv;
  const __nv_bfloat16 *qh, *ql, *kh, *kl, *vh, *vl;  // split-bf16 inputs (used when qh != nullptr)
  int n, nh;
  float scale;
  Drop drop; uint32_t thresh; float dscale;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// split two floats into packed (hi, lo) bf16x2 words
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) { split_pair(x, y, hi, lo); }

// dropout factors for two consecutive elements idx, idx+1: ONE hash when idx is even (both elements live in the same 32-bit
// word of the generator, see dropout_rng.h) -- always the case for an even token count such as the ViT's n = 50
__device__ __forceinline__ void drop_pair(const VitAttnP& p, const DropKey& key, unsigned long long idx, float& f0, float& f1) {
  if (p.drop.p <= 0.f) { f0 = f1 = 1.0f; return; }
  const uint32_t w = dropout_word(key, idx >> 1);
  if ((idx & 1ull) == 0ull) {
    f0 = ((w << 16) >= p.thresh) ? p.dscale : 0.0f;
    f1 = ((w & 0xffff0000u) >= p.thresh) ? p.dscale : 0.0f;
  } else {
    const uint32_t w1 = dropout_word(key, (idx >> 1) + 1ull);
    f0 = ((w & 0xffff0000u) >= p.thresh) ? p.dscale : 0.0f;
    f1 = ((w1 << 16) >= p.thresh) ? p.dscale : 0.0f;
  }
}

// cooperative: rows [0,64) x 64 fp32 columns from global (row stride ld) -> split-bf16 smem tiles; rows >= n are zero
__device__ __forceinline__ void stage_split(const float* src, long long ld, int n, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  for (int idx = threadIdx.x; idx < NT * (HD / 4); idx += VA_THREADS) {
    const int r = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n) v = *reinterpret_cast<const float4*>(src + (long long)r * ld + c);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    uint2 h, l;
    split4(vv, h, l);
    *reinterpret_cast<uint2*>(hi + r * LDH + c) = h;
    *reinterpret_cast<uint2*>(lo + r * LDH + c) = l;
  }
}

// same staging when the source already is split-bf16: asynchronous 16-byte copies (cp.async, zero-fill for the padding
// rows), no conversion and no registers; the caller waits once (cp_async_wait_all + __syncthreads) for all tiles
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stage_copy(const __nv_bfloat16* src_hi, const __nv_bfloat16* src_lo, long long ld, int n,
                                           __nv_bfloat16* hi, __nv_bfloat16* lo) {
#pragma unroll
  for (int it = 0; it < NT * (HD / 8) / VA_THREADS; ++it) {
    const int idx = threadIdx.x + it * VA_THREADS;
    const int r = idx / (HD / 8), c = (idx % (HD / 8)) * 8;
    const int ok = r < n ? 16 : 0;
    const long long off = (long long)(r < n ? r : 0) * ld + c;
    cp_async16(hi + r * LDH + c, src_hi + off, ok);
    cp_async16(lo + r * LDH + c, src_lo + off, ok);
  }
}

// A fragments of this warp's 16 rows straight from global split-bf16
__device__ __forceinline__ void load_a_global_split(const __nv_bfloat16* bh, const __nv_bfloat16* bl, long long ld, int row0, int n,
                                                    int g, int t, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = row0 + g + rr * 8, col = kk * 16 + 2 * t + half * 8;
        uint32_t h = 0u, l = 0u;
        if (row < n) {
          h = *reinterpret_cast<const uint32_t*>(bh + (long long)row * ld + col);
          l = *reinterpret_cast<const uint32_t*>(bl + (long long)row * ld + col);
        }
        ah[kk][half * 2 + rr] = h;
        al[kk][half * 2 + rr] = l;
      }
    }
  }
}

// A fragments (16 rows x 64 cols, split) of this warp's rows straight from global fp32
__device__ __forceinline__ void load_a_global(const float* base, long long ld, int row0, int n, int g, int t, uint32_t (&ah)[4][4],
                                              uint32_t (&al)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = row0 + g + rr * 8, col = kk * 16 + 2 * t + half * 8;
        float2 v = make_float2(0.f, 0.f);
        if (row < n) v = *reinterpret_cast<const float2*>(base + (long long)row * ld + col);
        split2(v.x, v.y, ah[kk][half * 2 + rr], al[kk][half * 2 + rr]);
      }
    }
  }
}

// acc[8][4] (16 x 64) += A(16 x 64, split regs) * B^T where B is a [64 x 64] split smem tile stored [n][k] (non-transposed read).
// Per k16 step the B fragments of all four column pairs are fetched first and the three passes (hi*hi, lo*hi, hi*lo) are issued
// pass by pass over the eight accumulators: consecutive mma.sync instructions never accumulate into the same registers (issuing
// the three passes of one accumulator back to back left only two independent chains in flight: the mma.sync pipe was 37 % busy,
// stall reason `wait`).  Each accumulator still receives its products in the same order, so results are bit-identical.
__device__ __forceinline__ void mma_ABt(float (&acc)[8][4], const uint32_t (&ah)[4][4], const uint32_t (&al)[4][4],
                                        const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t bh[4][4], bl[4][4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const int row = jp * 16 + (lane & 7) + (lane >> 4) * 8, col = kk * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4(bh[jp], Bh + row * LDH + col);
      ldsm_x4(bl[jp], Bl + row * LDH + col);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], ah[kk], bh[jp][0], bh[jp][1]);
      mma16816(acc[2 * jp + 1], ah[kk], bh[jp][2], bh[jp][3]);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], al[kk], bh[jp][0], bh[jp][1]);
      mma16816(acc[2 * jp + 1], al[kk], bh[jp][2], bh[jp][3]);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], ah[kk], bl[jp][0], bl[jp][1]);
      mma16816(acc[2 * jp + 1], ah[kk], bl[jp][2], bl[jp][3]);
    }
  }
}

// acc[8][4] (16 x 64) += A(16 x 64 keys, split regs) * B where B is a [64 keys x 64] split smem tile stored [k][n] (transposed read)
__device__ __forceinline__ void mma_AB(float (&acc)[8][4], const uint32_t (&ah)[4][4], const uint32_t (&al)[4][4],
                                       const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t bh[4][4], bl[4][4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = jp * 16 + (lane >> 4) * 8;
      ldsm_x4_t(bh[jp], Bh + row * LDH + col);
      ldsm_x4_t(bl[jp], Bl + row * LDH + col);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], ah[kk], bh[jp][0], bh[jp][1]);
      mma16816(acc[2 * jp + 1], ah[kk], bh[jp][2], bh[jp][3]);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], al[kk], bh[jp][0], bh[jp][1]);
      mma16816(acc[2 * jp + 1], al[kk], bh[jp][2], bh[jp][3]);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      mma16816(acc[2 * jp], ah[kk], bl[jp][0], bl[jp][1]);
      mma16816(acc[2 * jp + 1], ah[kk], bl[jp][2], bl[jp][3]);
    }
  }
}

// C-layout accumulators (16 x 64) -> split A fragments for a following MMA that contracts over those 64 columns
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    split2(c[2 * kk][0], c[2 * kk][1], ah[kk][0], al[kk][0]);
    split2(c[2 * kk][2], c[2 * kk][3], ah[kk][1], al[kk][1]);
    split2(c[2 * kk + 1][0], c[2 * kk + 1][1], ah[kk][2], al[kk][2]);
    split2(c[2 * kk + 1][2], c[2 * kk + 1][3], ah[kk][3], al[kk][3]);
  }
}

__global__ void __launch_bounds__(VA_THREADS)
vit_attn_fwd_mma_kernel(const VitAttnP p, __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, long long ldo,
                        float* __restrict__ lse_out) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Kl = Kh + NT * LDH;
  __nv_bfloat16* Vh = Kl + NT * LDH;
  __nv_bfloat16* Vl = Vh + NT * LDH;
  const int h = blockIdx.x % p.nh, b = blockIdx.x / p.nh;
  const int n = p.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long rowbase = (long long)b * n;
  const DropKey dkey = drop_key_of(p.drop);
  uint32_t qh[4][4], ql[4][4];
  if (p.qh != nullptr) {
    stage_copy(p.kh + rowbase * p.ldk + (long long)h * HD, p.kl + rowbase * p.ldk + (long long)h * HD, p.ldk, n, Kh, Kl);
    stage_copy(p.vh + rowbase * p.ldv + (long long)h * HD, p.vl + rowbase * p.ldv + (long long)h * HD, p.ldv, n, Vh, Vl);
    load_a_global_split(p.qh + rowbase * p.ldq + (long long)h * HD, p.ql + rowbase * p.ldq + (long long)h * HD, p.ldq, warp * 16, n,
                        g, t, qh, ql);
  } else {
    stage_split(p.k + rowbase * p.ldk + (long long)h * HD, p.ldk, n, Kh, Kl);
    stage_split(p.v + rowbase * p.ldv + (long long)h * HD, p.ldv, n, Vh, Vl);
    load_a_global(p.q + rowbase * p.ldq + (long long)h * HD, p.ldq, warp * 16, n, g, t, qh, ql);
  }
  cp_async_wait_all();
  __syncthreads();

  float s[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
  mma_ABt(s, qh, ql, Kh, Kl, lane);

  // softmax over the n valid keys, rows r0 = warp*16+g (regs [0],[1]) and r1 = r0+8 (regs [2],[3])
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = warp * 16 + g + rr * 8;
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * j + 2 * t + e;
        float v = s[j][rr * 2 + e] * p.scale;
        if (col >= n) v = -INFINITY;
        s[j][rr * 2 + e] = v;
        m = fmaxf(m, v);
      }
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pv = __expf(s[j][rr * 2 + e] - m);
        s[j][rr * 2 + e] = pv;
        sum += pv;
      }
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const unsigned long long base = (((unsigned long long)b * p.nh + h) * n + (unsigned long long)(row < n ? row : 0)) * (unsigned long long)n;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * j + 2 * t;
      float f0 = 1.f, f1 = 1.f;
      if (col < n) drop_pair(p, dkey, base + col, f0, f1);  // f1 is unused when col + 1 == n (its probability is 0)
      s[j][rr * 2 + 0] *= inv * f0;
      s[j][rr * 2 + 1] *= inv * f1;
    }
    if (row < n && t == 0) lse_out[((long long)b * p.nh + h) * n + row] = m + __logf(sum);
  }

  uint32_t ph[4][4], pl[4][4];
  c_to_a(s, ph, pl);
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  mma_AB(o, ph, pl, Vh, Vl, lane);

#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = warp * 16 + g + rr * 8;
    if (row < n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t hi, lo;
        split2(o[j][rr * 2], o[j][rr * 2 + 1], hi, lo);
        const long long off = (rowbase + row) * ldo + (long long)h * HD + 8 * j + 2 * t;
        *reinterpret_cast<uint32_t*>(o_hi + off) = hi;
        if (o_lo) *reinterpret_cast<uint32_t*>(o_lo + off) = lo;
      }
    }
  }
}

__device__ __forceinline__ void store_pair(const VitBwdOut& out, int which, long long row, int col, float x, float y) {
  // explicit selects instead of dynamic indexing into the kernel-parameter arrays (that forced a local-memory copy)
  if (out.s_hi[0] != nullptr) {
    __nv_bfloat16* ph = which == 0 ? out.s_hi[0] : (which == 1 ? out.s_hi[1] : out.s_hi[2]);
    __nv_bfloat16* pl = which == 0 ? out.s_lo[0] : (which == 1 ? out.s_lo[1] : out.s_lo[2]);
    uint32_t hi, lo;
    split2(x, y, hi, lo);
    *reinterpret_cast<uint32_t*>(ph + row * out.lds + col) = hi;
    *reinterpret_cast<uint32_t*>(pl + row * out.lds + col) = lo;
  } else {
    float* pd = which == 0 ? out.d[0] : (which == 1 ? out.d[1] : out.d[2]);
    const long long ldd = which == 0 ? out.ldd[0] : (which == 1 ? out.ldd[1] : out.ldd[2]);
    *reinterpret_cast<float2*>(pd + row * ldd + col) = make_float2(x, y);
  }
}

__global__ void __launch_bounds__(VA_THREADS)
vit_attn_bwd_mma_kernel(const VitAttnP p, const float* __restrict__ lse, const float* __restrict__ dout,
                        const __nv_bfloat16* __restrict__ dout_hi, const __nv_bfloat16* __restrict__ dout_lo, long long lddo,
                        const VitBwdOut out) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16 *Qh = base, *Ql = Qh + NT * LDH, *Kh = Ql + NT * LDH, *Kl = Kh + NT * LDH, *Vh = Kl + NT * LDH, *Vl = Vh + NT * LDH,
                *Dh = Vl + NT * LDH, *Dl = Dh + NT * LDH;
  // P~ and dS (phase 2 operands) reuse the K / V tiles, which are dead once every warp has finished phase 1
  __nv_bfloat16 *Ph = Kh, *Pl = Kl, *Sh = Vh, *Sl = Vl;
  const int h = blockIdx.x % p.nh, b = blockIdx.x / p.nh;
  const int n = p.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long rowbase = (long long)b * n;
  const DropKey dkey = drop_key_of(p.drop);

  // two cp.async groups: {Q, K} first so that S = Q K^T can start while {V, dO} are still in flight
  if (p.qh != nullptr) {
    stage_copy(p.qh + rowbase * p.ldq + (long long)h * HD, p.ql + rowbase * p.ldq + (long long)h * HD, p.ldq, n, Qh, Ql);
    stage_copy(p.kh + rowbase * p.ldk + (long long)h * HD, p.kl + rowbase * p.ldk + (long long)h * HD, p.ldk, n, Kh, Kl);
    cp_async_commit();
    stage_copy(p.vh + rowbase * p.ldv + (long long)h * HD, p.vl + rowbase * p.ldv + (long long)h * HD, p.ldv, n, Vh, Vl);
  } else {
    stage_split(p.q + rowbase * p.ldq + (long long)h * HD, p.ldq, n, Qh, Ql);
    stage_split(p.k + rowbase * p.ldk + (long long)h * HD, p.ldk, n, Kh, Kl);
    cp_async_commit();
    stage_split(p.v + rowbase * p.ldv + (long long)h * HD, p.ldv, n, Vh, Vl);
  }
  if (dout != nullptr) {
    stage_split(dout + rowbase * lddo + (long long)h * HD, lddo, n, Dh, Dl);
  } else {
    stage_copy(dout_hi + rowbase * lddo + (long long)h * HD, dout_lo + rowbase * lddo + (long long)h * HD, lddo, n, Dh, Dl);
  }
  cp_async_commit();
  cp_async_wait_group<1>();
  __syncthreads();

  // ---------------- phase 1: this warp's 16 query rows
  {
    uint32_t ah[4][4], al[4][4];
    // A fragments of Q and dO from shared memory (non-transposed)
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldsm_x4(ah[kk], Qh + (warp * 16 + (lane & 15)) * LDH + kk * 16 + (lane >> 4) * 8);
      ldsm_x4(al[kk], Ql + (warp * 16 + (lane & 15)) * LDH + kk * 16 + (lane >> 4) * 8);
    }
    mma_ABt(s, ah, al, Kh, Kl, lane);
    cp_async_wait_group<0>();
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldsm_x4(ah[kk], Dh + (warp * 16 + (lane & 15)) * LDH + kk * 16 + (lane >> 4) * 8);
      ldsm_x4(al[kk], Dl + (warp * 16 + (lane & 15)) * LDH + kk * 16 + (lane >> 4) * 8);
    }
    mma_ABt(dp, ah, al, Vh, Vl, lane);

    // P (into s) and delta_i = dO_i . O_i = sum_j P~_ij dP_ij  (O = P~ V and dP = dO V^T, so the forward output is never
    // re-read); then P~ (into s) and scale*dS (into dp).  The dropout decisions of pass 1 are kept as a bit mask.
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = warp * 16 + g + rr * 8;
      const bool rv = row < n;
      const float l = rv ? lse[((long long)b * p.nh + h) * n + row] : 0.f;
      const unsigned long long ibase = (((unsigned long long)b * p.nh + h) * n + (unsigned long long)(rv ? row : 0)) * (unsigned long long)n;
      uint32_t keep = 0u;
      float dl = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = 8 * j + 2 * t;
        float f0 = 1.f, f1 = 1.f;
        if (rv && col < n) drop_pair(p, dkey, ibase + col, f0, f1);  // f1 is unused when col + 1 == n (p1 = 0)
        const float p0 = (rv && col < n) ? __expf(s[j][rr * 2] * p.scale - l) : 0.f;
        const float p1 = (rv && col + 1 < n) ? __expf(s[j][rr * 2 + 1] * p.scale - l) : 0.f;
        keep |= (f0 != 0.f ? 1u : 0u) << (2 * j);
        keep |= (f1 != 0.f ? 1u : 0u) << (2 * j + 1);
        s[j][rr * 2] = p0;
        s[j][rr * 2 + 1] = p1;
        dl = fmaf(p0 * f0, dp[j][rr * 2], dl);
        dl = fmaf(p1 * f1, dp[j][rr * 2 + 1], dl);
      }
      dl += __shfl_xor_sync(0xffffffffu, dl, 1);
      dl += __shfl_xor_sync(0xffffffffu, dl, 2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f0 = ((keep >> (2 * j)) & 1u) ? p.dscale : 0.f;
        const float f1 = ((keep >> (2 * j + 1)) & 1u) ? p.dscale : 0.f;
        const float p0 = s[j][rr * 2], p1 = s[j][rr * 2 + 1];
        s[j][rr * 2] = p0 * f0;
        s[j][rr * 2 + 1] = p1 * f1;
        dp[j][rr * 2] = p0 * (dp[j][rr * 2] * f0 - dl) * p.scale;
        dp[j][rr * 2 + 1] = p1 * (dp[j][rr * 2 + 1] * f1 - dl) * p.scale;
      }
    }
    // dQ = dS K   (contract over keys; K read transposed)
    c_to_a(dp, ah, al);
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    mma_AB(acc, ah, al, Kh, Kl, lane);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = warp * 16 + g + rr * 8;
      if (row < n) {
#pragma unroll
        for (int j = 0; j < 8; ++j) store_pair(out, 0, rowbase + row, h * HD + 8 * j + 2 * t, acc[j][rr * 2], acc[j][rr * 2 + 1]);
      }
    }
    __syncthreads();  // every warp is done reading K and V: their tiles become the P~ / dS tiles
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = warp * 16 + g + rr * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = 8 * j + 2 * t;
        uint32_t hi, lo;
        split2(s[j][rr * 2], s[j][rr * 2 + 1], hi, lo);
        *reinterpret_cast<uint32_t*>(Ph + row * LDH + col) = hi;
        *reinterpret_cast<uint32_t*>(Pl + row * LDH + col) = lo;
        split2(dp[j][rr * 2], dp[j][rr * 2 + 1], hi, lo);
        *reinterpret_cast<uint32_t*>(Sh + row * LDH + col) = hi;
        *reinterpret_cast<uint32_t*>(Sl + row * LDH + col) = lo;
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: this warp's 16 key rows: dV = P~^T dO, dK = dS^T Q (contract over the 64 query rows)
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* Ah = which == 0 ? Ph : Sh;
    const __nv_bfloat16* Al = which == 0 ? Pl : Sl;
    const __nv_bfloat16* Bh = which == 0 ? Dh : Qh;
    const __nv_bfloat16* Bl = which == 0 ? Dl : Ql;
    uint32_t ah[4][4], al[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      // A[m = key][k = query row] = X[query row][key]: transposed read; matrices (k0-7,m0-7),(k0-7,m8-15),(k8-15,m0-7),(k8-15,m8-15)
      const int q8 = lane >> 3;
      const int row = kk * 16 + (lane & 7) + (q8 >> 1) * 8, col = warp * 16 + (q8 & 1) * 8;
      ldsm_x4_t(ah[kk], Ah + row * LDH + col);
      ldsm_x4_t(al[kk], Al + row * LDH + col);
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    mma_AB(acc, ah, al, Bh, Bl, lane);
    const int which_out = which == 0 ? 2 : 1;  // dv : dk
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = warp * 16 + g + rr * 8;
      if (row < n) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          store_pair(out, which_out, rowbase + row, h * HD + 8 * j + 2 * t, acc[j][rr * 2], acc[j][rr * 2 + 1]);
      }
    }
  }
}

VitAttnP make_p(const AttnDesc& a) {
  VitAttnP p;
  p.q = a.q; p.k = a.k; p.v = a.v; p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv;
  p.qh = reinterpret_cast<const __nv_bfloat16*>(a.q_hi); p.ql = reinterpret_cast<const __nv_bfloat16*>(a.q_lo);
  p.kh = reinterpret_cast<const __nv_bfloat16*>(a.k_hi); p.kl = reinterpret_cast<const __nv_bfloat16*>(a.k_lo);
  p.vh = reinterpret_cast<const __nv_bfloat16*>(a.v_hi); p.vl = reinterpret_cast<const __nv_bfloat16*>(a.v_lo);
  p.n = a.Tq; p.nh = a.nh; p.scale = a.scale;
  p.drop = a.drop;
  p.thresh = dropout_threshold(a.drop.p);
  p.dscale = a.drop.p > 0.f ? 1.0f / (1.0f - a.drop.p) : 1.0f;
  return p;
}

constexpr size_t FWD_SMEM = (size_t)4 * NT * LDH * 2;
constexpr size_t BWD_SMEM = (size_t)8 * NT * LDH * 2;

}  // namespace

bool vit_attention_eligible(const AttnDesc& a) {
  const int al = a.q_hi != nullptr ? 8 : 4;  // 16-byte rows for the bf16 copies, float4 for the fp32 conversion
  if (a.q_hi != nullptr && !(a.q_lo && a.k_hi && a.k_lo && a.v_hi && a.v_lo)) return false;
  return a.mask == VC_MASK_NONE && a.d == HD && a.Tq == a.Tk && a.Tq <= NT && a.ldq % al == 0 && a.ldk % al == 0 && a.ldv % al == 0;
}

int vit_attention_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s) {
  if (ldo % 2 != 0) return set_error("vit_attention_fwd: ldo must be even");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = true;
  }
  VC_LAUNCH((vit_attn_fwd_mma_kernel), a.B * a.nh, VA_THREADS, FWD_SMEM, reinterpret_cast<cudaStream_t>(s), 
      make_p(a), reinterpret_cast<__nv_bfloat16*>(o_hi), reinterpret_cast<__nv_bfloat16*>(o_lo), ldo, lse);
  return check_launch("vit_attn_fwd_mma_kernel");
}

static int launch_vit_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                          const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, const VitBwdOut& out,
                          stream_t s) {
  (void)o_hi; (void)o_lo; (void)ldo;  // delta = rowsum(dO * O) is rebuilt from P~ and dP inside the kernel
  if (lddo % (dout ? 4 : 8) != 0) return set_error("vit_attention_bwd: unsupported strides");
  if (!dout && !(dout_hi && dout_lo)) return set_error("vit_attention_bwd: no upstream gradient");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    configured = true;
  }
  VC_LAUNCH((vit_attn_bwd_mma_kernel), a.B * a.nh, VA_THREADS, BWD_SMEM, reinterpret_cast<cudaStream_t>(s), 
      make_p(a), lse, dout,
      reinterpret_cast<const __nv_bfloat16*>(dout_hi), reinterpret_cast<const __nv_bfloat16*>(dout_lo), lddo, out);
  return check_launch("vit_attn_bwd_mma_kernel");
}

int vit_attention_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
                      int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, stream_t s) {
  if (lddq % 2 != 0 || lddk % 2 != 0 || lddv % 2 != 0) return set_error("vit_attention_bwd: unsupported strides");
  VitBwdOut out;
  out.d[0] = dq; out.d[1] = dk; out.d[2] = dv;
  out.ldd[0] = lddq; out.ldd[1] = lddk; out.ldd[2] = lddv;
  for (int i = 0; i < 3; ++i) out.s_hi[i] = out.s_lo[i] = nullptr;
  out.lds = 0;
  return launch_vit_bwd(a, o_hi, o_lo, ldo, lse, dout, nullptr, nullptr, lddo, out, s);
}

int vit_attention_bwd_split(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                            const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi, bf16_t* dk_lo,
                            bf16_t* dv_hi, bf16_t* dv_lo, int64_t lds, stream_t s) {
  if (lds % 2 != 0) return set_error("vit_attention_bwd_split: unsupported stride");
  VitBwdOut out;
  for (int i = 0; i < 3; ++i) { out.d[i] = nullptr; out.ldd[i] = 0; }
  out.s_hi[0] = reinterpret_cast<__nv_bfloat16*>(dq_hi); out.s_lo[0] = reinterpret_cast<__nv_bfloat16*>(dq_lo);
  out.s_hi[1] = reinterpret_cast<__nv_bfloat16*>(dk_hi); out.s_lo[1] = reinterpret_cast<__nv_bfloat16*>(dk_lo);
  out.s_hi[2] = reinterpret_cast<__nv_bfloat16*>(dv_hi); out.s_lo[2] = reinterpret_cast<__nv_bfloat16*>(dv_lo);
  out.lds = lds;
  return launch_vit_bwd(a, o_hi, o_lo, ldo, lse, dout, dout_hi, dout_lo, lddo, out, s);
}

}  // namespace vck
