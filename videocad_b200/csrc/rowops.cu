// Row-wise / element-wise sm_100a kernels of the VideoCAD hot path (HBM-bound; 128-bit vectorised, warp-shuffle
// reductions): LayerNorm fwd/bwd (vit_pytorch LayerNorms, TransformerDecoderLayer.norm1-3), patchify+LayerNorm
// (vit_pytorch to_patch_embedding[0:2]), CLS/pos token assembly + embedding dropout, activation/dropout backward,
// fp32 -> split-bf16 conversion, and the narrow linears that are not tensor-core shaped (embed_action K=7,
// predict_action_class_0_4 N=5; /root/reference/model/autoregressive_transformer.py:64,77,110-113).
#include <cuda_runtime.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

inline cudaStream_t cs(stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ void drop4(float (&v)[4], const DropKey& key, uint32_t thresh, float scale, unsigned long long idx) {
  const Rand4 w = dropout_words(key, idx >> 2);
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = (w.v[i] >= thresh) ? v[i] * scale : 0.0f;
}

// ------------------------------------------------------------------------------------------- split
__global__ void split_kernel(const float* __restrict__ x, long long ldx, long long rows, long long cols4,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldo) {
  pdl_grid_sync();
  const long long total = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4, c = (i % cols4) * 4;
    const float4 v4 = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    uint2 h, l;
    split4(v, h, l);
    *reinterpret_cast<uint2*>(hi + r * ldo + c) = h;
    *reinterpret_cast<uint2*>(lo + r * ldo + c) = l;
  }
}

// one launch for all weight matrices: block -> item by binary search over block_start (n_items is small)
__global__ void __launch_bounds__(256) split_many_kernel(const vc_split_item* __restrict__ items, int n_items) {
  pdl_grid_sync();
  int lo = 0, hi = n_items - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].block_start <= b) lo = mid; else hi = mid - 1;
  }
  const vc_split_item it = items[lo];
  const long long base = (b - it.block_start) * 1024;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < it.n4) {
      const float4 v4 = *reinterpret_cast<const float4*>(it.src + i * 4);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      uint2 h, l;
      split4(v, h, l);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(it.hi) + i * 4) = h;
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(it.lo) + i * 4) = l;
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm
constexpr int LN_MAXV = 8;  // float4 chunks per lane -> C <= 1024
constexpr int LN_WARPS = 4;

struct RowLoader {  // plain strided rows
  const float* x; long long ldx;
  static constexpr bool kContiguous = true;
  __device__ __forceinline__ float4 load(long long row, int c) const {
    return *reinterpret_cast<const float4*>(x + row * ldx + c);
  }
  __device__ __forceinline__ const float* ptr(long long row, int c) const { return x + row * ldx + c; }
};
struct PatchLoader {  // 'f 1 (h 32) (w 32) -> (f h w) (32 32)' gather; C = 1024
  const float* img; int S, wp, N;
  static constexpr bool kContiguous = false;
  __device__ __forceinline__ const float* ptr(long long, int) const { return nullptr; }
  __device__ __forceinline__ float4 load(long long row, int c) const {
    const long long f = row / N;
    const int pidx = (int)(row % N);
    const int ph = pidx / wp, pw = pidx % wp;
    const int p1 = c >> 5, p2 = c & 31;
    return *reinterpret_cast<const float4*>(img + (f * S + (ph * 32 + p1)) * (long long)S + pw * 32 + p2);
  }
};

template <class Loader, int NV>
__device__ __forceinline__ void ln_load_stats(const Loader& ld, long long row, int C, int lane, float4 (&v)[NV],
                                              float& mean, float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      v[i] = ld.load(row, c);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  rstd = rsqrtf(warp_sum(q) / (float)C + eps);
}

template <class Loader, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(Loader ld, long long rows, int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              float* __restrict__ y, long long ldy, __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo,
              long long ldys, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[NV];
  float mean, rstd;
  ln_load_stats<Loader, NV>(ld, row, C, lane, v, mean, rstd, eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float o[4];
      o[0] = (v[i].x - mean) * rstd * g.x + b.x;
      o[1] = (v[i].y - mean) * rstd * g.y + b.y;
      o[2] = (v[i].z - mean) * rstd * g.z + b.z;
      o[3] = (v[i].w - mean) * rstd * g.w + b.w;
      if (y) *reinterpret_cast<float4*>(y + row * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (y_hi) {
        uint2 h, l;
        split4(o, h, l);
        *reinterpret_cast<uint2*>(y_hi + row * ldys + c) = h;
        if (y_lo) *reinterpret_cast<uint2*>(y_lo + row * ldys + c) = l;
      }
    }
  }
}

// backward: block loops over rows (grid-stride by warps); dgamma/dbeta partials in registers, one atomic pass at the end
struct LnFuse {  // optional second output: g = dx * dropout_mask -> split-bf16 (+ column sums), see layernorm_bwd_fused
  Drop drop; uint32_t thresh; float scale;
  __nv_bfloat16 *g_hi, *g_lo; long long ldg;
  float* colsum;
};

// one row's operands in registers (loaded one row ahead of their use, see ln_bwd_kernel)
template <int NV>
struct LnBwdRow {
  float4 x[NV], d[NV], r[NV];
  float mu, rs;
};

template <class Loader, bool kNeedDx, int NV>
__device__ __forceinline__ void ln_bwd_load(const Loader& ld, const float* __restrict__ dy, long long lddy, const float* __restrict__ mean,
                                            const float* __restrict__ rstd, const float* __restrict__ dres, long long lddres,
                                            long long row, int C, int lane, LnBwdRow<NV>& R) {
  R.mu = mean[row];
  R.rs = rstd[row];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      R.x[i] = ld.load(row, c);
      R.d[i] = *reinterpret_cast<const float4*>(dy + row * lddy + c);
      if (kNeedDx && dres) R.r[i] = *reinterpret_cast<const float4*>(dres + row * lddres + c);
    }
  }
}

template <bool kNeedDx, bool kFuse, int NV>
__device__ __forceinline__ void ln_bwd_row(LnBwdRow<NV>& R, long long row, int C, int lane, float invC, const float* __restrict__ gamma,
                                           bool has_res, float* __restrict__ dx, long long lddx, const LnFuse& fuse, const DropKey& fkey, float4 (&dg)[NV],
                                           float4 (&db)[NV], float4 (&cs)[kFuse ? NV : 1]) {
  const float mu = R.mu, rs = R.rs;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      const float4 xv = R.x[i];
      const float4 d = R.d[i];
      const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      R.x[i] = xh;
      dg[i].x += d.x * xh.x; dg[i].y += d.y * xh.y; dg[i].z += d.z * xh.z; dg[i].w += d.w * xh.w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      if (kNeedDx) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        R.d[i] = g;
        s1 += g.x + g.y + g.z + g.w;
        s2 += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
      }
    }
  }
  if (kNeedDx) {
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane * 4 + i * 128;
      if (c < C) {
        const float4 xh = R.x[i], g = R.d[i];
        float4 o;
        o.x = rs * (g.x - s1 - xh.x * s2);
        o.y = rs * (g.y - s1 - xh.y * s2);
        o.z = rs * (g.z - s1 - xh.z * s2);
        o.w = rs * (g.w - s1 - xh.w * s2);
        if (has_res) {
          const float4 r = R.r[i];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dx + row * lddx + c) = o;
        if (kFuse) {
          float v[4] = {o.x, o.y, o.z, o.w};
          if (fuse.drop.p > 0.f) drop4(v, fkey, fuse.thresh, fuse.scale, (unsigned long long)row * C + c);
          uint2 h, l;
          split4(v, h, l);
          *reinterpret_cast<uint2*>(fuse.g_hi + row * fuse.ldg + c) = h;
          *reinterpret_cast<uint2*>(fuse.g_lo + row * fuse.ldg + c) = l;
          cs[i].x += v[0]; cs[i].y += v[1]; cs[i].z += v[2]; cs[i].w += v[3];
        }
      }
    }
  }
}

// Each warp walks its rows with a stride.  The operands of the next LNB_DEPTH - 1 rows (x, dy, residual gradient: 12 x 128-bit
// loads per lane and row at C = 512) are kept in flight with cp.async into a per-warp shared-memory ring: every lane copies and
// later reads only ITS OWN 16-byte slots, so the only synchronisation is cp.async.wait_group.  History: alternating load and
// compute phases reached ~3 TB/s; a one-row register prefetch needed 183 registers (two resident CTAs per SM); the ring needs
// registers for one row only (145, three CTAs per SM).  Measured: inside the training step the ring version is the faster one
// (image-encoder backward 4.05 -> 3.95 ms, +1.3 % frames/s), but ISOLATED under ncu it is slower than the register prefetch
// (43.2 vs 38.5 us at [12800, 512], long-scoreboard and end-of-kernel barrier stalls up; profiles/r01l_ncu_attn_ln_summary.txt):
// the kernel still reaches only ~3 TB/s of its 131 MB and is a listed open item.
constexpr int LNB_DEPTH = 3;

__device__ __forceinline__ void ln_cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ln_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ln_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class Loader, bool kNeedDx, bool kFuse, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(Loader ld, const float* __restrict__ dy, long long lddy, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, long long rows, int C,
              const float* __restrict__ dres, long long lddres, float* __restrict__ dx, long long lddx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, const LnFuse fuse) {
  pdl_grid_sync();
  __shared__ float4 red[LN_WARPS][32];
  extern __shared__ float4 ln_ring[];  // [LN_WARPS][LNB_DEPTH][3 (x, dy, dres)][NV][32 lanes] (ring variant only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 dg[NV], db[NV], cs[kFuse ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (kFuse ? NV : 1); ++i) cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float invC = 1.0f / (float)C;
  const DropKey fkey = kFuse ? drop_key_of(fuse.drop) : DropKey{0u, 0u};
  const bool has_res = kNeedDx && dres != nullptr;
  const long long stride = (long long)gridDim.x * LN_WARPS;
  long long row = (long long)blockIdx.x * LN_WARPS + warp;
  constexpr bool kRing = Loader::kContiguous && NV <= 4;  // contiguous rows of <= 512 columns: the image encoders and the decoder
  if constexpr (kRing) {
    float4* wring = ln_ring + (size_t)warp * (LNB_DEPTH * 3 * NV * 32);
    auto issue = [&](long long r, int slot) {
      if (r < rows) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane * 4 + i * 128;
          if (c < C) {
            ln_cp_async16(&wring[((slot * 3 + 0) * NV + i) * 32 + lane], ld.ptr(r, c));
            ln_cp_async16(&wring[((slot * 3 + 1) * NV + i) * 32 + lane], dy + r * lddy + c);
            if (has_res) ln_cp_async16(&wring[((slot * 3 + 2) * NV + i) * 32 + lane], dres + r * lddres + c);
          }
        }
      }
      ln_cp_async_commit();  // one group per ring step, empty past the last row: the wait below counts groups
    };
#pragma unroll
    for (int dpt = 0; dpt < LNB_DEPTH - 1; ++dpt) issue(row + dpt * stride, dpt);
    float mu_n = 0.f, rs_n = 0.f;
    if (row < rows) { mu_n = mean[row]; rs_n = rstd[row]; }
    int slot = 0;
    while (row < rows) {
      int nslot = slot + LNB_DEPTH - 1;
      if (nslot >= LNB_DEPTH) nslot -= LNB_DEPTH;
      issue(row + (LNB_DEPTH - 1) * stride, nslot);  // overwrites the slot consumed in the previous iteration
      LnBwdRow<NV> A;
      A.mu = mu_n; A.rs = rs_n;
      if (row + stride < rows) { mu_n = mean[row + stride]; rs_n = rstd[row + stride]; }
      ln_cp_async_wait<LNB_DEPTH - 1>();  // this row's group has landed
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < C) {
          A.x[i] = wring[((slot * 3 + 0) * NV + i) * 32 + lane];
          A.d[i] = wring[((slot * 3 + 1) * NV + i) * 32 + lane];
          if (has_res) A.r[i] = wring[((slot * 3 + 2) * NV + i) * 32 + lane];
        }
      }
      ln_bwd_row<kNeedDx, kFuse, NV>(A, row, C, lane, invC, gamma, has_res, dx, lddx, fuse, fkey, dg, db, cs);
      row += stride;
      slot = slot + 1 == LNB_DEPTH ? 0 : slot + 1;
    }
    ln_cp_async_wait<0>();
  } else {
    LnBwdRow<NV> A;
    for (; row < rows; row += stride) {
      ln_bwd_load<Loader, kNeedDx, NV>(ld, dy, lddy, mean, rstd, dres, lddres, row, C, lane, A);
      ln_bwd_row<kNeedDx, kFuse, NV>(A, row, C, lane, invC, gamma, has_res, dx, lddx, fuse, fkey, dg, db, cs);
    }
  }
  // cross-warp reduction of the parameter grads, then one atomicAdd per column per block
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c >= C) break;  // uniform across the warp's lanes only for full chunks; C % 128 == 0 is required
    for (int pass = 0; pass < (kFuse ? 3 : 2); ++pass) {
      if (pass == 2 && fuse.colsum == nullptr) break;  // uniform
      red[warp][lane] = pass == 0 ? dg[i] : (pass == 1 ? db[i] : cs[kFuse ? i : 0]);
      __syncthreads();
      if (warp == 0) {
        float4 a = red[0][lane];
#pragma unroll
        for (int w = 1; w < LN_WARPS; ++w) {
          a.x += red[w][lane].x; a.y += red[w][lane].y; a.z += red[w][lane].z; a.w += red[w][lane].w;
        }
        float* dst = (pass == 0 ? dgamma : (pass == 1 ? dbeta : fuse.colsum)) + c;
        atomicAdd(dst + 0, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------- ViT token assembly
__global__ void vit_assemble_fwd_kernel(const float* __restrict__ e, int F, int N, int C, const float* __restrict__ cls,
                                        const float* __restrict__ pos, Drop drop, uint32_t thresh, float scale,
                                        float* __restrict__ x) {
  pdl_grid_sync();
  const int n = N + 1, C4 = C / 4;
  const DropKey dkey = drop_key_of(drop);
  const long long total = (long long)F * n * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long ft = i / C4;
    const int t = (int)(ft % n);
    const long long f = ft / n;
    const float4 a = (t == 0) ? __ldg(reinterpret_cast<const float4*>(cls + c))
                              : *reinterpret_cast<const float4*>(e + (f * N + (t - 1)) * (long long)C + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * C + c));
    float v[4] = {a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w};
    if (drop.p > 0.f) drop4(v, dkey, thresh, scale, (unsigned long long)ft * C + c);
    *reinterpret_cast<float4*>(x + ft * C + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// grid: (ceil(n*C4/128), f-chunks); thread <-> (t, c4); loops over its f-chunk
__global__ void vit_assemble_bwd_kernel(const float* __restrict__ dx, int F, int N, int C, Drop drop, uint32_t thresh,
                                        float scale, float* __restrict__ de, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int f_per_block) {
  pdl_grid_sync();
  const int n = N + 1, C4 = C / 4;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n * C4) return;
  const DropKey dkey = drop_key_of(drop);
  const int t = j / C4, c = (j % C4) * 4;
  const int f0 = blockIdx.y * f_per_block;
  const int f1 = min(F, f0 + f_per_block);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int f = f0; f < f1; ++f) {
    const long long ft = (long long)f * n + t;
    const float4 d = *reinterpret_cast<const float4*>(dx + ft * C + c);
    float v[4] = {d.x, d.y, d.z, d.w};
    if (drop.p > 0.f) drop4(v, dkey, thresh, scale, (unsigned long long)ft * C + c);
    if (t > 0) *reinterpret_cast<float4*>(de + ((long long)f * N + (t - 1)) * C + c) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(dpos + (long long)t * C + c + i, acc[i]);
    if (t == 0) atomicAdd(dcls + c + i, acc[i]);
  }
}

// ------------------------------------------------------------------------------------------- act/dropout backward
// block = (128 column quads) x (ADB_RG row groups); grid = (ceil(N4/128), row chunks).  Each thread walks the rows
// r0 + ty, r0 + ty + ADB_RG, ... of its chunk; the ADB_RG partial column sums are combined through shared memory so that
// every block issues ONE atomicAdd per column (bias gradient).
constexpr int ADB_RG = 4;
__global__ void __launch_bounds__(128 * ADB_RG)
act_dropout_bwd_kernel(const float* __restrict__ dy, long long lddy, long long M, int N, int act,
                       const float* __restrict__ aux, long long ldaux, const __nv_bfloat16* __restrict__ aux_hi,
                       long long ldaux_hi, Drop drop, uint32_t thresh, float scale, float* __restrict__ g, long long ldg,
                       __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, long long ldgs,
                       float* __restrict__ colsum, int rows_per_block) {
  pdl_grid_sync();
  __shared__ float4 red[ADB_RG][128];
  const int q = blockIdx.x * 128 + threadIdx.x;
  const bool active = q * 4 < N;
  const int c = q * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const DropKey dkey = drop_key_of(drop);
  if (active) {
    for (long long r = r0 + threadIdx.y; r < r1; r += ADB_RG) {
      const float4 d = *reinterpret_cast<const float4*>(dy + r * lddy + c);
      float v[4] = {d.x, d.y, d.z, d.w};
      if (drop.p > 0.f) drop4(v, dkey, thresh, scale, (unsigned long long)r * N + c);
      if (act == ACT_GELU) {
        const float4 a = *reinterpret_cast<const float4*>(aux + r * ldaux + c);
        v[0] *= gelu_grad_f(a.x); v[1] *= gelu_grad_f(a.y); v[2] *= gelu_grad_f(a.z); v[3] *= gelu_grad_f(a.w);
      } else if (act == ACT_TANH) {
        const float4 a = *reinterpret_cast<const float4*>(aux + r * ldaux + c);
        v[0] *= 1.f - a.x * a.x; v[1] *= 1.f - a.y * a.y; v[2] *= 1.f - a.z * a.z; v[3] *= 1.f - a.w * a.w;
      } else if (act == ACT_RELU) {
        const uint2 a = *reinterpret_cast<const uint2*>(aux_hi + r * ldaux_hi + c);
        if ((a.x & 0x7fffu) == 0u) v[0] = 0.f;
        if ((a.x & 0x7fff0000u) == 0u) v[1] = 0.f;
        if ((a.y & 0x7fffu) == 0u) v[2] = 0.f;
        if ((a.y & 0x7fff0000u) == 0u) v[3] = 0.f;
      }
      if (g) *reinterpret_cast<float4*>(g + r * ldg + c) = make_float4(v[0], v[1], v[2], v[3]);
      if (g_hi) {
        uint2 h, l;
        split4(v, h, l);
        *reinterpret_cast<uint2*>(g_hi + r * ldgs + c) = h;
        if (g_lo) *reinterpret_cast<uint2*>(g_lo + r * ldgs + c) = l;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += v[i];
    }
  }
  if (colsum) {
    red[threadIdx.y][threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (threadIdx.y == 0 && active) {
      float4 a = red[0][threadIdx.x];
#pragma unroll
      for (int k = 1; k < ADB_RG; ++k) {
        const float4 b = red[k][threadIdx.x];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      atomicAdd(colsum + c + 0, a.x); atomicAdd(colsum + c + 1, a.y); atomicAdd(colsum + c + 2, a.z); atomicAdd(colsum + c + 3, a.w);
    }
  }
}

__global__ void row_reduce_mod_kernel(const float* __restrict__ x, long long ldx, long long M, int N, int div, int mod,
                                      float* __restrict__ out) {
  pdl_grid_sync();
  const long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int n = (int)(i % N);
    atomicAdd(out + ((m / div) % mod) * N + n, x[m * ldx + n]);
  }
}

__global__ void broadcast_rows_kernel(const float* __restrict__ src, long long lds, long long M, int N, int div,
                                      float* __restrict__ dst, long long ldd, __nv_bfloat16* __restrict__ d_hi,
                                      __nv_bfloat16* __restrict__ d_lo, long long ldds) {
  pdl_grid_sync();
  const int N4 = N / 4;
  const long long total = M * N4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N4;
    const int c = (int)(i % N4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(src + (m / div) * lds + c);
    if (dst) *reinterpret_cast<float4*>(dst + m * ldd + c) = a;
    if (d_hi) {
      const float v[4] = {a.x, a.y, a.z, a.w};
      uint2 h, l;
      split4(v, h, l);
      *reinterpret_cast<uint2*>(d_hi + m * ldds + c) = h;
      if (d_lo) *reinterpret_cast<uint2*>(d_lo + m * ldds + c) = l;
    }
  }
}

// ------------------------------------------------------------------------------------------- embed_action (K = 7)
__global__ void embed_action_fwd_kernel(const float* __restrict__ actions, long long R, int A, int H,
                                        const float* __restrict__ W, const float* __restrict__ b,
                                        const float* __restrict__ E, int T, float* __restrict__ y,
                                        __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo) {
  pdl_grid_sync();
  const long long total = R * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int h = (int)(i % H);
    float acc = b[h];
    for (int a = 0; a < A; ++a) acc += actions[r * A + a] * W[(long long)h * A + a];
    if (E) acc += E[(r % T) * H + h];
    const float v = tanhf(acc);
    if (y) y[i] = v;
    if (y_hi) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      y_hi[i] = hi;
      if (y_lo) y_lo[i] = lo;
    }
  }
}

// grid: (ceil(H/128), row chunks); thread <-> h
__global__ void embed_action_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                        const float* __restrict__ actions, long long R, int A, int H, int T,
                                        float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dE,
                                        int rows_per_block) {
  pdl_grid_sync();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float accW[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float accb = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float yv = y[r * H + h];
    const float d = dy[r * H + h] * (1.f - yv * yv);
    accb += d;
    for (int a = 0; a < A && a < 8; ++a) accW[a] += d * actions[r * A + a];
    if (dE) atomicAdd(dE + (r % T) * H + h, d);
  }
  atomicAdd(db + h, accb);
  for (int a = 0; a < A && a < 8; ++a) atomicAdd(dW + (long long)h * A + a, accW[a]);
}

// ------------------------------------------------------------------------------------------- narrow head (N = 5)
constexpr int HEAD_MAXC = 8;
__global__ void head_small_fwd_kernel(const float* __restrict__ x, long long R, int H, const float* __restrict__ W,
                                      const float* __restrict__ b, int C, float* __restrict__ out) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float acc[HEAD_MAXC];
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) acc[c] = 0.f;
  for (int h = lane * 4; h < H; h += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + r * H + h);
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) {
      if (c < C) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)c * H + h));
        acc[c] += xv.x * w.x + xv.y * w.y + xv.z * w.z + xv.w * w.w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) {
    if (c < C) {
      const float s = warp_sum(acc[c]);
      if (lane == 0) out[r * C + c] = s + b[c];
    }
  }
}

// grid: (ceil(H4/128), row chunks); thread <-> h quad
__global__ void head_small_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, long long R, int H,
                                      const float* __restrict__ W, int C, float* __restrict__ dx, int accumulate_dx,
                                      float* __restrict__ dW, float* __restrict__ db, int rows_per_block) {
  pdl_grid_sync();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= H) return;
  const int h = q * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float4 w[HEAD_MAXC], aw[HEAD_MAXC];
  float ab[HEAD_MAXC];
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) {
    w[c] = (c < C) ? __ldg(reinterpret_cast<const float4*>(W + (long long)c * H + h)) : make_float4(0.f, 0.f, 0.f, 0.f);
    aw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[c] = 0.f;
  }
  for (long long r = r0; r < r1; ++r) {
    const float4 xv = *reinterpret_cast<const float4*>(x + r * H + h);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) {
      if (c < C) {
        const float d = dout[r * C + c];
        o.x += d * w[c].x; o.y += d * w[c].y; o.z += d * w[c].z; o.w += d * w[c].w;
        aw[c].x += d * xv.x; aw[c].y += d * xv.y; aw[c].z += d * xv.z; aw[c].w += d * xv.w;
        ab[c] += d;
      }
    }
    float4* dst = reinterpret_cast<float4*>(dx + r * H + h);
    if (accumulate_dx) {
      const float4 p = *dst;
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *dst = o;
  }
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) {
    if (c < C) {
      float* dst = dW + (long long)c * H + h;
      atomicAdd(dst + 0, aw[c].x); atomicAdd(dst + 1, aw[c].y); atomicAdd(dst + 2, aw[c].z); atomicAdd(dst + 3, aw[c].w);
      if (q == 0) atomicAdd(db + c, ab[c]);
    }
  }
}

// Linear layer on M <= 16 rows: the rows live in shared memory, one warp per output column streams that column's K weights once
constexpr int LR_MAXM = 16, LR_WARPS = 8;
template <int MM>
__global__ void __launch_bounds__(LR_WARPS * 32)
linear_rows_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo, long long ldx,
                   int M, const float* __restrict__ W, const float* __restrict__ bias, int N, int K, int act,
                   const float* __restrict__ residual, long long ld_res, float* __restrict__ out_f32, long long ldo,
                   __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ldo_split) {
  pdl_grid_sync();
  extern __shared__ float4 lr_smem4[];
  float* xs = reinterpret_cast<float*>(lr_smem4);  // [MM][K], rows >= M zero
  const int K4 = K >> 2;
  for (int idx = threadIdx.x; idx < MM * K4; idx += LR_WARPS * 32) {
    const int m = idx / K4, k = (idx - m * K4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) {
      if (x != nullptr) {
        v = *reinterpret_cast<const float4*>(x + (long long)m * ldx + k);
      } else {
        const uint2 h = *reinterpret_cast<const uint2*>(x_hi + (long long)m * ldx + k);
        v = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16),
                        __uint_as_float(h.y & 0xffff0000u));
        if (x_lo != nullptr) {
          const uint2 l = *reinterpret_cast<const uint2*>(x_lo + (long long)m * ldx + k);
          v.x += __uint_as_float(l.x << 16); v.y += __uint_as_float(l.x & 0xffff0000u);
          v.z += __uint_as_float(l.y << 16); v.w += __uint_as_float(l.y & 0xffff0000u);
        }
      }
    }
    *reinterpret_cast<float4*>(xs + m * K + k) = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n = blockIdx.x * LR_WARPS + warp; n < N; n += gridDim.x * LR_WARPS) {
    float acc[MM];
#pragma unroll
    for (int m = 0; m < MM; ++m) acc[m] = 0.f;
    const float* wrow = W + (long long)n * K;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + m * K + k);
        acc[m] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[m]))));
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int m = 0; m < MM; ++m) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], off);
    }
    // lane m finishes row m
    float v = 0.f;
#pragma unroll
    for (int m = 0; m < MM; ++m) if (lane == m) v = acc[m];
    if (lane < M) {
      if (bias != nullptr) v += __ldg(bias + n);
      v = apply_act(v, act);
      if (residual != nullptr) v += residual[(long long)lane * ld_res + n];
      if (out_f32 != nullptr) out_f32[(long long)lane * ldo + n] = v;
      if (out_hi != nullptr) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        out_hi[(long long)lane * ldo_split + n] = h;
        if (out_lo != nullptr) out_lo[(long long)lane * ldo_split + n] = l;
      }
    }
  }
}

// uint8 frames -> normalised fp32 (16 pixels per thread: one 128-bit load, four 128-bit stores)
__global__ void frames_u8_normalize_kernel(const uint8_t* __restrict__ src, long long n, float mean, float std, float* __restrict__ dst) {
  pdl_grid_sync();
  const long long n16 = n >> 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 w = *reinterpret_cast<const uint4*>(src + 16 * i);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o;
      o.x = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(ws[k] & 0xffu), 255.0f), mean), std);
      o.y = __fdiv_rn(__fsub_rn(__fdiv_rn((float)((ws[k] >> 8) & 0xffu), 255.0f), mean), std);
      o.z = __fdiv_rn(__fsub_rn(__fdiv_rn((float)((ws[k] >> 16) & 0xffu), 255.0f), mean), std);
      o.w = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(ws[k] >> 24), 255.0f), mean), std);
      *reinterpret_cast<float4*>(dst + 16 * i + 4 * k) = o;
    }
  }
  // tail (n not a multiple of 16)
  for (long long i = (n16 << 4) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[i], 255.0f), mean), std);
}

__global__ void add_kernel(const float* __restrict__ a, const float* b, float* out, long long n) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

__global__ void dropout_mask_kernel(Drop drop, uint32_t thresh, float scale, long long n, float* out) {
  pdl_grid_sync();
  const long long n4 = (n + 3) / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const Rand4 w = dropout_words(drop_seed(drop), drop.site, (unsigned long long)i);
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < n) out[i * 4 + k] = (w.v[k] >= thresh) ? scale : 0.f;
  }
}

inline int ew_grid(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
inline float drop_scale(const Drop& d) { return d.p > 0.f ? 1.0f / (1.0f - d.p) : 1.0f; }

}  // namespace

// =================================================================================================== launchers
int split_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, bf16_t* hi, bf16_t* lo, int64_t ldo, stream_t s) {
  if (cols % 4 != 0 || ldx % 4 != 0 || ldo % 4 != 0) return set_error("split_f32: cols/ld must be multiples of 4");
  if (rows <= 0 || cols <= 0) return 0;
  VC_LAUNCH((split_kernel), ew_grid(rows * cols / 4, 256), 256, 0, cs(s), x, ldx, rows, cols / 4, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                reinterpret_cast<__nv_bfloat16*>(lo), ldo);
  return check_launch("split_kernel");
}

int split_many(const vc_split_item* items, int n_items, int64_t total_blocks, stream_t s) {
  if (n_items <= 0 || total_blocks <= 0) return 0;
  VC_LAUNCH((split_many_kernel), (unsigned)total_blocks, 256, 0, cs(s), items, n_items);
  return check_launch("split_many_kernel");
}

int layernorm_fwd(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  float* y, int64_t ldy, bf16_t* y_hi, bf16_t* y_lo, int64_t ldy_split, float* mean, float* rstd,
                  stream_t s) {
  if (C % 128 != 0 || C > 128 * LN_MAXV) return set_error("layernorm_fwd: C must be a multiple of 128 and <= 1024");
  if (rows <= 0) return 0;
  RowLoader ld{x, ldx};
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(y_hi);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(y_lo);
  const int grid = cdiv(rows, LN_WARPS);
  if (C <= 256) VC_LAUNCH((ln_fwd_kernel<RowLoader, 2>), grid, LN_WARPS * 32, 0, cs(s), ld, rows, C, gamma, beta, eps, y, ldy, yh, yl, ldy_split, mean, rstd);
  else if (C <= 512) VC_LAUNCH((ln_fwd_kernel<RowLoader, 4>), grid, LN_WARPS * 32, 0, cs(s), ld, rows, C, gamma, beta, eps, y, ldy, yh, yl, ldy_split, mean, rstd);
  else VC_LAUNCH((ln_fwd_kernel<RowLoader, 8>), grid, LN_WARPS * 32, 0, cs(s), ld, rows, C, gamma, beta, eps, y, ldy, yh, yl, ldy_split, mean, rstd);
  return check_launch("ln_fwd_kernel");
}

int layernorm_bwd_fused(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                        const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, Drop gdrop, bf16_t* g_hi, bf16_t* g_lo, int64_t ldg, float* g_colsum,
                        stream_t s) {
  if (C % 128 != 0 || C > 128 * LN_MAXV) return set_error("layernorm_bwd: C must be a multiple of 128 and <= 1024");
  if (rows <= 0) return 0;
  RowLoader ld{x, ldx};
  // one resident wave (3 CTAs/SM: 145 registers, 72 KiB of operand ring at C = 512): fewer blocks also means fewer gradient
  // atomics per column; decoder-sized problems (a few hundred rows) get one row per warp instead of four
  int grid = cdiv(rows, LN_WARPS * 4);
  if (grid < 148) grid = cdiv(rows, LN_WARPS) < 148 ? cdiv(rows, LN_WARPS) : 148;
  if (grid > 148 * 3) grid = 148 * 3;
  LnFuse f;
  f.drop = gdrop; f.thresh = dropout_threshold(gdrop.p); f.scale = drop_scale(gdrop);
  f.g_hi = reinterpret_cast<__nv_bfloat16*>(g_hi); f.g_lo = reinterpret_cast<__nv_bfloat16*>(g_lo); f.ldg = ldg;
  f.colsum = g_colsum;
#define VC_LN_BWD(FUSE, NVV)                                                                                              \
  do {                                                                                                                    \
    const size_t ring = (NVV) <= 4 ? sizeof(float4) * LN_WARPS * LNB_DEPTH * 3 * (NVV) * 32 : 0;                         \
    static bool configured = false;                                                                                       \
    if (ring > 48 * 1024 && !configured) {                                                                                \
      cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel<RowLoader, true, FUSE, NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring); \
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));                                                      \
      configured = true;                                                                                                  \
    }                                                                                                                     \
    VC_LAUNCH((ln_bwd_kernel<RowLoader, true, FUSE, NVV>), grid, LN_WARPS * 32, ring, cs(s), ld, dy, lddy, mean, rstd, gamma, rows, C, dres, \
                                                                                 lddres, dx, lddx, dgamma, dbeta, f);     \
  } while (0)
  if (g_hi != nullptr) {
    if (g_lo == nullptr) return set_error("layernorm_bwd_fused: g_lo required with g_hi");
    if (C <= 256) VC_LN_BWD(true, 2); else if (C <= 512) VC_LN_BWD(true, 4); else VC_LN_BWD(true, 8);
  } else {
    if (C <= 256) VC_LN_BWD(false, 2); else if (C <= 512) VC_LN_BWD(false, 4); else VC_LN_BWD(false, 8);
  }
#undef VC_LN_BWD
  return check_launch("ln_bwd_kernel");
}

int layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                  const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                  float* dgamma, float* dbeta, stream_t s) {
  return layernorm_bwd_fused(dy, lddy, x, ldx, mean, rstd, gamma, rows, C, dres, lddres, dx, lddx, dgamma, dbeta, no_drop(),
                             nullptr, nullptr, 0, nullptr, s);
}

int patch_layernorm_fwd(const float* img, int F, int S, const float* gamma, const float* beta, float eps, bf16_t* y_hi,
                        bf16_t* y_lo, float* mean, float* rstd, stream_t s) {
  if (S % 32 != 0) return set_error("patch_layernorm_fwd: image size must be a multiple of 32");
  const int wp = S / 32, N = wp * wp;
  const long long rows = (long long)F * N;
  if (rows <= 0) return 0;
  PatchLoader ld{img, S, wp, N};
  VC_LAUNCH((ln_fwd_kernel<PatchLoader, 8>), cdiv(rows, LN_WARPS), LN_WARPS * 32, 0, cs(s), 
      ld, rows, 1024, gamma, beta, eps, nullptr, 0, reinterpret_cast<__nv_bfloat16*>(y_hi),
      reinterpret_cast<__nv_bfloat16*>(y_lo), 1024, mean, rstd);
  return check_launch("ln_fwd_kernel<patch>");
}

int patch_layernorm_bwd_params(const float* img, int F, int S, const float* mean, const float* rstd, const float* dy,
                               float* dgamma, float* dbeta, stream_t s) {
  if (S % 32 != 0) return set_error("patch_layernorm_bwd_params: image size must be a multiple of 32");
  const int wp = S / 32, N = wp * wp;
  const long long rows = (long long)F * N;
  if (rows <= 0) return 0;
  PatchLoader ld{img, S, wp, N};
  int grid = cdiv(rows, LN_WARPS * 8);
  if (grid > 148 * 4) grid = 148 * 4;
  LnFuse f;
  f.drop = no_drop(); f.thresh = 0; f.scale = 1.f; f.g_hi = f.g_lo = nullptr; f.ldg = 0; f.colsum = nullptr;
  VC_LAUNCH((ln_bwd_kernel<PatchLoader, false, false, 8>), grid, LN_WARPS * 32, 0, cs(s), ld, dy, 1024, mean, rstd, nullptr, rows, 1024,
                                                                              nullptr, 0, nullptr, 0, dgamma, dbeta, f);
  return check_launch("ln_bwd_kernel<patch>");
}

int vit_assemble_fwd(const float* e, int F, int N, int C, const float* cls, const float* pos, Drop drop, float* x,
                     stream_t s) {
  if (C % 4 != 0) return set_error("vit_assemble_fwd: C % 4 != 0");
  const long long total = (long long)F * (N + 1) * (C / 4);
  if (total <= 0) return 0;
  VC_LAUNCH((vit_assemble_fwd_kernel), ew_grid(total, 256), 256, 0, cs(s), e, F, N, C, cls, pos, drop, dropout_threshold(drop.p),
                                                                  drop_scale(drop), x);
  return check_launch("vit_assemble_fwd_kernel");
}

int vit_assemble_bwd(const float* dx, int F, int N, int C, Drop drop, float* de, float* dcls, float* dpos, stream_t s) {
  if (C % 4 != 0) return set_error("vit_assemble_bwd: C % 4 != 0");
  if (F <= 0) return 0;
  const int f_per_block = 8;
  dim3 grid(cdiv((long long)(N + 1) * (C / 4), 128), cdiv(F, f_per_block));
  VC_LAUNCH((vit_assemble_bwd_kernel), grid, 128, 0, cs(s), dx, F, N, C, drop, dropout_threshold(drop.p), drop_scale(drop), de, dcls,
                                                   dpos, f_per_block);
  return check_launch("vit_assemble_bwd_kernel");
}

int act_dropout_bwd(const float* dy, int64_t lddy, int64_t M, int N, int act, const float* aux, int64_t ldaux,
                    const bf16_t* aux_hi, int64_t ldaux_hi, Drop drop, float* g, int64_t ldg, bf16_t* g_hi, bf16_t* g_lo,
                    int64_t ldg_split, float* colsum, stream_t s) {
  if (N % 4 != 0) return set_error("act_dropout_bwd: N % 4 != 0");
  if ((act == VC_ACT_GELU || act == VC_ACT_TANH) && aux == nullptr) return set_error("act_dropout_bwd: aux required");
  if (act == VC_ACT_RELU && aux_hi == nullptr) return set_error("act_dropout_bwd: aux_hi required for relu");
  if (M <= 0) return 0;
  // 8 rows per thread for the image-encoder sizes; decoder-sized problems (a few hundred rows) get one row per thread,
  // otherwise a handful of CTAs would each walk their rows one dependent load at a time
  int rows_per_block = 32;
  while (rows_per_block > ADB_RG && (long long)cdiv(N / 4, 128) * cdiv(M, rows_per_block) < 148) rows_per_block >>= 1;
  dim3 grid(cdiv(N / 4, 128), cdiv(M, rows_per_block));
  VC_LAUNCH((act_dropout_bwd_kernel), grid, dim3(128, ADB_RG), 0, cs(s), dy, lddy, M, N, act, aux, ldaux,
                                                  reinterpret_cast<const __nv_bfloat16*>(aux_hi), ldaux_hi, drop,
                                                  dropout_threshold(drop.p), drop_scale(drop), g, ldg,
                                                  reinterpret_cast<__nv_bfloat16*>(g_hi),
                                                  reinterpret_cast<__nv_bfloat16*>(g_lo), ldg_split, colsum, rows_per_block);
  return check_launch("act_dropout_bwd_kernel");
}

int row_reduce_mod(const float* x, int64_t ldx, int64_t M, int N, int div, int mod, float* out, stream_t s) {
  if (M <= 0 || N <= 0) return 0;
  if (div < 1 || mod < 1) return set_error("row_reduce_mod: div/mod must be >= 1");
  VC_LAUNCH((row_reduce_mod_kernel), ew_grid(M * N, 256), 256, 0, cs(s), x, ldx, M, N, div, mod, out);
  return check_launch("row_reduce_mod_kernel");
}

int broadcast_rows(const float* src, int64_t lds, int64_t M, int N, int div, float* dst, int64_t ldd, bf16_t* d_hi,
                   bf16_t* d_lo, int64_t ldd_split, stream_t s) {
  if (N % 4 != 0) return set_error("broadcast_rows: N % 4 != 0");
  if (M <= 0) return 0;
  VC_LAUNCH((broadcast_rows_kernel), ew_grid(M * N / 4, 256), 256, 0, cs(s), src, lds, M, N, div, dst, ldd,
                                                                   reinterpret_cast<__nv_bfloat16*>(d_hi),
                                                                   reinterpret_cast<__nv_bfloat16*>(d_lo), ldd_split);
  return check_launch("broadcast_rows_kernel");
}

int embed_action_fwd(const float* actions, int64_t R, int A, int H, const float* W, const float* b, const float* E, int T,
                     float* y, bf16_t* y_hi, bf16_t* y_lo, stream_t s) {
  if (R <= 0) return 0;
  VC_LAUNCH((embed_action_fwd_kernel), ew_grid(R * H, 256), 256, 0, cs(s), actions, R, A, H, W, b, E, T, y,
                                                                 reinterpret_cast<__nv_bfloat16*>(y_hi),
                                                                 reinterpret_cast<__nv_bfloat16*>(y_lo));
  return check_launch("embed_action_fwd_kernel");
}

int embed_action_bwd(const float* dy, const float* y, const float* actions, int64_t R, int A, int H, int T, float* dW,
                     float* db, float* dE, stream_t s) {
  if (A > 8) return set_error("embed_action_bwd: act_dim > 8 unsupported");
  if (R <= 0) return 0;
  const int rows_per_block = 16;
  dim3 grid(cdiv(H, 128), cdiv(R, rows_per_block));
  VC_LAUNCH((embed_action_bwd_kernel), grid, 128, 0, cs(s), dy, y, actions, R, A, H, T, dW, db, dE, rows_per_block);
  return check_launch("embed_action_bwd_kernel");
}

int head_small_fwd(const float* x, int64_t R, int H, const float* W, const float* b, int C, float* out, stream_t s) {
  if (C > HEAD_MAXC || H % 4 != 0) return set_error("head_small_fwd: C <= 8 and H % 4 == 0 required");
  if (R <= 0) return 0;
  VC_LAUNCH((head_small_fwd_kernel), cdiv(R, 4), 128, 0, cs(s), x, R, H, W, b, C, out);
  return check_launch("head_small_fwd_kernel");
}

int head_small_bwd(const float* dout, const float* x, int64_t R, int H, const float* W, int C, float* dx, int accumulate_dx,
                   float* dW, float* db, stream_t s) {
  if (C > HEAD_MAXC || H % 4 != 0) return set_error("head_small_bwd: C <= 8 and H % 4 == 0 required");
  if (R <= 0) return 0;
  const int rows_per_block = 16;
  dim3 grid(cdiv(H / 4, 128), cdiv(R, rows_per_block));
  VC_LAUNCH((head_small_bwd_kernel), grid, 128, 0, cs(s), dout, x, R, H, W, C, dx, accumulate_dx, dW, db, rows_per_block);
  return check_launch("head_small_bwd_kernel");
}

int linear_rows_fwd(const float* x, const bf16_t* x_hi, const bf16_t* x_lo, int64_t ldx, int M, const float* W, const float* bias, int N,
                    int K, int act, const float* residual, int64_t ld_res, float* out_f32, int64_t ldo, bf16_t* out_hi, bf16_t* out_lo,
                    int64_t ldo_split, stream_t s) {
  if (M <= 0 || N <= 0) return 0;
  if (M > LR_MAXM) return set_error("linear_rows_fwd: at most 16 rows");
  if (K <= 0 || K % 128 != 0 || ldx % 4 != 0) return set_error("linear_rows_fwd: K must be a multiple of 128 and ldx of 4");
  if ((x == nullptr) == (x_hi == nullptr)) return set_error("linear_rows_fwd: give x either as fp32 or as split-bf16");
  if (!W || (out_f32 == nullptr && out_hi == nullptr)) return set_error("linear_rows_fwd: null weight or no output");
  const int MM = M <= 8 ? 8 : 16;
  const size_t smem = sizeof(float) * (size_t)MM * K;
  if (smem > 200 * 1024) return set_error("linear_rows_fwd: K too large for the shared-memory row buffer");
  int grid = cdiv(N, LR_WARPS);
  if (grid > 148 * 4) grid = 148 * 4;
#define VC_LR(MMV)                                                                                                          \
  do {                                                                                                                      \
    static size_t configured = 0;                                                                                           \
    if (smem > 48 * 1024 && smem > configured) {                                                                            \
      cudaError_t e = cudaFuncSetAttribute(linear_rows_kernel<MMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) return set_error(cudaGetErrorString(e));                                                        \
      configured = 200 * 1024;                                                                                              \
    }                                                                                                                       \
    VC_LAUNCH((linear_rows_kernel<MMV>), grid, LR_WARPS * 32, smem, cs(s), x, reinterpret_cast<const __nv_bfloat16*>(x_hi),  \
              reinterpret_cast<const __nv_bfloat16*>(x_lo), (long long)ldx, M, W, bias, N, K, act, residual, (long long)ld_res, out_f32,        \
              (long long)ldo, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), (long long)ldo_split);         \
  } while (0)
  if (MM == 8) VC_LR(8); else VC_LR(16);
#undef VC_LR
  return check_launch("linear_rows_kernel");
}

int frames_u8_normalize(const uint8_t* src, int64_t n, float mean, float std, float* dst, stream_t s) {
  if (n <= 0) return 0;
  if (!src || !dst) return set_error("frames_u8_normalize: null pointer");
  if (!(std != 0.f)) return set_error("frames_u8_normalize: std must be non-zero");
  if ((reinterpret_cast<uintptr_t>(src) & 15u) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15u) != 0)
    return set_error("frames_u8_normalize: src and dst must be 16-byte aligned");
  VC_LAUNCH((frames_u8_normalize_kernel), ew_grid((n + 15) / 16, 256), 256, 0, cs(s), src, (long long)n, mean, std, dst);
  return check_launch("frames_u8_normalize_kernel");
}

int add_f32(const float* a, const float* b, float* out, int64_t n, stream_t s) {
  if (n <= 0) return 0;
  VC_LAUNCH((add_kernel), ew_grid(n, 256), 256, 0, cs(s), a, b, out, n);
  return check_launch("add_kernel");
}

int zero_f32(float* x, int64_t n, stream_t s) {
  if (n <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(x, 0, (size_t)n * sizeof(float), cs(s));
  return e == cudaSuccess ? 0 : set_error(cudaGetErrorString(e));
}

int dropout_mask_debug(Drop drop, int64_t n, float* out, stream_t s) {
  if (n <= 0) return 0;
  VC_LAUNCH((dropout_mask_kernel), ew_grid((n + 3) / 4, 256), 256, 0, cs(s), drop, dropout_threshold(drop.p), drop_scale(drop), n, out);
  return check_launch("dropout_mask_kernel");
}

}  // namespace vck
