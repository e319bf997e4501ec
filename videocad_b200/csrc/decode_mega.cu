// One position of the key/value-cached rollout (/root/reference/model/autoregressive_transformer.py:222-275) as ONE persistent
// kernel.  decode.cu runs the step as 66 dependent launches (6.7 us each inside a CUDA graph: the chain, not the 293 MB of weights,
// bounds it).  Here dec_mega_ctas() = 128 co-resident CTAs (cooperative launch, one per SM) walk the same phases -- the DecGemv /
// DecAttn / DecSelect descriptors of seq_decode_step_dev, uploaded once per rollout as a program -- and meet at a grid-wide barrier
// (one release-add + acquire-poll on a global counter) after each phase instead of at a kernel boundary:
//   * weights: every CTA owns the same output columns of a Linear for the whole rollout.  Their fp32 rows stream through a ring of
//     MG_SLOTS x 32 KB shared-memory slots, filled by TMA bulk copies (cp.async.bulk -> mbarrier) that warp 0 issues as soon as a slot
//     has been consumed -- up to MG_SLOTS chunks, i.e. two to four phases, AHEAD of the phase being computed, across the barriers
//     (weights never change during a rollout).  HBM latency and the barrier wait overlap; a phase starts with its weights in smem.
//   * activations ([B, H] rows, B <= 16) live in global memory (L2) between phases and are read with ld.global.cg: L1 is not
//     coherent across the SMs that produced them within one launch.
//   * the phase bodies are those of decode.cu (input rows built on load: plain / LayerNorm / action embedding; halving-butterfly
//     row reductions; split-key attention merged by the last CTA to arrive; device-side argmax + action mask + normalise).
// The position counter and the barrier epoch advance on the device, so one captured launch serves every position.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

constexpr int MG_CTAS = 128;
constexpr int MG_THREADS = 256;
constexpr int MG_WARPS = MG_THREADS / 32;
constexpr int MG_SLOTS = 5;
constexpr int MG_SLOT_BYTES = 32 * 1024;
constexpr int MG_MAX_COLS = 48;  // output columns of one CTA in one GEMV phase
constexpr int MG_PART_LD = MG_MAX_COLS + 1;  // partial sums [warp][row][column]: odd row pitch, conflict-free for the dot-product writes and the epilogue reads
constexpr int MG_PART_BYTES = ((MG_WARPS * 16 * MG_PART_LD * 4 + 127) / 128) * 128;  // also the attention / selection scratch (>= 8.1 KB)
constexpr float MG_LN_EPS = 1e-5f;
constexpr long long MG_TIMEOUT_CYCLES = 4000000000LL;  // ~2 s: a barrier that never completes poisons the output instead of hanging

inline cudaStream_t cs(stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// VC_MEGA_DEBUG: global-timer stamps of every CTA at every barrier of the traced position: [phase][cta][arrive, leave]
__device__ unsigned long long g_mega_trace[VC_MEGA_MAX_PHASES * MG_CTAS * 2];
__device__ int g_mega_trace_valid;
__device__ __forceinline__ unsigned long long mg_globaltimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

struct MegaSync {
  unsigned int* ctr;    // arrivals, monotonic over the whole rollout                          (8-byte aligned; err is the next word:
  unsigned int* err;    // set when a barrier timed out: every later barrier is skipped         the poll reads both at once)
  unsigned int* epoch;  // launches completed so far
  const int* t_ptr;
  int debug_t;          // VC_MEGA_DEBUG=<position>: CTAs 0 and 64 print where the cycles of that position went (experiment)
  int fenced;           // VC_MEGA_FENCED=1: acquire fence after the last poll of a barrier (experiment)
};

// The weight rows are read once per step and there are 293 MB of them: fetched with an L2 evict-first policy they do not push the
// step's re-used data (cached keys / values, LayerNorm parameters, biases, activations) out of the 126 MB L2.
__device__ __forceinline__ uint64_t mg_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void mg_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void mg_red_release(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// {p[0], p[1]} in one 8-byte load.  Relaxed: an acquire load costs an L1 invalidation (CCTL.IVALL) per poll; the barrier issues ONE
// acquire fence after the last poll instead.
__device__ __forceinline__ unsigned long long mg_ld_relaxed64(const unsigned int* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mg_fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int mg_atom_add_acq_rel(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Sum v[0..MM) over the 32 lanes with MM - 1 + log2(32 / MM) shuffles (see decode.cu): returns the total of row (lane >> log2(32 / MM)).
template <int MM>
__device__ __forceinline__ float mg_reduce_rows(float (&v)[MM], int lane) {
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

// ---- this CTA's share of a GEMV phase and the chunks it is streamed in
struct MgShare {
  int n0, cols, ch;  // first column, number of columns (0: none), columns per ring slot
};
__device__ __forceinline__ MgShare mg_share(const DecGemv& g, int cta) {
  MgShare s;
  s.n0 = cta * g.cols_per_cta;
  s.cols = max(0, min(g.cols_per_cta, g.N - s.n0));
  s.ch = (MG_SLOT_BYTES / (g.K * 4)) & ~3;
  return s;
}

// producer state of the weight ring (lives in the lanes of the LAST warp: all values are warp-uniform)
struct MgProducer {
  int ph;        // phase of the next chunk to request
  int c;         // first column (within this CTA's share) of that chunk
  unsigned int issued;
};

__device__ __forceinline__ void mg_issue_one(MgProducer& pr, const MegaPhase* prog, int nph, int cta, uint8_t* ring, uint64_t* full, int lane) {
  while (pr.ph < nph) {
    if (prog[pr.ph].kind == VC_MEGA_GEMV) {
      const MgShare sh = mg_share(prog[pr.ph].g, cta);
      if (pr.c < sh.cols) break;
    }
    ++pr.ph;
    pr.c = 0;
  }
  if (pr.ph >= nph) return;
  const DecGemv& g = prog[pr.ph].g;
  const MgShare sh = mg_share(g, cta);
  const int ncol = min(sh.ch, sh.cols - pr.c);
  const unsigned int slot = pr.issued % MG_SLOTS;
  const uint32_t row_bytes = (uint32_t)g.K * 4u;
  if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)ncol * row_bytes);
  __syncwarp();
  uint8_t* dst = ring + (size_t)slot * MG_SLOT_BYTES;
  const float* src = g.W + (size_t)(sh.n0 + pr.c) * g.K;
  const uint64_t policy = mg_policy_evict_first();
  for (int c = lane; c < ncol; c += 32) mg_bulk_g2s(dst + (size_t)c * row_bytes, src + (size_t)c * g.K, row_bytes, &full[slot], policy);
  pr.c += ncol;
  ++pr.issued;
}

struct MgSmem {
  uint64_t* full;    // [MG_SLOTS]
  MegaPhase* prog;   // [nph]
  float* part;       // GEMV: [MG_WARPS][MM][MG_PART_LD] partial sums; attention / selection scratch
  float* red;        // [128]
  float* stat;       // [32]
  unsigned int* flag;  // [4]: ticket broadcast, barrier failure
  uint8_t* ring;     // [MG_SLOTS][MG_SLOT_BYTES]
  long long* tl;     // [16] debug: cycles per category (thread 0), tl[15] = last time stamp
  bool dbg;
  bool trace;
};

// debug time line: charge the cycles since the previous stamp to category `cat`
enum { TL_XLOAD = 0, TL_RING = 1, TL_DOTS = 2, TL_EPI = 3, TL_BAR_SYNC = 4, TL_BAR_WAIT = 5, TL_ATTN = 6, TL_SELECT = 7, TL_SETUP = 8, TL_LN_ISSUE = 9, TL_LN_ROW = 10, TL_LN_STATS = 11, TL_LN_SYNC = 12 };
__device__ __forceinline__ void mg_stamp(const MgSmem& sm, int cat, int tid) {
  if (sm.dbg && tid == 0) {
    const long long now = clock64();
    sm.tl[cat] += now - sm.tl[15];
    sm.tl[15] = now;
  }
}

// ------------------------------------------------------------------------------------------------- grid-wide barrier
__device__ __forceinline__ void mg_grid_barrier(const MegaSync& sy, unsigned int target, MgSmem& sm, int tid, int ph) {
  __syncthreads();  // every thread's global writes of this phase precede thread 0's release below
  mg_stamp(sm, TL_BAR_SYNC, tid);
  if (tid == 0 && sm.flag[1] == 0u) {
    // bar.sync above orders the CTA's writes before this thread; the release is cumulative over them, the acquire below covers the reads
    // of every thread after the closing bar.sync
    if (sm.trace) g_mega_trace[((size_t)ph * MG_CTAS + blockIdx.x) * 2] = mg_globaltimer();
    mg_red_release(sy.ctr, 1u);
    const long long t0 = clock64();
    for (;;) {
      const unsigned long long v = mg_ld_relaxed64(sy.ctr);
      if ((unsigned int)(v >> 32) != 0u) {  // another CTA gave up: the step is lost, do not wait for it
        sm.flag[1] = 1u;
        break;
      }
      if ((int)((unsigned int)v - target) >= 0) break;
      if (clock64() - t0 > MG_TIMEOUT_CYCLES) {
        sm.flag[1] = 1u;
        atomicExch(sy.err, 1u);
        break;
      }
    }
    // No acquire fence after the last poll (VC_MEGA_FENCED=1 adds one): every load of data another CTA wrote in this launch is an
    // ld.global.cg issued after the bar.sync below -- i.e. after this observation -- and reads L2, where the producers' writes landed
    // before their release-add.  The fence would cost an L1 invalidation + MEMBAR.GPU (~0.4 us) per barrier for nothing L1 holds.
    if (sy.fenced) mg_fence_acq_rel();
    if (sm.trace) g_mega_trace[((size_t)ph * MG_CTAS + blockIdx.x) * 2 + 1] = mg_globaltimer();
  }
  __syncthreads();
  mg_stamp(sm, TL_BAR_WAIT, tid);
}

// ------------------------------------------------------------------------------------------------- GEMV phase
template <int MM>
__device__ __forceinline__ void mg_gemv_phase(const DecGemv& a, int cta, int t, MgSmem& sm, MgProducer& pr, unsigned int& consumed, int nph,
                                              int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const MgShare sh = mg_share(a, cta);
  if (sh.cols == 0) return;
  const int K = a.K, M = a.M, n0 = sh.n0, cols = sh.cols;
  float* part = sm.part;
  // ---------------------------------------------------------------- input rows: this thread's float4 slot of K, all rows
  // Thread <-> float4 slot of K, ROTATED by the CTA index: right after a grid barrier all 128 CTAs read the same [B, K] rows, and without
  // the rotation they would all ask L2 for the same lines at the same moment (measured: 5300 cycles for these 32 KB per phase).
  const int k0 = ((tid + cta * 2) & (MG_THREADS - 1)) * 4;
  const bool active = k0 < K;
  float4 x[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) x[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.in_mode == VC_DEC_IN_PLAIN || a.in_mode == VC_DEC_IN_LN) {
    const bool ln = a.in_mode == VC_DEC_IN_LN;
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      if (ln) {  // parameters: requested together with the rows
        g = __ldg(reinterpret_cast<const float4*>(a.gamma + k0));
        bt = __ldg(reinterpret_cast<const float4*>(a.beta + k0));
      }
#pragma unroll
      for (int m = 0; m < MM; ++m)
        if (m < M) x[m] = ldcg4(a.x + (size_t)m * a.ldx + k0);
    }
    if (ln) mg_stamp(sm, TL_LN_ISSUE, tid);
    if (ln) {
      // LayerNorm over K, two-pass statistics (biased variance, eps inside the square root) as ln_fwd_kernel -- computed per ROW by one
      // warp that reads the whole row itself (second read of 4 KB from L2, in flight together with the loads above): two warp
      // reductions and ONE CTA barrier instead of two CTA-wide reductions with four
      for (int m = warp; m < M; m += MG_WARPS) {
        float4 r[8];  // K <= 1024: 8 float4 per lane
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = lane * 4 + ((i + cta) & 7) * 128;  // chunk order rotated by the CTA index, like k0 above
          r[i] = k < K ? ldcg4(a.x + (size_t)m * a.ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          sum += r[i].x + r[i].y + r[i].z + r[i].w;
        }
        mg_stamp(sm, TL_LN_ROW, tid);
        const float mean = warp_sum(sum) / (float)K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (lane * 4 + ((i + cta) & 7) * 128 < K) {
            const float d0 = r[i].x - mean, d1 = r[i].y - mean, d2 = r[i].z - mean, d3 = r[i].w - mean;
            q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        }
        q = warp_sum(q);
        if (lane == 0) {
          sm.stat[m] = mean;
          sm.stat[16 + m] = rsqrtf(q / (float)K + MG_LN_EPS);
        }
      }
      mg_stamp(sm, TL_LN_STATS, tid);
      __syncthreads();
      mg_stamp(sm, TL_LN_SYNC, tid);
      if (active) {
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          if (m < M) {
            const float mean = sm.stat[m], rstd = sm.stat[16 + m];
            x[m].x = (x[m].x - mean) * rstd * g.x + bt.x;
            x[m].y = (x[m].y - mean) * rstd * g.y + bt.y;
            x[m].z = (x[m].z - mean) * rstd * g.z + bt.z;
            x[m].w = (x[m].w - mean) * rstd * g.w + bt.w;
          }
        }
      }
    }
  } else if (a.in_mode == VC_DEC_IN_EMBED) {
    // token of this position: tanh(embed_action(a_t) + E[t])  (autoregressive_transformer.py:110-113, 176-178).  The fed-back actions
    // were written by the previous launch's selection: one L2 read each into shared memory (read per use with ld.cg they cost an L2
    // round trip per multiply: 224 dependent round trips, measured 160 000 cycles)
    float* sact = sm.part;  // [M][act_dim]
    for (int i = tid; i < M * a.act_dim; i += MG_THREADS) sact[i] = __ldcg(a.actions + i);
    __syncthreads();
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
            for (int i = 0; i < a.act_dim; ++i) acc += sact[m * a.act_dim + i] * a.emb_W[(size_t)(k0 + j) * a.act_dim + i];
            v[j] = tanhf(acc + e4[j]);
          }
          x[m] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  if (a.in_mode == VC_DEC_IN_EMBED) __syncthreads();  // the staged actions share the partial-sum buffer of the dot products below
  if (a.x_out != nullptr && cta == 0 && active) {  // the rows just built are the residual operand of a later phase
#pragma unroll
    for (int m = 0; m < MM; ++m)
      if (m < M) *reinterpret_cast<float4*>(a.x_out + (size_t)m * K + k0) = x[m];
  }
  mg_stamp(sm, TL_XLOAD, tid);
  // ---------------------------------------------------------------- epilogue operands of the first round: request them now
  const int e_m = tid / cols, e_c = tid - e_m * cols;
  const bool e_on = tid < cols * M;
  float e_bias = 0.f, e_res = 0.f;
  if (e_on) {
    if (a.bias) e_bias = __ldg(a.bias + n0 + e_c);
    if (a.residual) e_res = __ldcg(a.residual + (size_t)e_m * a.ld_res + n0 + e_c);
  }
  // ---------------------------------------------------------------- dot products, one ring slot (sh.ch columns) at a time
  constexpr int kShift = MM == 8 ? 2 : 1;
  for (int c0 = 0; c0 < cols; c0 += sh.ch) {
    const int ncol = min(sh.ch, cols - c0);
    const unsigned int slot = consumed % MG_SLOTS;
    mbar_wait(&sm.full[slot], (consumed / MG_SLOTS) & 1u);
    mg_stamp(sm, TL_RING, tid);
    const float* wsm = reinterpret_cast<const float*>(sm.ring + (size_t)slot * MG_SLOT_BYTES);
    int c = 0;
    for (; c + 4 <= ncol; c += 4) {  // four columns per iteration: four independent load / FMA / shuffle chains
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
      const float r0 = mg_reduce_rows<MM>(p0, lane), r1 = mg_reduce_rows<MM>(p1, lane), r2 = mg_reduce_rows<MM>(p2, lane),
                  r3 = mg_reduce_rows<MM>(p3, lane);
      if ((lane & ((1 << kShift) - 1)) == 0) {
        float* dst = part + ((size_t)warp * MM + (lane >> kShift)) * MG_PART_LD + c0 + c;
        dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
      }
    }
    for (; c < ncol; ++c) {
      float p[MM];
      if (active) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsm + (size_t)c * K + k0);
#pragma unroll
        for (int m = 0; m < MM; ++m) p[m] = fmaf(w4.x, x[m].x, fmaf(w4.y, x[m].y, fmaf(w4.z, x[m].z, w4.w * x[m].w)));
      } else {
#pragma unroll
        for (int m = 0; m < MM; ++m) p[m] = 0.f;
      }
      const float r = mg_reduce_rows<MM>(p, lane);
      if ((lane & ((1 << kShift) - 1)) == 0) part[((size_t)warp * MM + (lane >> kShift)) * MG_PART_LD + c0 + c] = r;
    }
    __syncthreads();  // every warp is done with the slot (and its partial sums are visible)
    mg_stamp(sm, TL_DOTS, tid);
    ++consumed;
    // refill the slot just released: by the LAST warp, which has no part in the epilogue (issuing 8 bulk copies costs ~2000 cycles)
    if (warp == MG_WARPS - 1) mg_issue_one(pr, sm.prog, nph, cta, sm.ring, sm.full, lane);
  }
  // ---------------------------------------------------------------- epilogue: consecutive threads <-> consecutive columns
  for (int idx = tid; idx < cols * M; idx += MG_THREADS) {
    const int m = idx / cols, cc = idx - m * cols;
    const int n = n0 + cc;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < MG_WARPS; ++w) v += part[((size_t)w * MM + m) * MG_PART_LD + cc];
    const bool first = idx == tid;
    if (a.bias) v += first ? e_bias : __ldg(a.bias + n);
    v = apply_act(v, a.act);
    if (a.residual) v += first ? e_res : __ldcg(a.residual + (size_t)m * a.ld_res + n);
    a.out[(size_t)m * a.out_row_stride + (size_t)t * a.out_t_stride + n] = v;
  }
  mg_stamp(sm, TL_EPI, tid);
}

// ------------------------------------------------------------------------------------------------- attention phase
template <int DPL>
__device__ __forceinline__ void mg_ldcg_vec(const float* p, float (&v)[DPL]) {
  if constexpr (DPL == 2) {
    const float2 a = __ldcg(reinterpret_cast<const float2*>(p));
    v[0] = a.x; v[1] = a.y;
  } else {
#pragma unroll
    for (int i = 0; i < DPL / 4; ++i) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
  }
}

constexpr int MG_KEYS = 4;  // keys per warp iteration

template <int DPL>
__device__ __forceinline__ void mg_attn_phase(const DecAttn& a, int B, int cta, int ncta, int t, MgSmem& sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  float* sm_m = sm.part;                      // [MG_WARPS]
  float* sm_l = sm.part + MG_WARPS;           // [MG_WARPS]
  float* sm_acc = sm.part + 2 * MG_WARPS;     // [MG_WARPS][32 * DPL]
  const int items = B * a.nh * a.nsplit;
  const int j_lo = a.window > 0 ? max(0, t - a.window + 1) : 0;
  const int nkeys = t - j_lo + 1;
  for (int item = cta; item < items; item += ncta) {
    const int s = item % a.nsplit;
    const int bh = item / a.nsplit;
    const int h = bh % a.nh, b = bh / a.nh;
    const int d0 = h * a.dh + lane * DPL;
    float q[DPL];
    mg_ldcg_vec<DPL>(a.q + (size_t)b * a.q_bstride + (size_t)t * a.q_tstride + d0, q);
#pragma unroll
    for (int i = 0; i < DPL; ++i) q[i] *= a.scale;
    float mrun = -INFINITY, lrun = 0.f, acc[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
    const float* kb = a.k + (size_t)b * a.kv_bstride + d0;
    const float* vb = a.v + (size_t)b * a.kv_bstride + d0;
    const int stride = a.nsplit * MG_WARPS;
    for (int i = s * MG_WARPS + warp; i < nkeys; i += MG_KEYS * stride) {
      float kk[MG_KEYS][DPL], vv[MG_KEYS][DPL], sc[MG_KEYS];
#pragma unroll
      for (int u = 0; u < MG_KEYS; ++u) {
        const int iu = i + u * stride;
        const size_t row = (size_t)(j_lo + (iu < nkeys ? iu : i)) * a.kv_rstride;  // past the end: re-read a valid row, weight 0
        mg_ldcg_vec<DPL>(kb + row, kk[u]);
        mg_ldcg_vec<DPL>(vb + row, vv[u]);
      }
      float mnew = mrun;
#pragma unroll
      for (int u = 0; u < MG_KEYS; ++u) {
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
      for (int u = 0; u < MG_KEYS; ++u) {
        const float p = __expf(sc[u] - mnew);  // exp(-inf) = 0 for the keys past the end
        lrun += p;
#pragma unroll
        for (int e = 0; e < DPL; ++e) acc[e] = fmaf(p, vv[u][e], acc[e]);
      }
      mrun = mnew;
    }
    if (lane == 0) { sm_m[warp] = mrun; sm_l[warp] = lrun; }
#pragma unroll
    for (int e = 0; e < DPL; ++e) sm_acc[warp * 32 * DPL + lane * DPL + e] = acc[e];
    __syncthreads();
    // merge the 8 warps of this CTA
    float mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < MG_WARPS; ++w) mx = fmaxf(mx, sm_m[w]);
    float o = 0.f, l = 0.f;
    if (mx != -INFINITY) {
#pragma unroll
      for (int w = 0; w < MG_WARPS; ++w) {
        if (sm_m[w] != -INFINITY) {
          const float e = __expf(sm_m[w] - mx);
          l += e * sm_l[w];
          if (tid < a.dh) o += e * sm_acc[w * 32 * DPL + tid];
        }
      }
    }
    float* outp = a.out + (size_t)b * a.ld_out + h * a.dh;
    if (a.nsplit == 1) {  // a single part: finished
      if (tid < a.dh) outp[tid] = o / l;
    } else {
      const size_t pbase = ((size_t)b * a.nh + h) * a.nsplit;
      if (tid < a.dh) a.part_o[(pbase + s) * a.dh + tid] = o;
      if (tid == 0) { a.part_ml[(pbase + s) * 2] = mx; a.part_ml[(pbase + s) * 2 + 1] = l; }
      // the last part of this (sequence, head) to arrive merges all of them (bar.sync + acq_rel ticket: the CTA's partials are
      // released, the earlier arrivals' partials acquired)
      __syncthreads();
      if (tid == 0) sm.flag[0] = mg_atom_add_acq_rel(a.counters + (size_t)b * a.nh + h, 1u);
      __syncthreads();
      if (sm.flag[0] == (unsigned int)(a.nsplit - 1)) {
        if (tid == 0) a.counters[(size_t)b * a.nh + h] = 0u;  // ready for the next position
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
    }
    __syncthreads();  // the shared scratch is reused by this CTA's next item
  }
}

// ------------------------------------------------------------------------------------------------- heads' argmax + feedback
// command -> which of the 6 parameters it carries (autoregressive_transformer.py:83-89)
__constant__ int kMegaActionMask[5][6] = {{1, 1, 0, 0, 0, 0}, {0, 0, 1, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 0}};

__device__ __forceinline__ void mg_select_phase(const DecSelect& a, int B, int cta, int t, const MegaSync& sy, unsigned int epoch, MgSmem& sm,
                                                int tid) {
  if (cta >= B) return;
  const int lane = tid & 31, warp = tid >> 5;
  float* ds_x = sm.part;  // [H] normalised row
  float* red = sm.red;    // [MG_WARPS]
  float* bcast = sm.stat;
  float* logit_c = sm.stat + 1;                       // [8]
  int* arg_p = reinterpret_cast<int*>(sm.red + 32);   // [8]
  const int b = cta, H = a.H;
  const float* y = a.y + (size_t)b * H;
  float s = 0.f;
  for (int k = tid; k < H; k += MG_THREADS) s += __ldcg(y + k);
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < MG_WARPS; ++w) v += red[w]; bcast[0] = v / (float)H; }
  __syncthreads();
  const float mean = bcast[0];
  float qv = 0.f;
  for (int k = tid; k < H; k += MG_THREADS) { const float d = __ldcg(y + k) - mean; qv += d * d; }
  qv = warp_sum(qv);
  __syncthreads();
  if (lane == 0) red[warp] = qv;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < MG_WARPS; ++w) v += red[w]; bcast[0] = rsqrtf(v / (float)H + MG_LN_EPS); }
  __syncthreads();
  const float rstd = bcast[0];
  for (int k = tid; k < H; k += MG_THREADS) ds_x[k] = (__ldcg(y + k) - mean) * rstd * a.gamma[k] + a.beta[k];
  __syncthreads();
  // command head: warp w <-> class w
  if (warp < a.NC) {
    const float* wr = a.Wc + (size_t)warp * H;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(wr[k], ds_x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + a.bc[warp];
      if (sm.flag[1] != 0u) v = __int_as_float(0x7fc00000);  // a grid barrier timed out: the rollout is invalid, make it visible
      logit_c[warp] = v;
      a.cmds_all[((size_t)b * a.T + t) * a.NC + warp] = v;
    }
  }
  // parameter heads: warp i <-> parameter i, argmax over its NV logits (written by the head phase); first index on ties
  if (warp < a.NPAR) {
    const float* z = a.params_all + ((size_t)b * a.T + t) * ((size_t)a.NPAR * a.NV) + (size_t)warp * a.NV;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < a.NV; k += 32) {
      const float v = __ldcg(z + k);
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
    // apply_action_mask (autoregressive_transformer.py:91-108) then normalize_actions (:115-118)
    float par[8];
    for (int i = 0; i < a.NPAR; ++i) par[i] = (cmd < 5 && i < 6 && kMegaActionMask[cmd][i]) ? (float)arg_p[i] : -1.0f;
    if (a.NPAR > 3 && !(par[2] >= 200.f && par[2] < 250.f)) par[3] = -1.0f;
    float* out = a.action_next + (size_t)b * (1 + a.NPAR);
    out[0] = (float)cmd / 4.0f;
    for (int i = 0; i < a.NPAR; ++i) out[1 + i] = par[i] / 1000.0f;
  }
  // the last of the B selecting CTAs advances the position counter and the barrier epoch (every CTA read both at kernel entry)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(a.done_ctr, 1u);
    if (done == (unsigned int)B - 1u) {
      *a.done_ctr = 0u;
      *a.t_ptr = t + 1;
      *sy.epoch = epoch + 1u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------- the kernel
template <int MM, int DPL>
__global__ void __launch_bounds__(MG_THREADS, 1) dec_mega_kernel(const MegaPhase* prog_g, int nph, const MegaSync sy) {
  extern __shared__ __align__(128) uint8_t mg_smem[];
  MgSmem sm;
  sm.full = reinterpret_cast<uint64_t*>(mg_smem);                     // 64 B reserved
  sm.flag = reinterpret_cast<unsigned int*>(mg_smem + 64);            // 16 B
  sm.stat = reinterpret_cast<float*>(mg_smem + 128);                  // 32 floats: row means | row rstd (selection: broadcast + command logits)
  sm.red = reinterpret_cast<float*>(mg_smem + 256);                   // 128 floats (selection: per-warp sums at 0..7, argmax indices at 32..39)
  sm.part = reinterpret_cast<float*>(mg_smem + 1024);                 // MG_MAX_COLS * MG_WARPS * 16 floats = 24 KB
  sm.prog = reinterpret_cast<MegaPhase*>(mg_smem + 1024 + MG_PART_BYTES);
  const size_t prog_bytes = ((size_t)VC_MEGA_MAX_PHASES * sizeof(MegaPhase) + 127) & ~(size_t)127;
  sm.ring = mg_smem + 1024 + MG_PART_BYTES + prog_bytes;
  sm.tl = reinterpret_cast<long long*>(mg_smem + 768);                // 128 B
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const long long t_entry = clock64();

  // the program -> shared memory (16-byte words)
  {
    const uint4* src = reinterpret_cast<const uint4*>(prog_g);
    uint4* dst = reinterpret_cast<uint4*>(sm.prog);
    const int words = nph * (int)(sizeof(MegaPhase) / 16);
    for (int i = tid; i < words; i += MG_THREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < MG_SLOTS; ++i) mbar_init(&sm.full[i], 1);
    fence_barrier_init();
    sm.flag[0] = 0u;
    sm.flag[1] = *sy.err;  // sticky across the launches of a rollout
  }
  const int t = *sy.t_ptr;
  const unsigned int epoch = *sy.epoch;
  sm.dbg = sy.debug_t >= 0 && t == sy.debug_t && (cta == 0 || cta == 64);
  sm.trace = sy.debug_t >= 0 && t == sy.debug_t;
  if (sm.trace && cta == 0 && tid == 0) g_mega_trace_valid = nph;
  if (sm.dbg && tid == 0) {
    for (int i = 0; i < 15; ++i) sm.tl[i] = 0;
    sm.tl[15] = t_entry;
  }
  __syncthreads();

  MgProducer pr;
  pr.ph = 0; pr.c = 0; pr.issued = 0u;
  if (warp == MG_WARPS - 1) {
    for (int i = 0; i < MG_SLOTS; ++i) mg_issue_one(pr, sm.prog, nph, cta, sm.ring, sm.full, lane);
  }
  unsigned int consumed = 0u;
  mg_stamp(sm, TL_SETUP, tid);
  // barriers of this launch: one after every phase but the last
  unsigned int target = epoch * (unsigned int)(nph - 1) * (unsigned int)ncta;
  for (int ph = 0; ph < nph; ++ph) {
    const MegaPhase& P = sm.prog[ph];
    if (P.kind == VC_MEGA_GEMV) {
      mg_gemv_phase<MM>(P.g, cta, t, sm, pr, consumed, nph, tid);
    } else if (P.kind == VC_MEGA_ATTN) {
      mg_attn_phase<DPL>(P.a, P.B, cta, ncta, t, sm, tid);
      mg_stamp(sm, TL_ATTN, tid);
    } else {
      mg_select_phase(P.s, P.B, cta, t, sy, epoch, sm, tid);
      mg_stamp(sm, TL_SELECT, tid);
    }
    if (ph + 1 < nph) {
      target += (unsigned int)ncta;
      mg_grid_barrier(sy, target, sm, tid, ph);
    }
  }
  if (sm.dbg && tid == 0)
    printf("mega timeline t=%d cta=%d (cycles): setup %lld | x rows %lld | ring wait %lld | dots %lld | epilogue %lld | barrier: cta sync %lld, "
           "grid wait %lld | attention %lld | select %lld | total %lld || LN: issue %lld, row landed %lld, stats %lld, cta sync %lld\n",
           t, cta, sm.tl[TL_SETUP], sm.tl[TL_XLOAD], sm.tl[TL_RING], sm.tl[TL_DOTS], sm.tl[TL_EPI], sm.tl[TL_BAR_SYNC], sm.tl[TL_BAR_WAIT],
           sm.tl[TL_ATTN], sm.tl[TL_SELECT], clock64() - t_entry, sm.tl[TL_LN_ISSUE], sm.tl[TL_LN_ROW], sm.tl[TL_LN_STATS], sm.tl[TL_LN_SYNC]);
}

// VC_MEGA_DEBUG (experiment): after every launch look whether the traced position ran; print per phase how the 128 CTAs reached and left
// its barrier.  Synchronises the stream: not for timed runs, and not under graph capture.
void mega_trace_dump(cudaStream_t st) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
  if (cudaStreamSynchronize(st) != cudaSuccess) return;
  int nph = 0;
  if (cudaMemcpyFromSymbol(&nph, g_mega_trace_valid, sizeof(int)) != cudaSuccess || nph <= 0) return;
  static std::vector<unsigned long long> h(VC_MEGA_MAX_PHASES * MG_CTAS * 2);
  if (cudaMemcpyFromSymbol(h.data(), g_mega_trace, h.size() * sizeof(unsigned long long)) != cudaSuccess) return;
  const int zero = 0;
  cudaMemcpyToSymbol(g_mega_trace_valid, &zero, sizeof(int));
  unsigned long long prev_leave = 0;
  printf("mega trace (ns): phase | work of the fastest / median / slowest CTA since the previous barrier opened | slowest CTA | barrier: last arrival -> last leave\n");
  for (int ph = 0; ph + 1 < nph; ++ph) {
    std::vector<unsigned long long> arr(MG_CTAS);
    unsigned long long last_arr = 0, last_leave = 0;
    int slow = 0;
    for (int c = 0; c < MG_CTAS; ++c) {
      arr[c] = h[((size_t)ph * MG_CTAS + c) * 2];
      if (arr[c] > last_arr) { last_arr = arr[c]; slow = c; }
      last_leave = std::max(last_leave, h[((size_t)ph * MG_CTAS + c) * 2 + 1]);
    }
    std::sort(arr.begin(), arr.end());
    if (ph > 0)
      printf("  %2d | %6lld %6lld %6lld | cta %3d | %5lld\n", ph, (long long)(arr[0] - prev_leave), (long long)(arr[MG_CTAS / 2] - prev_leave),
             (long long)(last_arr - prev_leave), slow, (long long)(last_leave - last_arr));
    prev_leave = last_leave;
  }
  fflush(stdout);
}

template <int MM, int DPL>
int launch_mega(const void* program_dev, int nph, unsigned int* sync, const int* t_ptr, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dec_mega_kernel<MM, DPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    int per_sm = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dec_mega_kernel<MM, DPL>, MG_THREADS, smem);
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    if (per_sm < 1 || per_sm * sms < MG_CTAS) return set_error("dec_mega: the persistent grid does not fit on this device");
    configured = true;
  }
  MegaSync sy;
  sy.ctr = sync; sy.err = sync + 1; sy.epoch = sync + 2; sy.t_ptr = t_ptr;
  static const int debug_t = getenv("VC_MEGA_DEBUG") ? atoi(getenv("VC_MEGA_DEBUG")) : -1;
  sy.debug_t = debug_t;
  static const int fenced = getenv("VC_MEGA_FENCED") ? atoi(getenv("VC_MEGA_FENCED")) : 0;
  sy.fenced = fenced;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(MG_CTAS);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barrier cannot wait for a CTA that was never scheduled
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, dec_mega_kernel<MM, DPL>, reinterpret_cast<const MegaPhase*>(program_dev), nph, sy);
  const int rc = check_launch("dec_mega_kernel");
  if (debug_t >= 0 && rc == 0) mega_trace_dump(st);
  return rc;
}

}  // namespace

int dec_mega_ctas() { return MG_CTAS; }

int dec_mega_cols(int N) {
  int cols = (N + MG_CTAS - 1) / MG_CTAS;
  cols = (cols + 3) / 4 * 4;
  return cols <= MG_MAX_COLS ? cols : 0;
}

int dec_mega_nsplit(int B, int nh) {
  int n = MG_CTAS / (B * nh);
  if (n < 1) n = 1;
  if (n > 8) n = 8;
  return n;
}

int dec_mega_upload(const MegaPhase* host, int nph, void* program_dev, stream_t s) {
  if (!host || !program_dev || nph < 1 || nph > VC_MEGA_MAX_PHASES) return set_error("dec_mega_upload: bad program");
  static_assert(sizeof(MegaPhase) % 16 == 0, "the kernel copies the program in 16-byte words");
  // pageable source: the runtime stages the bytes before the call returns, so the caller's array may go away
  cudaError_t e = cudaMemcpyAsync(program_dev, host, (size_t)nph * sizeof(MegaPhase), cudaMemcpyHostToDevice, cs(s));
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

int dec_mega(const void* program_dev, int nph, int M, int dh, unsigned int* sync, const int* t_ptr, stream_t s) {
  if (!program_dev || !sync || !t_ptr || nph < 2 || nph > VC_MEGA_MAX_PHASES) return set_error("dec_mega: bad arguments");
  if ((reinterpret_cast<uintptr_t>(sync) & 7) != 0) return set_error("dec_mega: sync words must be 8-byte aligned");
  if (M < 1 || M > 16) return set_error("dec_mega: 1..16 sequences");
  if (dh != 64 && dh != 128 && dh != 256) return set_error("dec_mega: head dim must be 64, 128 or 256");
  if ((reinterpret_cast<uintptr_t>(program_dev) & 15) != 0) return set_error("dec_mega: program must be 16-byte aligned");
  const size_t prog_bytes = ((size_t)VC_MEGA_MAX_PHASES * sizeof(MegaPhase) + 127) & ~(size_t)127;
  const size_t smem = 1024 + (size_t)MG_PART_BYTES + prog_bytes + (size_t)MG_SLOTS * MG_SLOT_BYTES;
  cudaStream_t st = cs(s);
  const bool m8 = M <= 8;
  if (dh == 64) return m8 ? launch_mega<8, 2>(program_dev, nph, sync, t_ptr, smem, st) : launch_mega<16, 2>(program_dev, nph, sync, t_ptr, smem, st);
  if (dh == 128) return m8 ? launch_mega<8, 4>(program_dev, nph, sync, t_ptr, smem, st) : launch_mega<16, 4>(program_dev, nph, sync, t_ptr, smem, st);
  return m8 ? launch_mega<8, 8>(program_dev, nph, sync, t_ptr, smem, st) : launch_mega<16, 8>(program_dev, nph, sync, t_ptr, smem, st);
}

}  // namespace vck
