// Split-bf16 tensor-core GEMM for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma (kind::f16,
// bf16 operands, fp32 accumulators in TMEM) -> tcgen05.ld epilogue with fused bias / row-add / activation /
// dropout / residual and fp32 + split-bf16 outputs.
//
// Replaces the cuBLAS calls behind every nn.Linear on the reference hot path
// (vit_pytorch Attention.to_qkv / to_out / FeedForward.net, nn.MultiheadAttention in/out_proj,
// TransformerDecoderLayer.linear1/2, embed_state / embed_image / image_projection / predict_action_class_0_999;
// /root/reference/model/autoregressive_transformer.py:54-65,76, base_transformer.py:53-54) and their
// autograd dgrad / wgrad GEMMs.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer        (one elected thread)
//   warp 1      tcgen05.mma issuer  (one elected thread); 3 MMAs per k16 step in fp32-grade mode
//   warps 2..5  epilogue            (TMEM lane quarter = warp % 4)
//   smem ring of STAGES x {A_hi, A_lo, B_hi, B_lo} tiles, TMEM double-buffered accumulators (2 x BN columns)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <functional>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"
#include "gemm_profile.h"

namespace vck {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int STG_LD = 36;  // floats per staged epilogue row (144 B): 16-byte aligned, conflict-free for 128-bit accesses

struct GemmParams {
  int M, N, K, passes, splitk, kb_per_split, num_kb;
  int a_mn, b_mn;
  const float* bias;
  const float* rowadd; long long ld_rowadd; int rowadd_div, rowadd_mod;
  float* preact; long long ld_preact;
  int act;
  float drop_scale; uint32_t drop_thresh; uint32_t drop_site; uint64_t drop_seed; const uint64_t* drop_seed_ptr; int drop_on;
  const float* residual; long long ld_res;
  float* out_f32; long long ldo;
  __nv_bfloat16 *out_hi, *out_lo; long long ldo_split;
  int act_backward;
  const float* act_aux; long long ld_act_aux;
  const __nv_bfloat16* act_aux_hi; long long ld_act_aux_hi;
  float* colsum;
  int debug;  // VC_GEMM_DEBUG (profiling experiments only): 1 = no global stores, 2 = TMEM loads only, 4 = no TMEM loads
};

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGES = (BN <= 128) ? 3 : 2;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;  // per-warp 32x32 fp32 transpose buffers of the epilogue
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two
};

// shared-memory matrix descriptor (SWIZZLE_128B, version 1).  K-major: SBO = 1024 B between 8-row groups.
// MN-major: LBO = 8192 B between 64-element MN groups (one TMA box each), SBO = 1024 B between 8-k groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int mn_major) {
  const uint32_t lbo = mn_major ? 8192u : 16u;
  const uint32_t sbo = 1024u;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Epilogue feature bits: the kernel is instantiated for a handful of feature sets so that each instance carries only the
// code it needs (the all-features body, unrolled, overflowed the instruction cache: ncu showed the epilogue warps in
// stall_no_inst).  A variant may be used for any launch whose required features are a SUBSET of its mask (each feature
// still checks its runtime pointer/flag).
enum : uint32_t {
  EF_BIAS = 1u << 0,     // per-column bias
  EF_PREACT = 1u << 1,
  EF_ACT = 1u << 2,      // forward activation
  EF_DROP = 1u << 3,
  EF_ACTBWD = 1u << 4,   // backward activation (act_backward)
  EF_RES = 1u << 5,
  EF_F32 = 1u << 6,      // fp32 output (plain store)
  EF_ATOMIC = 1u << 7,   // fp32 output accumulated with atomics (split-K)
  EF_SPLIT = 1u << 8,    // split-bf16 output
  EF_COLSUM = 1u << 9,
  EF_ROWADD = 1u << 10,  // row-broadcast add (timestep embedding rows)
  EF_ALL = (1u << 11) - 1
};

// Epilogue math on 4 consecutive columns of one output row, in the COALESCED layout (8 lanes cover 128 B of a row).
// `pre_res` / `pre_aux`: the residual / activation-backward operand of this quad when the caller has already loaded them
// (the pair kernel issues those global loads before it waits for TMEM, see below); otherwise they are loaded here.
// `dkey`: the dropout key of this launch, derived ONCE per thread by the caller (epilogue_seed) - a per-quad read of the
// device-resident seed word sat behind the preceding global stores (possible aliasing) and cost a load latency per quad.
template <uint32_t F>
__device__ __forceinline__ void epilogue_quad(const GemmParams& p, float (&v)[4], long long row, int col, bool add_bias,
                                              const DropKey& dkey,
                                              bool use_pre_res = false, float4 pre_res = float4{0.f, 0.f, 0.f, 0.f},
                                              bool use_pre_aux = false, float4 pre_aux = float4{0.f, 0.f, 0.f, 0.f},
                                              bool use_pre_relu = false, uint2 pre_relu = uint2{0u, 0u},
                                              bool use_pre_bias = false, float4 pre_bias = float4{0.f, 0.f, 0.f, 0.f}) {
  if constexpr ((F & EF_BIAS) != 0) {
    if (p.bias != nullptr && add_bias) {
      const float4 b = use_pre_bias ? pre_bias : __ldg(reinterpret_cast<const float4*>(p.bias + col));
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
  }
  if constexpr ((F & EF_ROWADD) != 0) {
    if (p.rowadd != nullptr && add_bias) {
      const int rr = ((int)row / p.rowadd_div) % p.rowadd_mod;
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowadd + (long long)rr * p.ld_rowadd + col));
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
  }
  // ACT_GELU_DSTORE: `preact` receives gelu'(v) * dropout mask * dropout scale (what the backward multiplies by) instead of v.
  // Two elements at a time, each pair stored as soon as it is finished: the kernel is register-capped (96), four more live
  // values per quad spilled.
  constexpr bool kDStore = (F & EF_PREACT) != 0 && (F & EF_ACT) != 0;
  bool dstore = false;
  if constexpr (kDStore) dstore = p.preact != nullptr && p.act == ACT_GELU_DSTORE && !p.act_backward;
  if (kDStore && dstore) {
    uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    float scale = 1.0f;
    uint32_t thresh = 0u;
    if constexpr ((F & EF_DROP) != 0) {
      if (p.drop_on) {
        const unsigned long long idx = (unsigned long long)row * (unsigned long long)p.N + (unsigned long long)col;
        const Rand4 r4 = dropout_words(dkey, idx >> 2);
        w[0] = r4.v[0]; w[1] = r4.v[1]; w[2] = r4.v[2]; w[3] = r4.v[3];
        scale = p.drop_scale;
        thresh = p.drop_thresh;
      }
    }
    float* dptr = p.preact + row * p.ld_preact + col;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float d0, d1;
      gelu_with_grad(v[2 * h], v[2 * h], d0);
      gelu_with_grad(v[2 * h + 1], v[2 * h + 1], d1);
      const bool k0 = w[2 * h] >= thresh, k1 = w[2 * h + 1] >= thresh;
      v[2 * h] = k0 ? v[2 * h] * scale : 0.0f;
      v[2 * h + 1] = k1 ? v[2 * h + 1] * scale : 0.0f;
      *reinterpret_cast<float2*>(dptr + 2 * h) = make_float2(k0 ? d0 * scale : 0.0f, k1 ? d1 * scale : 0.0f);
    }
  } else {
    if constexpr ((F & EF_PREACT) != 0) {
      if (p.preact != nullptr)
        *reinterpret_cast<float4*>(p.preact + row * p.ld_preact + col) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if constexpr ((F & EF_ACT) != 0) {
      if (p.act != ACT_NONE && !p.act_backward) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = apply_act(v[i], p.act);
      }
    }
    if constexpr ((F & EF_DROP) != 0) {
      if (p.drop_on) {
        const unsigned long long idx = (unsigned long long)row * (unsigned long long)p.N + (unsigned long long)col;
        const Rand4 w = dropout_words(dkey, idx >> 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (w.v[i] >= p.drop_thresh) ? v[i] * p.drop_scale : 0.0f;
      }
    }
  }
  if constexpr ((F & EF_ACTBWD) != 0) {
    if (p.act_backward) {
      if (p.act == ACT_GELU) {
        const float4 a = use_pre_aux ? pre_aux : *reinterpret_cast<const float4*>(p.act_aux + row * p.ld_act_aux + col);
        v[0] *= gelu_grad_f(a.x); v[1] *= gelu_grad_f(a.y); v[2] *= gelu_grad_f(a.z); v[3] *= gelu_grad_f(a.w);
      } else if (p.act == ACT_TANH) {
        const float4 a = use_pre_aux ? pre_aux : *reinterpret_cast<const float4*>(p.act_aux + row * p.ld_act_aux + col);
        v[0] *= 1.f - a.x * a.x; v[1] *= 1.f - a.y * a.y; v[2] *= 1.f - a.z * a.z; v[3] *= 1.f - a.w * a.w;
      } else if (p.act == ACT_MUL_AUX) {  // the forward stored the whole factor (ACT_GELU_DSTORE)
        const float4 a = use_pre_aux ? pre_aux : *reinterpret_cast<const float4*>(p.act_aux + row * p.ld_act_aux + col);
        v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; v[3] *= a.w;
      } else if (p.act == ACT_RELU) {
        const uint2 a = use_pre_relu ? pre_relu : *reinterpret_cast<const uint2*>(p.act_aux_hi + row * p.ld_act_aux_hi + col);
        if ((a.x & 0x7fffu) == 0u) v[0] = 0.f;
        if ((a.x & 0x7fff0000u) == 0u) v[1] = 0.f;
        if ((a.y & 0x7fffu) == 0u) v[2] = 0.f;
        if ((a.y & 0x7fff0000u) == 0u) v[3] = 0.f;
      }
    }
  }
  if constexpr ((F & EF_RES) != 0) {
    if (p.residual != nullptr) {
      const float4 b = use_pre_res ? pre_res : *reinterpret_cast<const float4*>(p.residual + row * p.ld_res + col);
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
  }
  if constexpr ((F & (EF_F32 | EF_ATOMIC)) != 0) {
    if (p.out_f32 != nullptr) {
      float* o = p.out_f32 + row * p.ldo + col;
      if constexpr ((F & EF_ATOMIC) != 0) {
        if (p.splitk > 1) {
          // one 128-bit vector reduction instead of four scalar atomics (sm_90+): the split-K weight-gradient GEMMs issue
          // 16 K reductions per 128x128 tile and split
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                       "f"(v[3])
                       : "memory");
        } else {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  if constexpr ((F & EF_SPLIT) != 0) {
    if (p.out_hi != nullptr) {
      uint2 hi, lo;
      split4(v, hi, lo);
      *reinterpret_cast<uint2*>(p.out_hi + row * p.ldo_split + col) = hi;
      if (p.out_lo != nullptr) *reinterpret_cast<uint2*>(p.out_lo + row * p.ldo_split + col) = lo;
    }
  }
}

// element o*U + j of an 8-entry register array with a compile-time j (U = unroll factor of the store loop): static indices
// plus selects, so that the array never becomes addressable (local memory)
template <int U, class T, int N>
__device__ __forceinline__ T epi_pick(const T (&arr)[N], int o, int j) {
  if constexpr (N < 8) {
    return arr[0];
  } else if constexpr (U >= 8) {
    return arr[j];
  } else if constexpr (U == 4) {
    return o == 0 ? arr[j] : arr[4 + j];
  } else {
    T r = arr[j];
    if (o == 1) r = arr[2 + j];
    if (o == 2) r = arr[4 + j];
    if (o == 3) r = arr[6 + j];
    return r;
  }
}

template <uint32_t F>
__device__ __forceinline__ DropKey epilogue_seed(const GemmParams& p) {
  if constexpr ((F & EF_DROP) != 0) {
    if (p.drop_on) return drop_key(p.drop_seed_ptr != nullptr ? *p.drop_seed_ptr : p.drop_seed, p.drop_site);
  }
  return DropKey{0u, 0u};
}

template <int BN, uint32_t F>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const GemmParams p) {
  pdl_trigger();  // the wait follows the barrier / TMEM set-up, which touches no global memory
  using C = Cfg<BN>;
  __shared__ long long tl[12];  // VC_GEMM_DEBUG & 8: clock64 timeline of CTA 0 (experiment)
  const bool tlon = (p.debug & 8) && blockIdx.x == 0;
  if (tlon && threadIdx.x == 0) tl[0] = clock64();
  extern __shared__ uint8_t smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_off = ((raw_addr + C::BAR_BYTES + 1023u) & ~1023u) - raw_addr;
  uint8_t* tiles = smem_raw + tiles_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (p.passes == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tlon && threadIdx.x == 0) tl[1] = clock64();
  pdl_wait();
  if (tlon && threadIdx.x == 0) tl[2] = clock64();

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int total_tiles = num_m * num_n * p.splitk;
  const uint32_t tx_bytes = (uint32_t)((p.passes == 3 ? 2 : 1) * (A_TILE_BYTES + C::B_TILE_BYTES));

  // Producer and MMA warps: ALL 32 lanes walk the (warp-uniform) loops and wait on the barriers; one elected lane issues the TMA /
  // tcgen05 instructions.  Running the loops under `if (lane == 0)` made the compiler treat every operand as per-thread data:
  // it wrapped each tcgen05.mma in an ELECT / 5x R2UR / BRA.U.ANY loop, ~84 cycles of issue per MMA -- more than the 33 (64-wide
  // tile) or 65 cycles (128-wide) the tensor pipe needs for one, so the single-CTA kernel's main loop was issue-bound
  // (clock64 timeline: 1000 cycles per k-block of 12 MMAs, profiles/r01n_gemm_timeline.txt).
  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_idx = tile % num_n;
        const int t2 = tile / num_n;
        const int m_idx = t2 % num_m;
        const int split = t2 / num_m;
        const int m0 = m_idx * BM, n0 = n_idx * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (!elect_one()) continue;
          uint8_t* st = tiles + s * C::STAGE_BYTES;
          uint8_t* sA_hi = st;
          uint8_t* sA_lo = st + A_TILE_BYTES;
          uint8_t* sB_hi = st + 2 * A_TILE_BYTES;
          uint8_t* sB_lo = sB_hi + C::B_TILE_BYTES;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d(sA_hi, &tmA_hi, &full_bar[s], k0, m0);
            if (p.passes == 3) tma_load_2d(sA_lo, &tmA_lo, &full_bar[s], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              tma_load_2d(sA_hi + j * 8192, &tmA_hi, &full_bar[s], m0 + 64 * j, k0);
              if (p.passes == 3) tma_load_2d(sA_lo + j * 8192, &tmA_lo, &full_bar[s], m0 + 64 * j, k0);
            }
          }
          if (!p.b_mn) {
            tma_load_2d(sB_hi, &tmB_hi, &full_bar[s], k0, n0);
            if (p.passes == 3) tma_load_2d(sB_lo, &tmB_lo, &full_bar[s], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              tma_load_2d(sB_hi + j * 8192, &tmB_hi, &full_bar[s], n0 + 64 * j, k0);
              if (p.passes == 3) tma_load_2d(sB_lo + j * 8192, &tmB_lo, &full_bar[s], n0 + 64 * j, k0);
            }
          }
          if (tlon && it == 0) tl[3] = clock64();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    {
      // instruction descriptor: D=f32, A=B=bf16, majors, N, M
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t adv_a = p.a_mn ? 2048u : 32u;  // bytes per k16 step
      const uint32_t adv_b = p.b_mn ? 2048u : 32u;
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
        const int t2 = tile / num_n;
        const int split = t2 / num_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const uint32_t a = acc_it & 1u;
        const uint32_t aph = (acc_it >> 1) & 1u;
        mbar_wait(&tempty_bar[a], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (tlon && it == 0 && lane == 0) tl[4] = clock64();
          const uint32_t sA_hi = smem_u32(tiles + s * C::STAGE_BYTES);
          const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
          const uint32_t sB_hi = sA_hi + 2 * A_TILE_BYTES;
          const uint32_t sB_lo = sB_hi + C::B_TILE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k16 = 0; k16 < BK / 16; ++k16) {
              const uint64_t da_hi = make_smem_desc(sA_hi + k16 * adv_a, p.a_mn);
              const uint64_t db_hi = make_smem_desc(sB_hi + k16 * adv_b, p.b_mn);
              const uint32_t accumulate = (kb > kb0 || k16 > 0) ? 1u : 0u;
              umma_bf16(d_tmem, da_hi, db_hi, idesc, accumulate);
              if (p.passes == 3) {
                const uint64_t da_lo = make_smem_desc(sA_lo + k16 * adv_a, p.a_mn);
                const uint64_t db_lo = make_smem_desc(sB_lo + k16 * adv_b, p.b_mn);
                umma_bf16(d_tmem, da_lo, db_hi, idesc, 1u);
                umma_bf16(d_tmem, da_hi, db_lo, idesc, 1u);
              }
            }
            umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&tfull_bar[a]);  // accumulator complete -> epilogue (same elected lane as the MMAs)
        __syncwarp();
        if (tlon && lane == 0) tl[5] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue warps (TMEM -> registers -> global)
    const int g = warp & 3;  // TMEM lane quarter this warp may access
    const DropKey seed = epilogue_seed<F>(p);
    // Operands of the epilogue that do not depend on the accumulator (residual, activation-backward operand, bias) are
    // requested one 32-column chunk AHEAD: the first chunk's loads are issued before the wait for the accumulator, so for
    // the decoder-sized problems (one tile per CTA) they fly during the whole main loop; later chunks' loads fly during the
    // previous chunk's TMEM read / staging / stores.  Loading them inside the store loop put every load behind the
    // preceding global stores (the compiler must assume out_f32 aliases residual): ~8 dependent L2 round trips per chunk,
    // which doubled the duration of the 16-CTA decoder GEMMs (profiles/r01j launch list: 19.5 us vs 10.7 us).
    constexpr bool kPreRes = (F & EF_RES) != 0, kPreAux = (F & EF_ACTBWD) != 0, kPreBias = (F & EF_BIAS) != 0;
    struct Pre {
      float4 res[kPreRes ? 8 : 1];
      float4 aux[kPreAux ? 8 : 1];
      uint2 relu[kPreAux ? 8 : 1];
      float4 bias;
    };
    const bool has_res = kPreRes && p.residual != nullptr;
    const bool has_aux = kPreAux && p.act_backward && (p.act == ACT_GELU || p.act == ACT_TANH || p.act == ACT_MUL_AUX);
    const bool has_relu = kPreAux && p.act_backward && p.act == ACT_RELU;
    const bool has_bias = kPreBias && p.bias != nullptr;
    auto preload = [&](Pre& P, int m0, int col0) {
      const int col = col0 + 4 * (lane & 7);
      const bool col_ok = col < p.N;
      if constexpr (kPreBias) {
        P.bias = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias && col_ok) P.bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long row = (long long)m0 + g * 32 + it * 4 + (lane >> 3);
        const bool ok = col_ok && row < p.M;
        if constexpr (kPreRes) {
          P.res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_res && ok) P.res[it] = *reinterpret_cast<const float4*>(p.residual + row * p.ld_res + col);
        }
        if constexpr (kPreAux) {
          P.aux[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          P.relu[it] = make_uint2(0u, 0u);
          if (has_aux && ok) P.aux[it] = *reinterpret_cast<const float4*>(p.act_aux + row * p.ld_act_aux + col);
          if (has_relu && ok) P.relu[it] = *reinterpret_cast<const uint2*>(p.act_aux_hi + row * p.ld_act_aux_hi + col);
        }
      }
    };
    uint32_t acc_it = 0;
    Pre nxt;
    bool nxt_valid = false;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
      const int n_idx = tile % num_n;
      const int t2 = tile / num_n;
      const int m_idx = t2 % num_m;
      const int split = t2 / num_m;
      const int m0 = m_idx * BM, n0 = n_idx * BN;
      const uint32_t a = acc_it & 1u;
      const uint32_t aph = (acc_it >> 1) & 1u;
      if (!nxt_valid && !(p.debug & 6)) preload(nxt, m0, n0);
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      if (tlon && warp == 2 && lane == 0) tl[6] = clock64();
      // Each thread owns one accumulator ROW in TMEM (tcgen05.ld 32x32b).  Storing rows directly would scatter every
      // 128-bit store over 32 different lines: stage 32x32 chunks through shared memory so that 8 lanes write 128
      // contiguous bytes of a row.  The TMEM stage is handed back to the MMA issuer as soon as the last chunk has been
      // read, before that chunk's store phase.
      float* stg = reinterpret_cast<float*>(tiles + C::STAGES * C::STAGE_BYTES) + (warp - 2) * 32 * STG_LD;
      // store-loop unroll factor: the heavy bodies (erf, dropout RNG) are not fully unrolled, to bound the code size (instruction
      // cache); the preloaded operands are then picked with static indices + selects (epi_pick)
      constexpr int kU = ((F & (EF_ACT | EF_ACTBWD | EF_DROP | EF_PREACT | EF_ATOMIC)) != 0) ? ((F & EF_ROWADD) ? 2 : 4) : 8;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        if (!(p.debug & 4)) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(g * 32) << 16) + a * BN + c * 32, r);
          tmem_ld_wait();
        }
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[a]);
        }
        if (p.debug & 6) {
          if ((p.debug & 2) && r[0] == 0x7fc12345u) p.out_f32[0] = 1.f;  // keep the loads alive
          continue;
        }
        const Pre cur = nxt;
        nxt_valid = false;
        if (c + 1 < BN / 32) {
          if (n0 + (c + 1) * 32 < p.N) preload(nxt, m0, n0 + (c + 1) * 32);  // warp-uniform
        } else {
          const int ntile = tile + gridDim.x;
          if (ntile < total_tiles) {  // first chunk of this CTA's next tile
            preload(nxt, ((ntile / num_n) % num_m) * BM, (ntile % num_n) * BN);
            nxt_valid = true;
          }
        }
        const int col0 = n0 + c * 32;
        if (col0 < p.N) {  // warp-uniform
          float* myrow = stg + lane * STG_LD;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(myrow + 4 * q) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                    __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
          __syncwarp();
          const int q = lane & 7;
          const int col = col0 + 4 * q;
          float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int o = 0; o < 8 / kU; ++o) {
#pragma unroll
          for (int j = 0; j < kU; ++j) {
            const int it = o * kU + j;
            const int rr = it * 4 + (lane >> 3);
            const long long row = (long long)m0 + g * 32 + rr;
            if (row < p.M && col < p.N) {
              const float4 t4 = *reinterpret_cast<const float4*>(stg + rr * STG_LD + 4 * q);
              float v[4] = {t4.x, t4.y, t4.z, t4.w};
              if (p.debug & 1) {
                if (v[0] == 123.456f) p.out_f32[0] = v[1];
              } else {
                epilogue_quad<F>(p, v, row, col, split == 0, seed, has_res, epi_pick<kU>(cur.res, o, j), has_aux,
                                 epi_pick<kU>(cur.aux, o, j), has_relu, epi_pick<kU>(cur.relu, o, j), has_bias, cur.bias);
              }
              if constexpr ((F & EF_COLSUM) != 0) { cs[0] += v[0]; cs[1] += v[1]; cs[2] += v[2]; cs[3] += v[3]; }
            }
          }
          }
          if constexpr ((F & EF_COLSUM) != 0) {
            if (p.colsum != nullptr) {  // column sums of the final values: combine the 4 lanes that share a column quad
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
                cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
              }
              if (lane < 8 && col < p.N) {
#pragma unroll
                for (int i = 0; i < 4; ++i) atomicAdd(p.colsum + col + i, cs[i]);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }

  if (tlon && warp == 2 && lane == 0) tl[7] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tlon && threadIdx.x == 0) {
    tl[8] = clock64();
    printf("gemm timeline M%d N%d K%d (cycles since entry): setup %lld | pdl_wait %lld | first TMA issued %lld | first stage landed %lld | "
           "last MMA committed %lld | accumulator ready %lld | epilogue done %lld | all warps done %lld\n",
           p.M, p.N, p.K, tl[1] - tl[0], tl[2] - tl[0], tl[3] - tl[0], tl[4] - tl[0], tl[5] - tl[0], tl[6] - tl[0], tl[7] - tl[0], tl[8] - tl[0]);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): a cluster of two CTAs on the two SMs of a TPC computes one 256 x 256 tile.  Each CTA stages
// 128 rows of A and 128 of the tile's 256 B rows per k-block and the leader's MMA reads both CTAs' shared memory, so the
// operand bytes fetched from L2 per FLOP are half those of the 128 x 128 single-CTA kernel above - which ncu and the
// VC_GEMM_DEBUG=4 experiment (profiles/r01b_gemm_epilogue_experiment.txt) showed to be bound by L2->SM operand traffic
// (64 KiB per k-block per SM in split-bf16 mode), not by the tensor pipe.
//   warp 0      TMA producer (both CTAs; transaction bytes of both land on the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only); tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs
//   warps 2..   epilogue (EW = 8 or 16 warps): TMEM lane quarter = warp % 4, column group = (warp - 2) / 4; 16-column chunks
//               staged through an XOR-swizzled (unpadded) 32 x 16 fp32 buffer per warp
// ---------------------------------------------------------------------------------------------------
constexpr int P_BM = 256;
constexpr int P_MAX_EW = 16;                         // epilogue warps per CTA (template parameter EW: 8 or 16)
constexpr int P_TILE_BYTES = 128 * BK * 2;           // 16 KiB: one 128-row operand tile (hi or lo)
constexpr int P_CW = 16;                             // epilogue chunk width (columns)
constexpr int P_STG_LD = 16;                         // floats per staged row: unpadded, float4 slots XOR-swizzled by (row >> 1) & 3
constexpr int P_STG_BYTES = P_MAX_EW * 32 * P_STG_LD * 4;
constexpr int P_BAR_BYTES = 256;
// Tile width PBN (template parameter): 256, or 128 for problems whose 256-wide tiling leaves the last wave mostly empty.
// The N = 512 image-encoder GEMMs at C1 are 100 tiles of 256 x 256 on 74 SM pairs: two waves for 1.35 waves of work, and
// the second wave's (heavy) epilogue runs on a third of the chip -- 26.6 us with the epilogue removed altogether against
// 14.6 us of MMA work (profiles/r01l_gemm_debug.txt).  As 200 tiles of 256 x 128 the same problem is 2.7 waves of half-size
// tiles; each CTA then stages 128 rows of A and 64 of the tile's 128 B rows per k-block (48 KiB, four stages).
template <int PBN>
struct PCfg {
  static constexpr int B_ROWS = PBN / 2;                             // B rows staged by each CTA of the pair
  static constexpr int B_TILE_BYTES = B_ROWS * BK * 2;               // 16 or 8 KiB
  static constexpr int STAGE_BYTES = 2 * P_TILE_BYTES + 2 * B_TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo
  static constexpr int STAGES = PBN == 256 ? 3 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + P_STG_BYTES + P_BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * PBN;                          // two accumulator stages
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <uint32_t F, int EW, int PBN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const GemmParams p) {
  pdl_trigger();  // the wait follows the barrier / TMEM set-up, which touches no global memory
  using C = PCfg<PBN>;
  constexpr int P_STAGES = C::STAGES, P_STAGE_BYTES = C::STAGE_BYTES, P_BN = PBN;
  __shared__ long long ptl[16];  // VC_GEMM_DEBUG & 8: clock64 timeline of the leader CTA of cluster 0 (experiment)
  const bool ptlon = (p.debug & 8) && blockIdx.x == 0;
  if (ptlon && threadIdx.x == 0) { for (int i = 0; i < 16; ++i) ptl[i] = 0; ptl[0] = clock64(); }
  extern __shared__ uint8_t smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);  // used in the leader CTA only
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tfull_bar = empty_bar + P_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;                        // used in the leader CTA only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_off = ((raw_addr + P_BAR_BYTES + 1023u) & ~1023u) - raw_addr;
  uint8_t* tiles = smem_raw + tiles_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (p.passes == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EW);  // EW epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (ptlon && threadIdx.x == 0) ptl[1] = clock64();
  pdl_wait();

  const int num_m = (p.M + P_BM - 1) / P_BM;
  const int num_n = (p.N + P_BN - 1) / P_BN;
  const int total_tiles = num_m * num_n * p.splitk;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t tx_bytes = (uint32_t)((p.passes == 3 ? 2 : 1) * (P_TILE_BYTES + C::B_TILE_BYTES)) * 2u;  // both CTAs' loads of one stage

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs); warp-uniform loop, one elected lane issues
    {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int n_idx = tile % num_n;
        const int t2 = tile / num_n;
        const int m_idx = t2 % num_m;
        const int split = t2 / num_m;
        const int m0 = m_idx * P_BM + (int)rank * 128, n0 = n_idx * P_BN + (int)rank * C::B_ROWS;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (!elect_one()) continue;
          uint8_t* st = tiles + s * P_STAGE_BYTES;
          uint8_t* sA_hi = st;
          uint8_t* sA_lo = st + P_TILE_BYTES;
          uint8_t* sB_hi = st + 2 * P_TILE_BYTES;
          uint8_t* sB_lo = sB_hi + C::B_TILE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[s], tx_bytes);
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d_pair(sA_hi, &tmA_hi, fb, k0, m0);
            if (p.passes == 3) tma_load_2d_pair(sA_lo, &tmA_lo, fb, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_2d_pair(sA_hi + j * 8192, &tmA_hi, fb, m0 + 64 * j, k0);
              if (p.passes == 3) tma_load_2d_pair(sA_lo + j * 8192, &tmA_lo, fb, m0 + 64 * j, k0);
            }
          }
          if (!p.b_mn) {
            tma_load_2d_pair(sB_hi, &tmB_hi, fb, k0, n0);
            if (p.passes == 3) tma_load_2d_pair(sB_lo, &tmB_lo, fb, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 64; ++j) {
              tma_load_2d_pair(sB_hi + j * 8192, &tmB_hi, fb, n0 + 64 * j, k0);
              if (p.passes == 3) tma_load_2d_pair(sB_lo + j * 8192, &tmB_lo, fb, n0 + 64 * j, k0);
            }
          }
        }
      }
      // tail: wait until every stage-free arrival the leader's MMA commits multicast to this CTA has landed, so that no
      // remote arrival can target this CTA's shared memory after it exits
      for (uint32_t j = 0; j < (uint32_t)P_STAGES; ++j, ++it) {
        const uint32_t s = it % P_STAGES;
        const uint32_t ph = (it / P_STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA); warp-uniform loop, one elected lane issues
    if (leader) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)(P_BM >> 4) << 24);
      const uint32_t adv_a = p.a_mn ? 2048u : 32u;  // bytes per k16 step
      const uint32_t adv_b = p.b_mn ? 2048u : 32u;
      uint32_t it = 0, acc_it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++acc_it) {
        const int t2 = tile / num_n;
        const int split = t2 / num_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const uint32_t a = acc_it & 1u;
        const uint32_t aph = (acc_it >> 1) & 1u;
        mbar_wait(&tempty_bar[a], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * P_BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (ptlon && lane == 0 && kb == kb0 && acc_it < 2) ptl[2 + 4 * acc_it] = clock64();  // first stage of the tile landed
          const uint32_t sA_hi = smem_u32(tiles + s * P_STAGE_BYTES);
          const uint32_t sA_lo = sA_hi + P_TILE_BYTES;
          const uint32_t sB_hi = sA_hi + 2 * P_TILE_BYTES;
          const uint32_t sB_lo = sB_hi + C::B_TILE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k16 = 0; k16 < BK / 16; ++k16) {
              const uint64_t da_hi = make_smem_desc(sA_hi + k16 * adv_a, p.a_mn);
              const uint64_t db_hi = make_smem_desc(sB_hi + k16 * adv_b, p.b_mn);
              const uint32_t accumulate = (kb > kb0 || k16 > 0) ? 1u : 0u;
              umma_bf16_pair(d_tmem, da_hi, db_hi, idesc, accumulate);
              if (p.passes == 3) {
                const uint64_t da_lo = make_smem_desc(sA_lo + k16 * adv_a, p.a_mn);
                const uint64_t db_lo = make_smem_desc(sB_lo + k16 * adv_b, p.b_mn);
                umma_bf16_pair(d_tmem, da_lo, db_hi, idesc, 1u);
                umma_bf16_pair(d_tmem, da_hi, db_lo, idesc, 1u);
              }
            }
            umma_commit_pair(&empty_bar[s]);  // frees this smem stage in both CTAs
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_pair(&tfull_bar[a]);  // accumulator complete -> both CTAs' epilogue warps
        __syncwarp();
        if (ptlon && lane == 0 && acc_it < 2) ptl[3 + 4 * acc_it] = clock64();  // last MMA of the tile issued
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue warps (this CTA's 128 rows x 256 columns)
    const int g = warp & 3;          // TMEM lane quarter this warp may access
    const int hc = (warp - 2) >> 2;  // column group
    constexpr int kCols = P_BN * 4 / EW;  // columns per epilogue warp: 128 (EW = 8) or 64 (EW = 16)
    const uint32_t tempty_remote0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_remote1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
    float* stg = reinterpret_cast<float*>(tiles + P_STAGES * P_STAGE_BYTES) + (warp - 2) * 32 * P_STG_LD;
    const DropKey seed = epilogue_seed<F>(p);
    uint32_t acc_it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++acc_it) {
      const int n_idx = tile % num_n;
      const int t2 = tile / num_n;
      const int m_idx = t2 % num_m;
      const int split = t2 / num_m;
      const int m0 = m_idx * P_BM + (int)rank * 128, n0 = n_idx * P_BN + hc * kCols;
      const uint32_t a = acc_it & 1u;
      const uint32_t aph = (acc_it >> 1) & 1u;
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      if (ptlon && warp == 2 && lane == 0 && acc_it < 2) ptl[4 + 4 * acc_it] = clock64();  // accumulator ready
      constexpr int kUnrollIt = 4;  // fully unrolled: the preloaded operands are indexed by `it`
#pragma unroll 1
      for (int c = 0; c < kCols / P_CW; ++c) {
        // The epilogue is latency-bound (4 warps per scheduler, each iteration behind a global load): issue this chunk's
        // residual / activation-backward operand loads first, so that they fly during the TMEM read and the staging.
        constexpr bool kPreRes = (F & EF_RES) != 0, kPreAux = (F & EF_ACTBWD) != 0;
        float4 pres[kPreRes ? 4 : 1], paux[kPreAux ? 4 : 1];
        bool has_res = false, has_aux = false;
        {
          const int pcol = n0 + c * P_CW + 4 * (lane & 3);
          if constexpr (kPreRes) has_res = p.residual != nullptr;
          if constexpr (kPreAux) has_aux = p.act_backward && (p.act == ACT_GELU || p.act == ACT_TANH || p.act == ACT_MUL_AUX);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const long long prow = (long long)m0 + g * 32 + it * 8 + (lane >> 2);
            const bool ok = prow < p.M && pcol < p.N;
            if constexpr (kPreRes) {
              pres[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_res && ok) pres[it] = *reinterpret_cast<const float4*>(p.residual + prow * p.ld_res + pcol);
            }
            if constexpr (kPreAux) {
              paux[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_aux && ok) paux[it] = *reinterpret_cast<const float4*>(p.act_aux + prow * p.ld_act_aux + pcol);
            }
          }
        }
        uint32_t r[P_CW];
        if (!(p.debug & 4)) {
          tmem_ld_32x16(tmem_base + ((uint32_t)(g * 32) << 16) + a * P_BN + hc * kCols + c * P_CW, r);
          tmem_ld_wait();
        }
        if (c == kCols / P_CW - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(a ? tempty_remote1 : tempty_remote0);
        }
        if (p.debug & 6) {  // VC_GEMM_DEBUG experiments: 2 = TMEM loads only, 4 = not even those
          if ((p.debug & 2) && r[0] == 0x7fc12345u) p.out_f32[0] = 1.f;  // keep the loads alive
          continue;
        }
        const int col0 = n0 + c * P_CW;
        float* myrow = stg + lane * P_STG_LD;
#pragma unroll
        for (int q = 0; q < P_CW / 4; ++q)
          *reinterpret_cast<float4*>(myrow + 4 * (q ^ ((lane >> 1) & 3))) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        __syncwarp();
        const int q = lane & 3;
        const int col = col0 + 4 * q;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll kUnrollIt
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2);
          const long long row = (long long)m0 + g * 32 + rr;
          if (row < p.M && col < p.N) {
            const float4 t4 = *reinterpret_cast<const float4*>(stg + rr * P_STG_LD + 4 * (q ^ ((rr >> 1) & 3)));
            float v[4] = {t4.x, t4.y, t4.z, t4.w};
            if (p.debug & 1) {  // VC_GEMM_DEBUG = 1: no epilogue math, no global stores
              if (v[0] == 123.456f) p.out_f32[0] = v[1];
            } else {
              epilogue_quad<F>(p, v, row, col, split == 0, seed, kPreRes && has_res, pres[kPreRes ? it : 0], kPreAux && has_aux,
                               paux[kPreAux ? it : 0]);
            }
            if constexpr ((F & EF_COLSUM) != 0) { cs[0] += v[0]; cs[1] += v[1]; cs[2] += v[2]; cs[3] += v[3]; }
          }
        }
        if constexpr ((F & EF_COLSUM) != 0) {
          if (p.colsum != nullptr) {  // combine the 8 lanes that share a column quad
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 4);
              cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
              cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
            }
            if (lane < 4 && col < p.N) {
#pragma unroll
              for (int i = 0; i < 4; ++i) atomicAdd(p.colsum + col + i, cs[i]);
            }
          }
        }
        __syncwarp();
      }
      if (ptlon && warp == 2 && lane == 0 && acc_it < 2) ptl[5 + 4 * acc_it] = clock64();  // this warp's epilogue of the tile done
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (ptlon && threadIdx.x == 0) {
    const long long z = ptl[0];
    printf("pair timeline M%d N%d K%d BN%d (cycles since entry): setup %lld | tile0: first stage %lld, MMAs issued %lld, acc ready %lld, epilogue done %lld | "
           "tile1: first stage %lld, MMAs issued %lld, acc ready %lld, epilogue done %lld | all done %lld\n",
           p.M, p.N, p.K, PBN, ptl[1] - z, ptl[2] - z, ptl[3] - z, ptl[4] - z, ptl[5] - z, ptl[6] ? ptl[6] - z : 0, ptl[7] ? ptl[7] - z : 0,
           ptl[8] ? ptl[8] - z : 0, ptl[9] ? ptl[9] - z : 0, clock64() - z);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// cuTensorMapEncodeTiled costs a few microseconds of host time; the same (buffer, shape) pairs recur every training step
// (weights, persistent workspaces), so encoded maps are cached.
struct TmapKey {
  const void* base; int64_t rows, cols, ld; int box_cols, box_rows;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.rows); mix((size_t)k.cols); mix((size_t)k.ld); mix((size_t)k.box_cols * 1024 + (size_t)k.box_rows);
    return h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int make_tmap_uncached(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows);

// bf16 matrix stored [rows, cols] row-major with leading dimension ld; box = {box_cols (inner), box_rows}
int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  TmapKey key{base, rows, cols, ld, box_cols, box_rows};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *tm = it->second; return 0; }
  }
  if (int rc = make_tmap_uncached(tm, base, rows, cols, ld, box_cols, box_rows)) return rc;
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 16384) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *tm);
  return 0;
}

int make_tmap_uncached(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d): base=%p rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
             base, (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows);
    return set_error(msg);
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

uint32_t required_features(const GemmDesc& d) {
  uint32_t f = 0;
  if (d.bias) f |= EF_BIAS;
  if (d.rowadd) f |= EF_ROWADD;
  if (d.preact) f |= EF_PREACT;
  if (d.act != VC_ACT_NONE && !d.act_backward) f |= EF_ACT;
  if (d.drop.p > 0.f) f |= EF_DROP;
  if (d.act_backward && d.act != VC_ACT_NONE) f |= EF_ACTBWD;
  if (d.residual) f |= EF_RES;
  if (d.out_f32) f |= (d.splitk > 1 ? EF_ATOMIC : EF_F32);
  if (d.out_hi) f |= EF_SPLIT;
  if (d.colsum) f |= EF_COLSUM;
  return f;
}

template <int BN, uint32_t F>
int launch_gemm_variant(const GemmDesc& d, cudaStream_t stream);
struct GemmParams;
void fill_params(GemmParams& p, const GemmDesc& d, int splitk_req);
int make_operand_maps(const GemmDesc& d, CUtensorMap& tA_hi, CUtensorMap& tA_lo, CUtensorMap& tB_hi, CUtensorMap& tB_lo, int bn);

// variants, most specific first; the first whose mask covers the required features is used
constexpr uint32_t V_F32 = EF_F32;
constexpr uint32_t V_F32_BIAS = EF_F32 | EF_BIAS;
constexpr uint32_t V_F32_RES = EF_F32 | EF_RES | EF_BIAS;
constexpr uint32_t V_F32_DROP_RES = EF_F32 | EF_BIAS | EF_DROP | EF_RES;
constexpr uint32_t V_SPLIT = EF_SPLIT | EF_BIAS;
constexpr uint32_t V_SPLIT_ACT = EF_SPLIT | EF_BIAS | EF_ACT | EF_DROP | EF_PREACT;
constexpr uint32_t V_SPLIT_BWD = EF_SPLIT | EF_ACTBWD | EF_DROP | EF_COLSUM;
constexpr uint32_t V_ATOMIC = EF_ATOMIC | EF_BIAS;

template <int BN>
int launch_gemm(const GemmDesc& d, cudaStream_t stream) {
  const uint32_t req = required_features(d);
#define VC_TRY_VARIANT(V) if ((req & ~(V)) == 0) return launch_gemm_variant<BN, (V)>(d, stream);
  VC_TRY_VARIANT(V_F32)
  VC_TRY_VARIANT(V_F32_BIAS)
  VC_TRY_VARIANT(V_ATOMIC)
  VC_TRY_VARIANT(V_F32_RES)
  VC_TRY_VARIANT(V_F32_DROP_RES)
  VC_TRY_VARIANT(V_SPLIT)
  VC_TRY_VARIANT(V_SPLIT_ACT)
  VC_TRY_VARIANT(V_SPLIT_BWD)
#undef VC_TRY_VARIANT
  return launch_gemm_variant<BN, EF_ALL>(d, stream);
}

template <int BN, uint32_t F>
int launch_gemm_variant(const GemmDesc& d, cudaStream_t stream) {
  using C = Cfg<BN>;
  GemmParams p;
  fill_params(p, d, d.splitk);
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  if (int rc = make_operand_maps(d, tA_hi, tA_lo, tB_hi, tB_lo, BN)) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_m = (d.M + BM - 1) / BM, num_n = (d.N + BN - 1) / BN;
  const long long tiles = (long long)num_m * num_n * p.splitk;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  int slot = -1;
  char tag[64];
  snprintf(tag, sizeof tag, "M%d N%d K%d a%d b%d s%d p%d e%d%d%d%d", d.M, d.N, d.K, p.a_mn, p.b_mn, p.splitk, p.passes,
           d.out_f32 ? 1 : 0, d.out_hi ? 1 : 0, d.residual ? 1 : 0, d.act);
  const bool prof = gemm_profile_begin(stream, 2.0 * (double)d.M * (double)d.N * (double)d.K, &slot, tag);
  VC_LAUNCH((gemm_tc_kernel<BN, F>), grid, GEMM_THREADS, C::SMEM_BYTES, stream, tA_hi, tA_lo, tB_hi, tB_lo, p);
  if (prof) gemm_profile_end(stream, slot);
  return check_launch("gemm_tc_kernel");
}


// split count for an accumulating (split-K) GEMM of `tiles` output tiles and `num_kb` k-blocks on `units` concurrent
// CTAs / CTA pairs: minimise waves x (k-blocks per split + a per-tile epilogue allowance)
int choose_splitk(long long tiles, int num_kb, int units) {
  int best = 1;
  long long best_cost = -1;
  const int smax = num_kb / 4 > 0 ? num_kb / 4 : 1;
  for (int s = 1; s <= smax; ++s) {
    const long long waves = (tiles * s + units - 1) / units;
    const long long cost = waves * ((num_kb + s - 1) / s + 4);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

void fill_params(GemmParams& p, const GemmDesc& d, int splitk_req) {
  memset(&p, 0, sizeof p);
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.passes = d.passes;
  p.a_mn = d.a_mn_major ? 1 : 0;
  p.b_mn = d.b_mn_major ? 1 : 0;
  p.num_kb = (d.K + BK - 1) / BK;
  int splitk = splitk_req < 1 ? 1 : splitk_req;
  if (splitk > p.num_kb) splitk = p.num_kb;
  p.kb_per_split = (p.num_kb + splitk - 1) / splitk;
  p.splitk = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.bias = d.bias;
  p.rowadd = d.rowadd; p.ld_rowadd = d.ld_rowadd;
  p.rowadd_div = d.rowadd_div > 0 ? d.rowadd_div : 1;
  p.rowadd_mod = d.rowadd_mod > 0 ? d.rowadd_mod : 1;
  p.preact = d.preact; p.ld_preact = d.ld_preact;
  p.act = d.act;
  p.drop_on = d.drop.p > 0.f ? 1 : 0;
  p.drop_scale = d.drop.p > 0.f ? 1.0f / (1.0f - d.drop.p) : 1.0f;
  p.drop_thresh = dropout_threshold(d.drop.p);
  p.drop_site = d.drop.site;
  p.drop_seed = d.drop.seed;
  p.drop_seed_ptr = d.drop.seed_ptr;
  p.residual = d.residual; p.ld_res = d.ld_res;
  p.out_f32 = d.out_f32; p.ldo = d.ldo;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(d.out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
  p.ldo_split = d.ldo_split;
  p.act_backward = d.act_backward ? 1 : 0;
  p.act_aux = d.act_aux; p.ld_act_aux = d.ld_act_aux;
  p.act_aux_hi = reinterpret_cast<const __nv_bfloat16*>(d.act_aux_hi); p.ld_act_aux_hi = d.ld_act_aux_hi;
  p.colsum = d.colsum;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VC_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
}

int make_operand_maps(const GemmDesc& d, CUtensorMap& tA_hi, CUtensorMap& tA_lo, CUtensorMap& tB_hi, CUtensorMap& tB_lo, int bn) {
  int rc = 0;
  const bf16_t* a_lo = d.passes == 3 ? d.a_lo : d.a_hi;
  const bf16_t* b_lo = d.passes == 3 ? d.b_lo : d.b_hi;
  if (!d.a_mn_major) {
    rc |= make_tmap(&tA_hi, d.a_hi, d.M, d.K, d.lda, BK, BM);
    rc |= make_tmap(&tA_lo, a_lo, d.M, d.K, d.lda, BK, BM);
  } else {
    rc |= make_tmap(&tA_hi, d.a_hi, d.K, d.M, d.lda, 64, BK);
    rc |= make_tmap(&tA_lo, a_lo, d.K, d.M, d.lda, 64, BK);
  }
  if (!d.b_mn_major) {
    rc |= make_tmap(&tB_hi, d.b_hi, d.N, d.K, d.ldb, BK, bn);
    rc |= make_tmap(&tB_lo, b_lo, d.N, d.K, d.ldb, BK, bn);
  } else {
    rc |= make_tmap(&tB_hi, d.b_hi, d.K, d.N, d.ldb, 64, BK);
    rc |= make_tmap(&tB_lo, b_lo, d.K, d.N, d.ldb, 64, BK);
  }
  return rc;
}

int pair_epilogue_warps() {
  static int ew = 0;
  if (ew == 0) {
    const char* e = getenv("VC_GEMM_PAIR_EW");
    ew = (e && atoi(e) == 8) ? 8 : 16;
  }
  return ew;
}

template <uint32_t F, int EW, int PBN>
int launch_gemm_pair_ew(const GemmDesc& d, int splitk, cudaStream_t stream);

template <uint32_t F>
int launch_gemm_pair_variant(const GemmDesc& d, int splitk, int pbn, cudaStream_t stream) {
  if (pbn == 128)
    return pair_epilogue_warps() == 8 ? launch_gemm_pair_ew<F, 8, 128>(d, splitk, stream) : launch_gemm_pair_ew<F, 16, 128>(d, splitk, stream);
  return pair_epilogue_warps() == 8 ? launch_gemm_pair_ew<F, 8, 256>(d, splitk, stream) : launch_gemm_pair_ew<F, 16, 256>(d, splitk, stream);
}

template <uint32_t F, int EW, int PBN>
int launch_gemm_pair_ew(const GemmDesc& d, int splitk, cudaStream_t stream) {
  using C = PCfg<PBN>;
  constexpr int P_SMEM_BYTES = C::SMEM_BYTES, P_BN = PBN;
  GemmParams p;
  fill_params(p, d, splitk);
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
  if (int rc = make_operand_maps(d, tA_hi, tA_lo, tB_hi, tB_lo, C::B_ROWS)) return rc;  // each CTA stages half of the tile's B rows
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_pair_kernel<F, EW, PBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    attr_set = true;
  }
  const long long tiles = (long long)((d.M + P_BM - 1) / P_BM) * ((d.N + P_BN - 1) / P_BN) * p.splitk;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (int)(tiles < pairs ? tiles : pairs);
  int slot = -1;
  char tag[64];
  snprintf(tag, sizeof tag, "M%d N%d K%d a%d b%d s%d p%d e%d%d%d%d P%d", d.M, d.N, d.K, p.a_mn, p.b_mn, p.splitk, p.passes,
           d.out_f32 ? 1 : 0, d.out_hi ? 1 : 0, d.residual ? 1 : 0, d.act, PBN);
  const bool prof = gemm_profile_begin(stream, 2.0 * (double)d.M * (double)d.N * (double)d.K, &slot, tag);
  VC_LAUNCH((gemm_tc_pair_kernel<F, EW, PBN>), grid, 64 + 32 * EW, P_SMEM_BYTES, stream, tA_hi, tA_lo, tB_hi, tB_lo, p);
  count_pair_launch();
  if (prof) gemm_profile_end(stream, slot);
  return check_launch("gemm_tc_pair_kernel");
}

int launch_gemm_pair(const GemmDesc& d, int splitk, int pbn, cudaStream_t stream) {
  GemmDesc dd = d;
  dd.splitk = splitk;
  const uint32_t req = required_features(dd);
#define VC_TRY_VARIANT(V) if ((req & ~(V)) == 0) return launch_gemm_pair_variant<(V)>(d, splitk, pbn, stream);
  VC_TRY_VARIANT(V_F32)
  VC_TRY_VARIANT(V_F32_BIAS)
  VC_TRY_VARIANT(V_ATOMIC)
  VC_TRY_VARIANT(V_F32_RES)
  VC_TRY_VARIANT(V_F32_DROP_RES)
  VC_TRY_VARIANT(V_SPLIT)
  VC_TRY_VARIANT(V_SPLIT_ACT)
  VC_TRY_VARIANT(V_SPLIT_BWD)
#undef VC_TRY_VARIANT
  return launch_gemm_pair_variant<EF_ALL>(d, splitk, pbn, stream);
}

// The CTA-pair kernel handles problems whose M is a multiple of 256 and whose N is a multiple of its tile width (256 or 128) and
// that keep at least half of the 74 SM pairs busy (the image-encoder GEMMs); everything else (decoder-sized, ragged N) uses the
// single-CTA kernel.  Tile width: 256 whenever N allows it.  A wave model (waves x (k-blocks per tile x MMA time per k-block +
// epilogue allowance); VC_GEMM_PAIR_BN=-2 selects it) picks 128 for the N = 512 GEMMs, whose 100 wide tiles leave the second
// wave two-thirds empty, and each such GEMM alone does run 4-9 us faster -- but the training step got 1.7 % SLOWER
// (profiles/r01l_tile_width_ab.txt): inside the step those idle SMs are not idle, they run the CAD encoder and the
// weight-gradient GEMMs of the auxiliary streams, and the narrow tile moves 1.5x the operand bytes per FLOP through L2.
int g_pair_force_bn = -1;  // -1: undecided (VC_GEMM_PAIR_BN in the environment), 0: default policy, 128 / 256: forced, -2: wave model

bool pair_eligible(const GemmDesc& d, int* splitk_out, int* pbn_out) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("VC_GEMM_PAIR"); enabled = e ? atoi(e) : 1; }
  if (g_pair_force_bn == -1) { const char* e = getenv("VC_GEMM_PAIR_BN"); g_pair_force_bn = e ? atoi(e) : 0; }
  const int force_bn = g_pair_force_bn;
  if (!enabled) return false;
  if (d.M % P_BM != 0 || d.N % 128 != 0) return false;
  const int num_kb = (d.K + BK - 1) / BK;
  const int pairs = num_sms() / 2;
  double best_cost = -1.0;
  for (int pbn = 256; pbn >= 128; pbn -= 128) {
    if (d.N % pbn != 0) continue;
    if (force_bn == 128 || force_bn == 256) { if (pbn != force_bn && d.N % force_bn == 0) continue; }
    const long long tiles = (long long)(d.M / P_BM) * (d.N / pbn);
    int s = 1;
    if (d.splitk > 1) s = choose_splitk(tiles, num_kb, pairs);
    if (tiles * s < pairs / 2) continue;
    const long long waves = (tiles * s + pairs - 1) / pairs;
    const double kb = (double)((num_kb + s - 1) / s);
    double cost = pbn == 256 ? (double)waves * (kb + 4.0) : (double)waves * (kb * 0.55 + 2.0);
    if (force_bn != -2 && pbn == 128) cost += 1e9;  // default policy: the narrow tile only where the wide one does not apply
    if (best_cost < 0.0 || cost < best_cost) { best_cost = cost; *splitk_out = s; *pbn_out = pbn; }
  }
  return best_cost >= 0.0;
}

}  // namespace

void gemm_pair_force_tile(int bn) { g_pair_force_bn = (bn == 128 || bn == 256 || bn == -2) ? bn : 0; }

void gemm_desc_init(GemmDesc* d) {
  memset(d, 0, sizeof *d);
  d->passes = 3;
  d->splitk = 1;
  d->rowadd_div = 1;
  d->rowadd_mod = 1;
}

int gemm(const GemmDesc& d, stream_t stream) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return set_error("gemm: empty problem");
  if (d.N % 8 != 0) return set_error("gemm: N must be a multiple of 8");
  if (d.lda % 8 != 0 || d.ldb % 8 != 0) return set_error("gemm: lda/ldb must be multiples of 8 elements");
  if (d.passes != 1 && d.passes != 3) return set_error("gemm: passes must be 1 or 3");
  if (d.a_hi == nullptr || d.b_hi == nullptr) return set_error("gemm: null operand");
  if (d.passes == 3 && (d.a_lo == nullptr || d.b_lo == nullptr)) return set_error("gemm: passes=3 needs lo operands");
  if (d.act_backward && ((d.act == VC_ACT_GELU || d.act == VC_ACT_TANH || d.act == VC_ACT_MUL_AUX) && !d.act_aux)) return set_error("gemm: act_aux required");
  if (d.act == VC_ACT_GELU_DSTORE && (d.act_backward || !d.preact)) return set_error("gemm: VC_ACT_GELU_DSTORE is a forward activation and needs preact");
  if (d.act == VC_ACT_MUL_AUX && !d.act_backward) return set_error("gemm: VC_ACT_MUL_AUX is a backward-activation mode");
  if (d.act_backward && d.act == VC_ACT_RELU && !d.act_aux_hi) return set_error("gemm: act_aux_hi required");
  if (d.splitk > 1 && (d.act != VC_ACT_NONE || d.drop.p > 0.f || d.residual || d.out_hi || d.preact || d.colsum))
    return set_error("gemm: split-K supports only the bias epilogue with an fp32 (atomic) output");
  if (d.out_f32 == nullptr && d.out_hi == nullptr) return set_error("gemm: no output");
  if ((d.out_f32 && d.ldo % 4 != 0) || (d.out_hi && d.ldo_split % 4 != 0) || (d.residual && d.ld_res % 4 != 0) ||
      (d.preact && d.ld_preact % 4 != 0) || (d.rowadd && d.ld_rowadd % 4 != 0))
    return set_error("gemm: epilogue leading dimensions must be multiples of 4");
  // decoder-sized problems fill only a few SMs with 128x128 tiles: halve the tile width so that twice as many CTAs each
  // run half the MMA work
  int pair_splitk = 1, pair_bn = 256;
  if (pair_eligible(d, &pair_splitk, &pair_bn)) return launch_gemm_pair(d, pair_splitk, pair_bn, reinterpret_cast<cudaStream_t>(stream));
  const long long tiles128 = (long long)((d.M + BM - 1) / BM) * ((d.N + 127) / 128) * (d.splitk > 1 ? d.splitk : 1);
  if (tiles128 <= 48 && d.N >= 64) return launch_gemm<64>(d, reinterpret_cast<cudaStream_t>(stream));
  return launch_gemm<128>(d, reinterpret_cast<cudaStream_t>(stream));
}

}  // namespace vck
