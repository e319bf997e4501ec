// Host-callable launch API of the videocad_b200 kernels.
//
// Two implementations exist with IDENTICAL semantics:
//   * videocad_b200/csrc/*.cu           -- the product: hand-written sm_100a CUDA kernels (device pointers)
//   * oracle/csrc/kernels_cpu.cpp       -- test infrastructure: plain C++ loops on host pointers, used only
//                                          to check the host orchestration (model_*.cpp) without a GPU.
// The orchestration code (model_vit.cpp, model_seq.cpp) is written against this header only.
//
// Conventions
//   * all matrices are row-major with an explicit leading dimension (elements);
//   * "split" tensors are pairs of bf16 arrays (hi, lo) with x ~= hi + lo  (see common.cuh);
//   * every function enqueues work on `stream` and returns 0 on success (no host sync);
//   * no function allocates device memory.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "videocad_b200.h"  // the C ABI structs (include/): vc_drop, vc_gemm_desc, vc_attn_desc

namespace vck {

typedef void* stream_t;
typedef vc_bf16 bf16_t;  // raw bfloat16 bits
typedef vc_drop Drop;    // dropout call site; p == 0 disables
static inline Drop no_drop() { Drop d; d.p = 0.f; d.site = 0; d.seed = 0; d.seed_ptr = nullptr; return d; }
static inline Drop make_drop(float p, uint32_t site, uint64_t seed, const uint64_t* seed_ptr = nullptr) {
  Drop d; d.p = p; d.site = site; d.seed = seed; d.seed_ptr = seed_ptr; return d;
}

// -------------------------------------------------------------------------------------------------
// tensor-core GEMM (tcgen05 / TMEM / TMA):   acc[m,n] = sum_k A[m,k] * B[n,k]
//   A: a_mn_major == 0 -> stored [M,K] (lda >= K);  a_mn_major == 1 -> stored [K,M] (lda >= M)
//   B: b_mn_major == 0 -> stored [N,K] (ldb >= K);  b_mn_major == 1 -> stored [K,N] (ldb >= N)
//   passes == 3: hi*hi + lo*hi + hi*lo (fp32-grade), passes == 1: hi*hi only (bf16-grade)
// epilogue, in this order:
//   v = acc (+ bias[n]) (+ rowadd[((m / rowadd_div) % rowadd_mod), n]);  if (preact) preact[m,n] = v;
//   v = act(v);  v = dropout(v);  if (residual) v += residual[m,n];
//   out_f32[m,n] = v (or atomicAdd when splitk > 1);  (out_hi, out_lo)[m,n] = split(v)
// constraints: contiguous dims and all ld multiples of 8 elements (16-byte TMA strides); N % 8 == 0;
//   splitk > 1 only with bias/none epilogue and out_f32 pre-zeroed by the caller.
// -------------------------------------------------------------------------------------------------
typedef vc_gemm_desc GemmDesc;
void gemm_desc_init(GemmDesc* d);
int gemm(const GemmDesc& d, stream_t stream);
// tile width of the 2-SM GEMM kernel: 0 = chosen per problem by the wave model (default), 128 / 256 = forced where N allows
// (tests, experiments).  No effect in the CPU emulation.
void gemm_pair_force_tile(int bn);

// x[rows, cols] fp32 -> (hi, lo)
int split_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, bf16_t* hi, bf16_t* lo, int64_t ldo, stream_t s);

// many contiguous fp32 -> split conversions in one launch; `items` lives in device memory (see vc_split_item)
int split_many(const vc_split_item* items, int n_items, int64_t total_blocks, stream_t s);

// LayerNorm over the last dim (biased variance, eps inside sqrt).  Outputs optional (may be null).
int layernorm_fwd(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  float* y, int64_t ldy, bf16_t* y_hi, bf16_t* y_lo, int64_t ldy_split, float* mean, float* rstd,
                  stream_t s);
// dx = (dres ? dres : 0) + LN_bwd(dy);  dgamma/dbeta are ACCUMULATED (atomicAdd) into caller-zeroed buffers.
int layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                  const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                  float* dgamma, float* dbeta, stream_t s);

// same, with a fused second output for the next backward GEMMs: g = dx * dropout_mask(gdrop; element index row*C + c)
// written as split-bf16 (g_hi, g_lo) and, if g_colsum != null, g_colsum[c] += sum_rows g (atomic, pre-zeroed).
int layernorm_bwd_fused(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                        const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, Drop gdrop, bf16_t* g_hi, bf16_t* g_lo, int64_t ldg, float* g_colsum,
                        stream_t s);

// ViT front end: 'f 1 (h 32) (w 32) -> (f h w) 1024' gather + LayerNorm(1024) -> split operand of the patch GEMM
int patch_layernorm_fwd(const float* img, int F, int S, const float* gamma, const float* beta, float eps, bf16_t* y_hi,
                        bf16_t* y_lo, float* mean, float* rstd, stream_t s);
// parameter grads of that LayerNorm from dy[F*N, 1024] (accumulated into zeroed dgamma/dbeta); no input grad
int patch_layernorm_bwd_params(const float* img, int F, int S, const float* mean, const float* rstd, const float* dy,
                               float* dgamma, float* dbeta, stream_t s);

// tokens: x[f, 0] = cls + pos[0]; x[f, 1+p] = e[f*N+p] + pos[1+p]; then dropout(site) on the whole [F*(N+1), C] tensor
int vit_assemble_fwd(const float* e, int F, int N, int C, const float* cls, const float* pos, Drop drop, float* x,
                     stream_t s);
// de[f*N+p] = dx[f,1+p]*mask; dcls += sum_f dx[f,0]*mask; dpos[t] += sum_f dx[f,t]*mask  (dcls/dpos pre-zeroed)
int vit_assemble_bwd(const float* dx, int F, int N, int C, Drop drop, float* de, float* dcls, float* dpos, stream_t s);

// multi-head attention core on fp32 q/k/v with row strides; rows are (b * T + t); head h uses columns [h*d, (h+1)*d)
//   s = scale * q k^T + mask;  p = softmax(s);  p~ = dropout(p) (element index ((b*nh+h)*Tq+i)*Tk+j);  o = p~ v
//   mask: NONE | CAUSAL (j <= i) | WINDOW (i - window < j <= i)
// fwd writes o as split [B*Tq, nh*d] (+ optional fp32) and lse[B, nh, Tq] (log-sum-exp of s) for the backward.
typedef vc_attn_desc AttnDesc;
int attention_fwd(const AttnDesc& a, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t s);
// backward: recomputes p from (q, k, lse); needs do[B*Tq, nh*d] fp32 and o (split, as written by fwd)
//   dq/dk/dv written (not accumulated) as fp32 with their own strides
int attention_bwd(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                  const float* dout, int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                  int64_t lddv, stream_t s);

// (q/k/v may alternatively be given as split-bf16 arrays through AttnDesc::q_hi..v_lo; the generic SIMT kernels accept
//  fp32 inputs only, the tensor-core ViT kernel and the CPU emulation accept both)
// same backward, but dq/dk/dv are delivered as split-bf16 operands for the following dgrad/wgrad GEMMs (columns
// [0,nh*d) of three [B*T, ld_split] matrices).  `scratch` is an fp32 workspace of 3*B*T*nh*d floats that an
// implementation may use for an fp32 intermediate (the tensor-core ViT kernel writes split directly and ignores it).
int attention_bwd_split(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                        const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, float* scratch, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi,
                        bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, stream_t s);

// the same for attention behind a packed in_proj with bias (nn.MultiheadAttention): fp32 upstream gradient only; additionally
// accumulates the bias gradient = column sums of dq / dk / dv into dbq / dbk / dbv ([nh*d] each, caller-zeroed, may be null)
int attention_bwd_split_bias(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                             const float* dout, int64_t lddo, float* scratch, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi,
                             bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, float* dbq, float* dbk, float* dbv,
                             stream_t s);
// 0 routes short-sequence decoder attention (Tq == Tk <= 32) through the generic kernels (tests / A-B timing); default 1.
// No effect in the CPU emulation.
void attention_small_enable(int enable);

// backward of "v = dropout(act(pre))":  g = dy * mask*scale * act'(.)
//   act GELU: aux = pre-activation fp32; TANH: aux = forward output fp32; RELU: aux_hi = forward output hi (bf16) != 0
//   outputs (all optional): g fp32, g split, colsum[n] += sum_m g[m,n] (atomic, pre-zeroed)
int act_dropout_bwd(const float* dy, int64_t lddy, int64_t M, int N, int act, const float* aux, int64_t ldaux,
                    const bf16_t* aux_hi, int64_t ldaux_hi, Drop drop, float* g, int64_t ldg, bf16_t* g_hi,
                    bf16_t* g_lo, int64_t ldg_split, float* colsum, stream_t s);

// out[((m / div) % mod), n] += x[m, n]   (atomic; out pre-zeroed by caller, ld = N)
int row_reduce_mod(const float* x, int64_t ldx, int64_t M, int N, int div, int mod, float* out, stream_t s);
// dst[m, n] = src[m / div, n] for m in [0, M): fp32 and/or split copies
int broadcast_rows(const float* src, int64_t lds, int64_t M, int N, int div, float* dst, int64_t ldd, bf16_t* d_hi,
                   bf16_t* d_lo, int64_t ldd_split, stream_t s);

// y[r, :] = tanh(a[r, 0:7] W^T + b + E[r % T])   (E may be null)  -> fp32 + split
int embed_action_fwd(const float* actions, int64_t R, int A, int H, const float* W, const float* b, const float* E, int T,
                     float* y, bf16_t* y_hi, bf16_t* y_lo, stream_t s);
// dpre = dy * (1 - y^2); dW[H,A] += dpre^T a; db += colsum(dpre); dE[r % T] += dpre   (all accumulated, pre-zeroed)
int embed_action_bwd(const float* dy, const float* y, const float* actions, int64_t R, int A, int H, int T, float* dW,
                     float* db, float* dE, stream_t s);

// narrow head: out[r, c] = x[r,:] . W[c,:] + b[c], C small (5)
// fused training loss (see include/videocad_b200.h, vc_loss_*)
typedef vc_loss_cfg LossCfg;
size_t loss_workspace_floats(int R, int NP);
int loss_forward(const LossCfg& cfg, const float* cmds, const float* params, const float* targets, float* ws, float* loss_out, stream_t s);
int loss_backward(const LossCfg& cfg, const float* cmds, const float* params, const float* targets, const float* ws,
                  const float* upstream, float* dcmds, float* dparams, stream_t s);
// metrics of compute_loss from the argmax predictions loss_forward left in ws (see vc_loss_metrics in include/videocad_b200.h)
int loss_metrics(const LossCfg& cfg, const vc_metrics_cfg& mc, const float* targets, const float* ws, int T, int64_t* counts, stream_t s);

// fused clip_grad_norm_ + Adam step (see include/videocad_b200.h, vc_clip_adam_step)
size_t clip_adam_scratch_floats();
int clip_adam_step(const vc_adam_tensor* tensors, int num_tensors, double beta1, double beta2, double eps, double max_norm, int64_t step,
                   float* scratch, float* total_norm_out, stream_t s);

int head_small_fwd(const float* x, int64_t R, int H, const float* W, const float* b, int C, float* out, stream_t s);
// dx[r,:] (+)= dout[r,:] W ; dW += dout^T x ; db += colsum(dout)   (dW/db accumulated, pre-zeroed)
int head_small_bwd(const float* dout, const float* x, int64_t R, int H, const float* W, int C, float* dx, int accumulate_dx,
                   float* dW, float* db, stream_t s);

// fork/join onto auxiliary streams, so that small independent kernels (decoder-sized wgrad GEMMs, K/V projections)
// can run next to the critical chain instead of serialising behind it.  Works under CUDA-graph capture (event edges).
//   stream_fork: auxiliary stream `i` (0..3) first waits for everything enqueued so far on `main`; returned in *side
//   stream_join: `main` waits for everything enqueued so far on auxiliary stream `i`
// The CPU emulation returns `main` itself (everything is serial there).
int stream_fork(stream_t main, int i, stream_t* side);
int stream_join(stream_t main, int i);
// 0: stream_fork hands back the caller's stream (serialised execution, used when timing kernels one by one); 1: default
void side_streams_enable(int enable);

// Linear layer on a HANDFUL of rows (M <= 16: one token per sequence of an incremental decoding step), exact fp32 SIMT:
//   out[m, n] = act(sum_k x[m, k] * W[n, k] + bias[n]) + residual[m, n]      W fp32 [N, K] row-major (nn.Linear.weight)
// x is fp32 (x != null) or split-bf16 (x_hi / x_lo, joined on the fly); outputs fp32 and/or split.  A 128-row tensor-core tile
// would be 94 % padding here and each of its CTAs would stream its weight slab alone; this kernel spreads the N x K weight
// matrix over all SMs (one warp per output column) so a step is bound by reading the weights once.
int linear_rows_fwd(const float* x, const bf16_t* x_hi, const bf16_t* x_lo, int64_t ldx, int M, const float* W, const float* bias, int N,
                    int K, int act, const float* residual, int64_t ld_res, float* out_f32, int64_t ldo, bf16_t* out_hi, bf16_t* out_lo,
                    int64_t ldo_split, stream_t s);

// ---- one position of the key/value-cached rollout for B <= 16 sequences (decode.cu; CPU twin in oracle/csrc/kernels_cpu.cpp) ----
// out[m, n] = act(sum_k x[m, k] W[n, k] + bias[n]) + residual[m, n], stored at out[m * out_row_stride + t * out_t_stride + n]
// with t = *t_ptr (0 if null).  The input rows x [M, K] are built on load according to in_mode:
//   PLAIN  x = rows of `x` (ldx)                                   LN     x = LayerNorm(rows of `x`; gamma, beta, eps 1e-5)
//   EMBED  x = tanh(actions[M, act_dim] emb_W^T + emb_b + emb_E[t])  (K = hidden size, emb_W [K, act_dim], emb_E [*, K] or null)
// x_out (optional, [M, K]): receives the rows that were built (the residual operand of a later call).
enum { VC_DEC_IN_PLAIN = 0, VC_DEC_IN_LN = 1, VC_DEC_IN_EMBED = 2 };
struct DecGemv {
  int in_mode;
  const float* x; long long ldx;
  const float* gamma; const float* beta;
  float* x_out;
  const float* actions; int act_dim; const float* emb_W; const float* emb_b; const float* emb_E;
  int M, N, K;
  const float* W; const float* bias;
  int act;
  const float* residual; long long ld_res;
  float* out; long long out_row_stride; long long out_t_stride;
  const int* t_ptr;
  int cols_per_cta;  // chosen by the launcher
};
int dec_gemv(const DecGemv& g, stream_t s);
// attention of ONE query row per (sequence b, head h) at position t = *t_ptr against keys j in [0, t] (window == 0) or
// (t - window, t] (window > 0):  q row = q + b*q_bstride + t*q_tstride, key/value row j = k|v + b*kv_bstride + j*kv_rstride, head h =
// columns [h*dh, (h+1)*dh).  out[b*ld_out + h*dh + d] receives softmax(q k^T) v.  The keys may be split over nsplit parts that run
// concurrently; they meet through part_o [B, nh, nsplit, dh] / part_ml [B, nh, nsplit, 2] / counters [B, nh] (zero before the first
// call, left zero by every call).
struct DecAttn {
  const float* q; long long q_bstride; long long q_tstride;
  const float* k; const float* v; long long kv_bstride; long long kv_rstride;
  int nh, dh, nsplit, window;
  float scale;
  const int* t_ptr;
  float* out; long long ld_out;
  float* part_o; float* part_ml; unsigned int* counters;
};
int dec_attn(const DecAttn& a, int B, stream_t s);
// end of a step: x = LayerNorm(y[b]; gamma, beta); cmds_all[b, t, :] = x Wc^T + bc; argmax of the command logits and of the NPAR
// parameter heads (params_all[b, t, i, :], already written); apply_action_mask + normalize_actions
// (/root/reference/model/autoregressive_transformer.py:91-118) -> action_next[b, 0:1+NPAR]; finally *t_ptr += 1.
struct DecSelect {
  const float* y; const float* gamma; const float* beta;
  const float* Wc; const float* bc; int NC;
  int H, B, NPAR, NV, T;
  float* cmds_all; float* params_all;
  float* action_next;
  int* t_ptr; unsigned int* done_ctr;
};
int dec_select(const DecSelect& a, stream_t s);

// frame ingestion: uint8 grey-level frames -> the fp32 tensor the reference's loader hands the model,
// i.e. torchvision ToTensor (u / 255) followed by Normalize(mean, std) (/root/reference/main.py:103-110: mean = std = 0.5):
// dst[i] = (float(src[i]) / 255 - mean) / std, evaluated with the same fp32 operations in the same order (bit-exact)
int frames_u8_normalize(const uint8_t* src, int64_t n, float mean, float std, float* dst, stream_t s);
// uint8 RGB frames [n, Hin, Win, 3] (as stored in the dataset) -> fp32 [n, 1, Hout, Wout]: the whole frame transform of the reference's
// loader, Resize((Hout, Wout)) -> Grayscale(1) -> ToTensor -> Normalize(mean, std) (main.py:103-108), bit-exact with Pillow's
// 8-bit arithmetic (ingest.cu).  kk_* / bounds_*: Pillow's fixed-point resampling coefficients [out, ks] and (first tap, tap count)
// pairs [out, 2] per axis, device-resident, needed only for an axis whose size changes; tmp: uint8 [n, Hin, Wout, 3], needed only
// when the width changes.
int frames_rgb_u8_ingest(const uint8_t* src, int64_t n, int Hin, int Win, int Hout, int Wout, const int* kk_h, const int* bounds_h, int ks_h,
                         const int* kk_v, const int* bounds_v, int ks_v, uint8_t* tmp, float mean, float std, float* dst, stream_t s);

// misc
int add_f32(const float* a, const float* b, float* out, int64_t n, stream_t s);  // out = a + b (b may alias out)
int zero_f32(float* x, int64_t n, stream_t s);
int dropout_mask_debug(Drop drop, int64_t n, float* out, stream_t s);  // out[i] = keep ? 1/(1-p) : 0 (tests only)

}  // namespace vck
