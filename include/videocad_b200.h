/* videocad_b200 -- C ABI of the B200-native (sm_100a) VideoCAD behaviour-cloning hot path.
 *
 * Drop-in boundary: the object returned by the reference's ModelFactory.create_model
 * (/root/reference/model/model_factory.py:15-36) whose forward is
 * AutoRegressiveTransformer.forward (/root/reference/model/autoregressive_transformer.py:121-220).
 * The reference is pure Python/PyTorch and has no FFI of its own; the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (videocad_b200/lib.py is that stub).
 *
 * Rules of this ABI
 *   - plain pointers and sizes only; all pointers are DEVICE pointers on the current CUDA device unless
 *     stated otherwise; `stream` is a cudaStream_t passed as void*;
 *   - no function allocates device memory or synchronises the host; outputs/workspaces are caller-allocated;
 *   - every function returns 0 on success, non-zero on error (message via vc_last_error());
 *   - matrices are row-major with explicit leading dimensions in ELEMENTS;
 *   - "split" operands are pairs of bf16 arrays (hi, lo) with x ~= hi + lo, hi = bf16_rn(x),
 *     lo = bf16_rn(x - hi); bf16 values are passed as uint16_t bit patterns.
 */
#ifndef VIDEOCAD_B200_H_
#define VIDEOCAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t vc_bf16;

enum { VC_ACT_NONE = 0, VC_ACT_GELU = 1, VC_ACT_RELU = 2, VC_ACT_TANH = 3,
       /* GEMM epilogues only.  VC_ACT_GELU_DSTORE (forward, needs preact): like VC_ACT_GELU, but `preact` receives
        * D = gelu'(v) * dropout_mask * dropout_scale instead of the pre-activation v -- the factor the backward multiplies by,
        * computed where erf(v) and the mask exist anyway.  VC_ACT_MUL_AUX (act_backward, needs act_aux): v *= act_aux[m,n],
        * i.e. the dgrad epilogue reads that D and needs neither erf nor the dropout hash. */
       VC_ACT_GELU_DSTORE = 4, VC_ACT_MUL_AUX = 5 };
enum { VC_MASK_NONE = 0, VC_MASK_CAUSAL = 1, VC_MASK_WINDOW = 2 };

/* one dropout call site: keep-mask is a pure function of (seed, site, element index); p == 0 disables.
 * If seed_ptr != NULL the kernels read the seed from that DEVICE location instead of `seed` (lets a captured CUDA graph
 * be replayed with a fresh seed). */
typedef struct vc_drop {
  float p;
  uint32_t site;
  uint64_t seed;
  const uint64_t* seed_ptr;
} vc_drop;

/* Tensor-core GEMM  acc[m,n] = sum_k A[m,k] * B[n,k]  with fused epilogue.
 * Replaces every nn.Linear GEMM (and its dgrad/wgrad) on the path; see csrc/gemm_tc.cu.
 *   A: a_mn_major == 0 -> stored [M,K] (lda >= K);  == 1 -> stored [K,M] (lda >= M)
 *   B: b_mn_major == 0 -> stored [N,K] (ldb >= K);  == 1 -> stored [K,N] (ldb >= N)
 *   passes: 3 = hi*hi + lo*hi + hi*lo (fp32-grade), 1 = hi*hi (bf16-grade)
 * epilogue order: +bias[n], +rowadd[(m/rowadd_div)%rowadd_mod, n], store preact, act, dropout, +residual,
 *   store out_f32 (atomicAdd when splitk > 1) and/or split (out_hi, out_lo).
 * backward-activation mode (act_backward != 0; used by dgrad GEMMs): instead of "act, dropout" the epilogue applies
 *   v = dropout_mask(v) * act'(.)  with act' taken from act_aux (GELU: pre-activation fp32; TANH: forward output fp32;
 *   MUL_AUX: the stored factor itself) or act_aux_hi (RELU: forward output hi != 0), and colsum[n] += sum_m v[m,n]
 *   (atomic, caller-zeroed) if colsum != NULL. */
typedef struct vc_gemm_desc {
  const vc_bf16 *a_hi, *a_lo; int64_t lda; int a_mn_major;
  const vc_bf16 *b_hi, *b_lo; int64_t ldb; int b_mn_major;
  int M, N, K, passes, splitk;
  const float* bias;
  const float* rowadd; int64_t ld_rowadd; int rowadd_div, rowadd_mod;
  float* preact; int64_t ld_preact;
  int act;
  vc_drop drop;
  const float* residual; int64_t ld_res;
  float* out_f32; int64_t ldo;
  vc_bf16 *out_hi, *out_lo; int64_t ldo_split;
  int act_backward;
  const float* act_aux; int64_t ld_act_aux;
  const vc_bf16* act_aux_hi; int64_t ld_act_aux_hi;
  float* colsum;
} vc_gemm_desc;

/* Multi-head attention core (replaces the bmm/softmax/dropout/bmm chain of vit_pytorch Attention and of
 * F.multi_head_attention_forward).  Rows are (b*T + t); head h owns columns [h*d, (h+1)*d). */
typedef struct vc_attn_desc {
  const float *q, *k, *v; int64_t ldq, ldk, ldv;
  /* optional split-bf16 inputs (as written by a GEMM epilogue): when q_hi != NULL they replace q/k/v and ldq/ldk/ldv are
   * the leading dimensions of the bf16 arrays */
  const vc_bf16 *q_hi, *q_lo, *k_hi, *k_lo, *v_hi, *v_lo;
  int B, Tq, Tk, nh, d;
  int mask, window;
  float scale;
  vc_drop drop;
  /* optional batch strides (elements) of q / k / v: sample b starts at row offset b * bs instead of b * T * ld.  0 = dense
   * (b * Tq * ldq, b * Tk * ldk, b * Tk * ldv).  Non-zero strides are accepted by the forward only (fp32 inputs): they let a
   * decoding step attend from ONE query row per sequence to the first Tk rows of a [B, Tmax, .] key/value cache. */
  int64_t bsq, bsk, bsv;
} vc_attn_desc;

const char* vc_last_error(void);
int vc_version(void);
/* 1 if this library was built from the CUDA sources (product), 0 for the CPU emulation used by host-logic tests */
int vc_is_cuda_build(void);

/* ---- instrumentation (bench.py): number of kernels launched by this library, and CUDA-event timing of every
 * tensor-core GEMM launch (events recorded on the launching stream around each launch while enabled) ---- */
long long vc_launch_count(void);
void vc_launch_count_reset(void);
long long vc_gemm_pair_launch_count(void);           /* GEMM launches that used the 2-SM (cta_group::2) 256x256 kernel */
void vc_side_streams_enable(int enable);           /* 0: run auxiliary-stream work on the caller's stream (per-kernel timing); default 1 */
void vc_gemm_profile(int enable);                  /* enable/disable + clear the event pool */
int vc_gemm_profile_read(double* total_ms, double* total_flops, long long* launches); /* synchronises the events */
int vc_gemm_profile_read_min(double min_flops, double* total_ms, double* total_flops, long long* launches); /* only launches >= min_flops */
int vc_gemm_profile_dump(const char* path);       /* per-launch CSV: tag (shape/majors/split/epilogue), flops, ms */
/* sizeof() of the ABI structs, for binding self-checks: 0 vc_drop, 1 vc_gemm_desc, 2 vc_attn_desc, 3 vc_linear,
 * 4 vc_norm, 5 vc_vit_weights, 6 vc_vit_call, 7 vc_dec_layer, 8 vc_seq_weights, 9 vc_seq_call */
size_t vc_abi_sizeof(int which);

/* ---- per-kernel entry points (unit parity tests call these; the model-level calls below chain them) ---- */
void vc_gemm_desc_init(vc_gemm_desc* d);
int vc_gemm(const vc_gemm_desc* d, void* stream);
/* tile width of the 2-SM (cta_group::2) GEMM kernel: 0 = chosen per problem (default), 128 / 256 = forced where N allows (parity
 * tests of both widths, experiments) */
void vc_gemm_pair_force_tile(int bn);
int vc_split_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, vc_bf16* hi, vc_bf16* lo, int64_t ldo, void* stream);
/* many fp32 -> split conversions in ONE launch (all weight matrices of a segment after an optimizer step).
 * `items` is a DEVICE array; item i converts n4 float4 groups and owns blocks [block_start, block_start + ceil(n4/1024)). */
typedef struct vc_split_item {
  const float* src; vc_bf16* hi; vc_bf16* lo;
  int64_t n4; int64_t block_start;
} vc_split_item;
int vc_split_many(const vc_split_item* items, int n_items, int64_t total_blocks, void* stream);
int vc_layernorm_fwd(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                     float* y, int64_t ldy, vc_bf16* y_hi, vc_bf16* y_lo, int64_t ldy_split, float* mean, float* rstd,
                     void* stream);
int vc_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                     const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                     float* dgamma, float* dbeta, void* stream);
/* layernorm_bwd + fused second output g = dx * dropout_mask (split-bf16) and its column sums (see csrc/kernels.h) */
int vc_layernorm_bwd_fused(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                           const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                           float* dgamma, float* dbeta, vc_drop gdrop, vc_bf16* g_hi, vc_bf16* g_lo, int64_t ldg,
                           float* g_colsum, void* stream);
/* attention backward delivering dq/dk/dv as split-bf16 GEMM operands (scratch: 3*B*T*nh*d floats, may be unused);
 * the upstream gradient is either fp32 (dout) or split-bf16 (dout_hi/dout_lo, when dout == NULL), leading dimension lddo */
int vc_attention_bwd_split(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                           const float* dout, const vc_bf16* dout_hi, const vc_bf16* dout_lo, int64_t lddo, float* scratch, vc_bf16* dq_hi, vc_bf16* dq_lo, vc_bf16* dk_hi,
                           vc_bf16* dk_lo, vc_bf16* dv_hi, vc_bf16* dv_lo, int64_t ld_split, void* stream);
/* the same for attention behind a packed in_proj WITH bias (torch nn.MultiheadAttention, F.multi_head_attention_forward's
 * in_proj_bias; reference: TransformerDecoderLayer built at model/autoregressive_transformer.py:54-62): fp32 upstream gradient,
 * and the bias gradient (column sums of dq / dk / dv) accumulated into dbq / dbk / dbv ([nh*d] each, caller-zeroed, may be NULL).
 * Sequences with Tq == Tk <= 32 take one fused launch (csrc/attention_small.cu). */
int vc_attention_bwd_split_bias(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                                const float* dout, int64_t lddo, float* scratch, vc_bf16* dq_hi, vc_bf16* dq_lo, vc_bf16* dk_hi,
                                vc_bf16* dk_lo, vc_bf16* dv_hi, vc_bf16* dv_lo, int64_t ld_split, float* dbq, float* dbk, float* dbv,
                                void* stream);
/* 0: short-sequence attention (Tq == Tk <= 32) runs on the generic kernels instead of csrc/attention_small.cu (parity tests, A/B
 * timing); default 1 */
void vc_attention_small_enable(int enable);
int vc_patch_layernorm_fwd(const float* img, int F, int S, const float* gamma, const float* beta, float eps,
                           vc_bf16* y_hi, vc_bf16* y_lo, float* mean, float* rstd, void* stream);
int vc_patch_layernorm_bwd_params(const float* img, int F, int S, const float* mean, const float* rstd, const float* dy,
                                  float* dgamma, float* dbeta, void* stream);
int vc_vit_assemble_fwd(const float* e, int F, int N, int C, const float* cls, const float* pos, vc_drop drop, float* x,
                        void* stream);
int vc_vit_assemble_bwd(const float* dx, int F, int N, int C, vc_drop drop, float* de, float* dcls, float* dpos,
                        void* stream);
int vc_attention_fwd(const vc_attn_desc* a, vc_bf16* o_hi, vc_bf16* o_lo, int64_t ldo, float* lse, void* stream);
int vc_attention_bwd(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                     const float* dout, int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                     int64_t lddv, void* stream);
int vc_act_dropout_bwd(const float* dy, int64_t lddy, int64_t M, int N, int act, const float* aux, int64_t ldaux,
                       const vc_bf16* aux_hi, int64_t ldaux_hi, vc_drop drop, float* g, int64_t ldg, vc_bf16* g_hi,
                       vc_bf16* g_lo, int64_t ldg_split, float* colsum, void* stream);
int vc_row_reduce_mod(const float* x, int64_t ldx, int64_t M, int N, int div, int mod, float* out, void* stream);
int vc_broadcast_rows(const float* src, int64_t lds, int64_t M, int N, int div, float* dst, int64_t ldd, vc_bf16* d_hi,
                      vc_bf16* d_lo, int64_t ldd_split, void* stream);
int vc_embed_action_fwd(const float* actions, int64_t R, int A, int H, const float* W, const float* b, const float* E,
                        int T, float* y, vc_bf16* y_hi, vc_bf16* y_lo, void* stream);
int vc_embed_action_bwd(const float* dy, const float* y, const float* actions, int64_t R, int A, int H, int T, float* dW,
                        float* db, float* dE, void* stream);
/* ---- training loss of the reference trainer, fused (MultiClassesTrainer.compute_loss with use_mse=True,
 * /root/reference/trainer.py:935-966 and flexible_cross_entropy :853-916; restated in videocad_b200/loss.py):
 *   loss = 2 * CE_w(cmds, tgt[:,0]; ignore -1, class weights cmd_w)
 *        + sum_i cmd_w[param_to_label[i]] * mean_{selected rows}( -mean_{c in [t, min(t+tol_i-1, NV-1)]} log_softmax(params[:,i,:])[c] )
 * where a row is selected when its target is not -1 and its argmax lies outside the window; an empty selection contributes 0
 * and a NaN term is dropped.  Three launches forward (row pass, finalize), one backward; nothing leaves the device.
 * cmds [R,NC] fp32, params [R,NP,NV] fp32, targets [R, 1+NP] fp32 (integers stored as floats, -1 = ignore).
 * ws: vc_loss_workspace_floats(R, NP) floats kept from forward to backward. */
#define VC_LOSS_MAX_PARAMS 8
#define VC_LOSS_MAX_CLASSES 16
typedef struct vc_loss_cfg {
  int R, NC, NP, NV;
  float cmd_w[VC_LOSS_MAX_CLASSES];     /* class weights of the command head */
  int tolerance[VC_LOSS_MAX_PARAMS];     /* window length of parameter i (the reference's tolerance_i) */
  int param_to_label[VC_LOSS_MAX_PARAMS];
} vc_loss_cfg;
size_t vc_loss_workspace_floats(int R, int NP);
int vc_loss_forward(const vc_loss_cfg* cfg, const float* cmds, const float* params, const float* targets, float* ws, float* loss_out,
                    void* stream);
/* dcmds [R,NC], dparams [R,NP,NV] = upstream[0] * d loss / d logits (upstream: device scalar) */
int vc_loss_backward(const vc_loss_cfg* cfg, const float* cmds, const float* params, const float* targets, const float* ws,
                     const float* upstream, float* dcmds, float* dparams, void* stream);
/* The metrics half of compute_loss (trainer.py:968-1061): integer counts from the argmax predictions that vc_loss_forward left
 * in ws (call it after vc_loss_forward on the same ws).  The reference gathers them with ~30 .item() synchronisations per step;
 * here one single-CTA kernel writes VC_METRIC_COUNT int64 counters to device memory, to be read back with ONE copy:
 *   CORRECT / TOTAL                  correct_predictions / total_predictions
 *   CMD_CORRECTS+k / CMD_COUNTS+k    cmd_corrects[k] / cmd_counts[k], k < NC
 *   PARAM_CORRECTS+i / PARAM_COUNTS+i  param_corrects[i] / param_counts[i], i < NP (parameter i counts only where the command is
 *                                    right; correct = 0 <= pred - target < tolerance[i] if above[i], else |pred - target| < abs_tolerance)
 *   *_TOPK                           the same sums over the first `topk` time steps of every sequence (rows are b * T + t)
 * perfect_sequences / perfect_commands / total_sequences are constant 0 in the reference (trainer.py:1017-1034 is commented out). */
enum {
  VC_METRIC_CORRECT = 0, VC_METRIC_TOTAL = 1,
  VC_METRIC_CMD_CORRECTS = 2, VC_METRIC_CMD_COUNTS = 2 + VC_LOSS_MAX_CLASSES,
  VC_METRIC_PARAM_CORRECTS = 2 + 2 * VC_LOSS_MAX_CLASSES, VC_METRIC_PARAM_COUNTS = 2 + 2 * VC_LOSS_MAX_CLASSES + VC_LOSS_MAX_PARAMS,
  VC_METRIC_CMD_CORRECT_TOPK = 2 + 2 * VC_LOSS_MAX_CLASSES + 2 * VC_LOSS_MAX_PARAMS,
  VC_METRIC_CMD_COUNTS_TOPK, VC_METRIC_PARAM_CORRECT_TOPK, VC_METRIC_PARAM_COUNTS_TOPK,
  VC_METRIC_COUNT
};
typedef struct vc_metrics_cfg {
  int above[VC_LOSS_MAX_PARAMS];        /* trainer.py:829 self.above */
  int tolerance[VC_LOSS_MAX_PARAMS];    /* trainer.py:827 self.tolerances */
  int abs_tolerance;                    /* trainer.py TOLERANCE (3) */
  int topk;                             /* trainer.py:1003 k = 30 */
} vc_metrics_cfg;
int vc_loss_metrics(const vc_loss_cfg* cfg, const vc_metrics_cfg* mcfg, const float* targets, const float* ws, int T, int64_t* counts,
                    void* stream);

/* ---- optimizer step of the reference trainer, fused: torch.nn.utils.clip_grad_norm_(params, max_norm) followed by
 * torch.optim.Adam(...).step() (/root/reference/trainer.py:493-494, Adam defaults betas/eps, no weight decay, no amsgrad)
 * over a few large (flat) tensors: a squared-norm pass, a single-CTA finalize producing the clip coefficient on the device,
 * and one update pass that reads g, p, m, v once and writes g (clipped, as clip_grad_norm_ does), p, m, v once.
 * step is the 1-based Adam step count of this call; lr is per tensor (parameter groups). */
#define VC_ADAM_MAX_TENSORS 16
typedef struct vc_adam_tensor {
  float* p; float* g; float* m; float* v;
  int64_t n;
  float lr;
} vc_adam_tensor;
size_t vc_clip_adam_scratch_floats(void);
int vc_clip_adam_step(const vc_adam_tensor* tensors, int num_tensors, double beta1, double beta2, double eps, double max_norm,
                      int64_t step, float* scratch, float* total_norm_out, void* stream);

int vc_head_small_fwd(const float* x, int64_t R, int H, const float* W, const float* b, int C, float* out, void* stream);
int vc_head_small_bwd(const float* dout, const float* x, int64_t R, int H, const float* W, int C, float* dx,
                      int accumulate_dx, float* dW, float* db, void* stream);
/* nn.Linear on M <= 16 rows (one token per sequence of a decoding step), exact fp32: out = act(x W^T + bias) + residual with W the
 * fp32 [N, K] weight; x fp32 or split-bf16 (exactly one of x / x_hi non-NULL); fp32 and/or split outputs.  Used by
 * vc_seq_decode_step, where a 128-row tensor-core tile would be almost entirely padding. */
int vc_linear_rows_fwd(const float* x, const vc_bf16* x_hi, const vc_bf16* x_lo, int64_t ldx, int M, const float* W, const float* bias, int N,
                       int K, int act, const float* residual, int64_t ld_res, float* out_f32, int64_t ldo, vc_bf16* out_hi, vc_bf16* out_lo,
                       int64_t ldo_split, void* stream);
/* ---- frame ingestion: uint8 grey-level frames -> the normalised fp32 tensor the model consumes.
 * Replaces transforms.ToTensor() + transforms.Normalize([0.5], [0.5]) of the reference's loader (main.py:103-110, applied per
 * frame on the CPU in DatasetBase.__getitem__, data_loader/data_loader.py:434-508) and lets the batch cross PCIe as bytes:
 * dst[i] = (float(src[i]) / 255 - mean) / std with the same fp32 operations in the same order (bit-exact).
 * src and dst 16-byte aligned, n = number of pixels. */
int vc_frames_u8_normalize(const uint8_t* src, int64_t n, float mean, float std, float* dst, void* stream);
/* The whole frame transform of the reference's loader on the device: uint8 RGB frames [n, Hin, Win, 3], exactly as the dataset
 * stores them (generate_dataset.py:194-199), -> the encoder's fp32 input [n, 1, Hout, Wout]; replaces
 * Resize((224, 224)) -> Grayscale(1) -> ToTensor() -> Normalize([0.5], [0.5]) (main.py:103-108; per frame, through PIL, in
 * DatasetBase.__getitem__, data_loader/data_loader.py:434-446).  Bit-exact with Pillow's 8-bit resampling / rgb2l arithmetic.
 * kk_h / kk_v: Pillow's fixed-point bilinear coefficients [out, ks] (22 fractional bits), bounds_*: (first tap, tap count) [out, 2],
 * both on the device (videocad_b200.ingest.resample_coeffs computes them); an axis whose size does not change needs none.
 * tmp: uint8 [n, Hin, Wout, 3] when Win != Wout.  src 4-byte aligned, dst 16-byte aligned. */
int vc_frames_rgb_u8_ingest(const uint8_t* src, int64_t n, int Hin, int Win, int Hout, int Wout, const int* kk_h, const int* bounds_h, int ks_h,
                            const int* kk_v, const int* bounds_v, int ks_v, uint8_t* tmp, float mean, float std, float* dst, void* stream);
int vc_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);
int vc_zero_f32(float* x, int64_t n, void* stream);
int vc_dropout_mask_debug(vc_drop drop, int64_t n, float* out, void* stream);


/* =====================================================================================================
 * Model-level entry points: the drop-in path.  Three segments mirror the three autograd nodes of the
 * host-side nn.Module (videocad_b200/model.py):
 *   vc_vit_*   one vit_pytorch.ViT encoder (image_size 224, patch 32, dim 512, depth 6, heads 16 x 64,
 *              mlp 512; reference call sites /root/reference/model/trajectory_model.py:52-67,90-100)
 *   vc_seq_*   token construction + nn.TransformerDecoder (post-norm, ReLU) + the two action heads
 *              (/root/reference/model/autoregressive_transformer.py:143-218)
 * Weight structs carry, per parameter, the fp32 tensor, its split-bf16 copies (vc_split_f32 output,
 * refreshed by the host whenever the fp32 tensor changes) and the gradient destination.
 * Gradient buffers must be ZEROED by the caller before a backward call; backward accumulates into them.
 * ===================================================================================================== */
#define VC_VIT_DEPTH 6
#define VC_VIT_DIM 512
#define VC_VIT_HEADS 16
#define VC_VIT_DHEAD 64
#define VC_VIT_MLP 512
#define VC_PATCH 32

typedef struct vc_linear {      /* y = x W^T + b, W [out, in] row-major */
  const float* w; const float* b;           /* b may be null (to_qkv has no bias) */
  const vc_bf16* w_hi; const vc_bf16* w_lo; /* split copies of w */
  float* dw; float* db;                     /* gradient destinations (null when not training) */
} vc_linear;

typedef struct vc_norm {        /* LayerNorm affine parameters */
  const float* w; const float* b;
  float* dw; float* db;
} vc_norm;

typedef struct vc_vit_layer {
  vc_norm ln1; vc_linear qkv; vc_linear out;
  vc_norm ln2; vc_linear fc1; vc_linear fc2;
} vc_vit_layer;

typedef struct vc_vit_weights {
  const float* pos; const float* cls;   /* pos [>= n, 512] (leading rows used), cls [512] */
  float* dpos; float* dcls;
  vc_norm pe_ln1; vc_linear pe; vc_norm pe_ln2;
  vc_vit_layer layer[VC_VIT_DEPTH];
  vc_norm norm;
} vc_vit_weights;

typedef struct vc_vit_call {
  const vc_vit_weights* w;
  const float* img; int F; int S;       /* [F, 1, S, S], S multiple of 32, (S/32)^2 <= 49 */
  float dropout_p; int training;        /* dropout applied only when training != 0 */
  uint64_t seed; uint32_t site_base;
  const uint64_t* seed_dev;             /* optional device-resident seed (overrides `seed`; for CUDA-graph replay) */
  int passes;                           /* 3 = fp32-grade GEMMs (parity mode), 1 = bf16-grade */
  int aux_streams;                      /* first of the TWO auxiliary-stream indices of the backward's weight-gradient GEMMs:
                                         * 0 = default (2, 3); encoders that run concurrently must use different pairs (4 = 4, 5) */
  void* ws; size_t ws_bytes;            /* activation workspace, >= vc_vit_workspace_bytes(F, S) */
  float* cls_out;                       /* [F, 512] CLS embedding after the final LayerNorm */
} vc_vit_call;

size_t vc_vit_workspace_bytes(int F, int S);
size_t vc_vit_scratch_bytes(int F, int S);
int vc_vit_forward(const vc_vit_call* c, void* stream);
/* needs the SAME call struct (and untouched workspace) as the forward; dcls [F,512] */
int vc_vit_backward(const vc_vit_call* c, const float* dcls, void* scratch, size_t scratch_bytes, void* stream);
/* The same backward in pieces: encoder layers l_hi, l_hi - 1, ..., l_lo (0 <= l_lo <= l_hi <= 5).  The range with l_hi == 5 also runs the
 * final LayerNorm (dcls required), the range with l_lo == 0 the token assembly and the patch embedding.  Ranges called in descending
 * order on the SAME scratch equal one vc_vit_backward; after a range returns, the parameter gradients of its layers are final, so the
 * host can hand them to the gradient all-reduce (DistributedDataParallel, experiment.py:104-109) while the lower layers still run. */
int vc_vit_backward_layers(const vc_vit_call* c, const float* dcls, void* scratch, size_t scratch_bytes, int l_hi, int l_lo, void* stream);

typedef struct vc_dec_layer {
  vc_linear sa_in; vc_linear sa_out;    /* self_attn.in_proj [3H,H], out_proj [H,H] */
  vc_linear ca_in; vc_linear ca_out;    /* multihead_attn */
  vc_linear lin1; vc_linear lin2;       /* [Ff,H], [H,Ff] */
  vc_norm n1; vc_norm n2; vc_norm n3;
} vc_dec_layer;

typedef struct vc_seq_weights {
  vc_linear embed_state; vc_linear embed_image; vc_linear image_proj; vc_linear head_params;
  vc_linear embed_multiview;            /* [H, 512*num_views]; all-null when num_views == 0 */
  const float* embed_action_w; const float* embed_action_b; float* d_embed_action_w; float* d_embed_action_b;
  const float* head_cmd_w; const float* head_cmd_b; float* d_head_cmd_w; float* d_head_cmd_b;
  const float* timestep_emb; float* d_timestep_emb;     /* [max_ep_len, H] or null */
  const vc_dec_layer* layers; int num_layers;
} vc_seq_weights;

typedef struct vc_seq_call {
  const vc_seq_weights* w;
  int B, T, H, nhead, Ff, window;
  int past_actions, past_states;        /* the three branches of forward(), autoregressive_transformer.py:152-213 */
  int act_dim, num_cmd, num_param_out;  /* 7, 5, 6000 */
  const float* state_cls;               /* [B*T, 512] frame embeddings (unused unless past_states) */
  const float* cad_cls;                 /* [B, 512] */
  const float* actions;                 /* [B*T, act_dim] normalised actions */
  int num_views; const float* mv_cls;   /* multiview conditioning: CAD-encoder embeddings of the views, [B*num_views, 512] */
  float dropout_p; int training;
  uint64_t seed; uint32_t site_base;
  const uint64_t* seed_dev;             /* optional device-resident seed (overrides `seed`) */
  int passes;
  void* ws; size_t ws_bytes;            /* >= vc_seq_workspace_bytes(...) */
  float* cmds;                          /* [B*T, num_cmd] */
  float* params;                        /* [B*T, num_param_out] */
} vc_seq_call;

size_t vc_seq_workspace_bytes(int B, int T, int H, int Ff, int num_layers, int nhead, int num_param_out, int num_views);
size_t vc_seq_scratch_bytes(int B, int T, int H, int Ff, int num_param_out, int num_views);
int vc_seq_forward(const vc_seq_call* c, void* stream);
/* Incremental decoding for the rollout with action feedback (sequential_inference, autoregressive_transformer.py:222-275): after ONE
 * vc_seq_forward over the full length T (any actions; eval mode) on call `c`, vc_seq_decode_step(c, t, ...) for t = 0, 1, ... pushes
 * one token per sequence through the decoder, using c->ws as the key/value cache (self-attention rows [0, t], cross-attention
 * window of the memory tokens) instead of re-running the prefix.  actions_t [B, act_dim] normalised actions of step t (zeros at
 * t = 0); cmds_t [B, num_cmd], params_t [B, num_param_out] receive the logits of position t; scratch >=
 * vc_seq_decode_scratch_bytes.  Requires past_actions; results equal vc_seq_forward on the prefix [0, t] at position t. */
size_t vc_seq_decode_scratch_bytes(int B, int H, int Ff, int nhead);
int vc_seq_decode_step(const vc_seq_call* c, int t, const float* actions_t, void* scratch, size_t scratch_bytes, float* cmds_t,
                       float* params_t, void* stream);
/* The same step with the position and the action feedback ON THE DEVICE, for B <= 16 sequences (the rollout of BASELINE config C4:
 * 8 per GPU): *t_dev (device int, 0 before the first step) selects the position and is incremented by the call; actions_io
 * [B, act_dim] (device, zeros before the first step) holds the normalised action of position *t_dev on entry and receives
 * normalize_actions(apply_action_mask(argmax cmd, argmax params)) of this position (autoregressive_transformer.py:91-118, 256-266)
 * on exit; the logits of the position go to row (b, t) of cmds_all [B, T, num_cmd] / params_all [B, T, num_param_out].  The launch
 * sequence does not depend on the position: one captured CUDA graph serves all T steps and the host is never consulted.
 * Every kernel is bound by reading its weights once (TMA bulk copies issued ahead of the dependency wait).
 * scratch (>= vc_seq_decode_dev_scratch_bytes) must be ZERO-FILLED before the first step of the first rollout (it holds the arrival
 * counters of the kernels, which every call leaves at zero again).
 * vc_seq_decode_dev_supported() != 0 tells whether the configuration is covered (otherwise use vc_seq_decode_step). */
int vc_seq_decode_dev_supported(const vc_seq_call* c);
size_t vc_seq_decode_dev_scratch_bytes(int B, int H, int Ff, int nhead);
int vc_seq_decode_step_dev(const vc_seq_call* c, int* t_dev, float* actions_io, void* scratch, size_t scratch_bytes, float* cmds_all,
                           float* params_all, void* stream);
/* dcmds [B*T,num_cmd], dparams [B*T,num_param_out]; d_state_cls [B*T,512] (may be null unless past_states),
 * d_cad_cls [B,512] and d_mv_cls [B*num_views,512] (may be null when num_views == 0) are OVERWRITTEN */
int vc_seq_backward(const vc_seq_call* c, const float* dcmds, const float* dparams, float* d_state_cls, float* d_cad_cls,
                    float* d_mv_cls, void* scratch, size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOCAD_B200_H_ */
