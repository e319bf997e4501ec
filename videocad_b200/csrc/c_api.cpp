// extern "C" surface of libvideocad_b200 (declared in include/videocad_b200.h).  Thin forwarding layer: every
// entry point maps 1:1 onto a launcher of kernels.h or onto a model-level routine of model_*.cpp.
#include "videocad_b200.h"
#include "kernels.h"
#include "host_util.h"

#ifndef VC_CUDA_BUILD
#define VC_CUDA_BUILD 0
#endif

extern "C" {

const char* vc_last_error(void) { return vck::last_error(); }
int vc_version(void) { return 100; }
int vc_is_cuda_build(void) { return VC_CUDA_BUILD; }
long long vc_launch_count(void) { return vck::launch_count(); }
void vc_launch_count_reset(void) { vck::launch_count_reset(); }
long long vc_gemm_pair_launch_count(void) { return vck::pair_launch_count(); }
void vc_side_streams_enable(int enable) { vck::side_streams_enable(enable); }
void vc_gemm_profile(int enable) { vck::gemm_profile_enable(enable); }
int vc_gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  return vck::gemm_profile_read(total_ms, total_flops, launches);
}
int vc_gemm_profile_read_min(double min_flops, double* total_ms, double* total_flops, long long* launches) {
  return vck::gemm_profile_read_min(min_flops, total_ms, total_flops, launches);
}
int vc_gemm_profile_dump(const char* path) { return vck::gemm_profile_dump(path); }
size_t vc_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(vc_drop);
    case 1: return sizeof(vc_gemm_desc);
    case 2: return sizeof(vc_attn_desc);
    case 3: return sizeof(vc_linear);
    case 4: return sizeof(vc_norm);
    case 5: return sizeof(vc_vit_weights);
    case 6: return sizeof(vc_vit_call);
    case 7: return sizeof(vc_dec_layer);
    case 8: return sizeof(vc_seq_weights);
    case 9: return sizeof(vc_seq_call);
    default: return 0;
  }
}

void vc_gemm_desc_init(vc_gemm_desc* d) { vck::gemm_desc_init(d); }
void vc_gemm_pair_force_tile(int bn) { vck::gemm_pair_force_tile(bn); }
int vc_gemm(const vc_gemm_desc* d, void* stream) {
  if (!d) return vck::set_error("vc_gemm: null descriptor");
  return vck::gemm(*d, stream);
}
int vc_split_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, vc_bf16* hi, vc_bf16* lo, int64_t ldo, void* stream) {
  return vck::split_f32(x, ldx, rows, cols, hi, lo, ldo, stream);
}
int vc_split_many(const vc_split_item* items, int n_items, int64_t total_blocks, void* stream) {
  return vck::split_many(items, n_items, total_blocks, stream);
}
int vc_layernorm_fwd(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                     float* y, int64_t ldy, vc_bf16* y_hi, vc_bf16* y_lo, int64_t ldy_split, float* mean, float* rstd,
                     void* stream) {
  return vck::layernorm_fwd(x, ldx, rows, C, gamma, beta, eps, y, ldy, y_hi, y_lo, ldy_split, mean, rstd, stream);
}
int vc_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                     const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                     float* dgamma, float* dbeta, void* stream) {
  return vck::layernorm_bwd(dy, lddy, x, ldx, mean, rstd, gamma, rows, C, dres, lddres, dx, lddx, dgamma, dbeta, stream);
}
int vc_layernorm_bwd_fused(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                           const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                           float* dgamma, float* dbeta, vc_drop gdrop, vc_bf16* g_hi, vc_bf16* g_lo, int64_t ldg,
                           float* g_colsum, void* stream) {
  return vck::layernorm_bwd_fused(dy, lddy, x, ldx, mean, rstd, gamma, rows, C, dres, lddres, dx, lddx, dgamma, dbeta, gdrop, g_hi,
                                  g_lo, ldg, g_colsum, stream);
}
int vc_attention_bwd_split(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                           const float* dout, const vc_bf16* dout_hi, const vc_bf16* dout_lo, int64_t lddo, float* scratch,
                           vc_bf16* dq_hi, vc_bf16* dq_lo, vc_bf16* dk_hi, vc_bf16* dk_lo, vc_bf16* dv_hi, vc_bf16* dv_lo,
                           int64_t ld_split, void* stream) {
  if (!a) return vck::set_error("vc_attention_bwd_split: null descriptor");
  return vck::attention_bwd_split(*a, o_hi, o_lo, ldo, lse, dout, dout_hi, dout_lo, lddo, scratch, dq_hi, dq_lo, dk_hi, dk_lo, dv_hi,
                                  dv_lo, ld_split, stream);
}
int vc_attention_bwd_split_bias(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                                const float* dout, int64_t lddo, float* scratch, vc_bf16* dq_hi, vc_bf16* dq_lo, vc_bf16* dk_hi,
                                vc_bf16* dk_lo, vc_bf16* dv_hi, vc_bf16* dv_lo, int64_t ld_split, float* dbq, float* dbk, float* dbv,
                                void* stream) {
  if (!a) return vck::set_error("vc_attention_bwd_split_bias: null descriptor");
  return vck::attention_bwd_split_bias(*a, o_hi, o_lo, ldo, lse, dout, lddo, scratch, dq_hi, dq_lo, dk_hi, dk_lo, dv_hi, dv_lo, ld_split,
                                       dbq, dbk, dbv, stream);
}
void vc_attention_small_enable(int enable) { vck::attention_small_enable(enable); }
int vc_patch_layernorm_fwd(const float* img, int F, int S, const float* gamma, const float* beta, float eps,
                           vc_bf16* y_hi, vc_bf16* y_lo, float* mean, float* rstd, void* stream) {
  return vck::patch_layernorm_fwd(img, F, S, gamma, beta, eps, y_hi, y_lo, mean, rstd, stream);
}
int vc_patch_layernorm_bwd_params(const float* img, int F, int S, const float* mean, const float* rstd, const float* dy,
                                  float* dgamma, float* dbeta, void* stream) {
  return vck::patch_layernorm_bwd_params(img, F, S, mean, rstd, dy, dgamma, dbeta, stream);
}
int vc_vit_assemble_fwd(const float* e, int F, int N, int C, const float* cls, const float* pos, vc_drop drop, float* x,
                        void* stream) {
  return vck::vit_assemble_fwd(e, F, N, C, cls, pos, drop, x, stream);
}
int vc_vit_assemble_bwd(const float* dx, int F, int N, int C, vc_drop drop, float* de, float* dcls, float* dpos,
                        void* stream) {
  return vck::vit_assemble_bwd(dx, F, N, C, drop, de, dcls, dpos, stream);
}
int vc_attention_fwd(const vc_attn_desc* a, vc_bf16* o_hi, vc_bf16* o_lo, int64_t ldo, float* lse, void* stream) {
  if (!a) return vck::set_error("vc_attention_fwd: null descriptor");
  return vck::attention_fwd(*a, o_hi, o_lo, ldo, lse, stream);
}
int vc_attention_bwd(const vc_attn_desc* a, const vc_bf16* o_hi, const vc_bf16* o_lo, int64_t ldo, const float* lse,
                     const float* dout, int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                     int64_t lddv, void* stream) {
  if (!a) return vck::set_error("vc_attention_bwd: null descriptor");
  return vck::attention_bwd(*a, o_hi, o_lo, ldo, lse, dout, lddo, dq, lddq, dk, lddk, dv, lddv, stream);
}
int vc_act_dropout_bwd(const float* dy, int64_t lddy, int64_t M, int N, int act, const float* aux, int64_t ldaux,
                       const vc_bf16* aux_hi, int64_t ldaux_hi, vc_drop drop, float* g, int64_t ldg, vc_bf16* g_hi,
                       vc_bf16* g_lo, int64_t ldg_split, float* colsum, void* stream) {
  return vck::act_dropout_bwd(dy, lddy, M, N, act, aux, ldaux, aux_hi, ldaux_hi, drop, g, ldg, g_hi, g_lo, ldg_split, colsum,
                              stream);
}
int vc_row_reduce_mod(const float* x, int64_t ldx, int64_t M, int N, int div, int mod, float* out, void* stream) {
  return vck::row_reduce_mod(x, ldx, M, N, div, mod, out, stream);
}
int vc_broadcast_rows(const float* src, int64_t lds, int64_t M, int N, int div, float* dst, int64_t ldd, vc_bf16* d_hi,
                      vc_bf16* d_lo, int64_t ldd_split, void* stream) {
  return vck::broadcast_rows(src, lds, M, N, div, dst, ldd, d_hi, d_lo, ldd_split, stream);
}
int vc_embed_action_fwd(const float* actions, int64_t R, int A, int H, const float* W, const float* b, const float* E,
                        int T, float* y, vc_bf16* y_hi, vc_bf16* y_lo, void* stream) {
  return vck::embed_action_fwd(actions, R, A, H, W, b, E, T, y, y_hi, y_lo, stream);
}
int vc_embed_action_bwd(const float* dy, const float* y, const float* actions, int64_t R, int A, int H, int T, float* dW,
                        float* db, float* dE, void* stream) {
  return vck::embed_action_bwd(dy, y, actions, R, A, H, T, dW, db, dE, stream);
}
size_t vc_loss_workspace_floats(int R, int NP) { return vck::loss_workspace_floats(R, NP); }
int vc_loss_forward(const vc_loss_cfg* cfg, const float* cmds, const float* params, const float* targets, float* ws, float* loss_out,
                    void* stream) {
  if (!cfg) return vck::set_error("vc_loss_forward: null cfg");
  return vck::loss_forward(*cfg, cmds, params, targets, ws, loss_out, stream);
}
int vc_loss_backward(const vc_loss_cfg* cfg, const float* cmds, const float* params, const float* targets, const float* ws,
                     const float* upstream, float* dcmds, float* dparams, void* stream) {
  if (!cfg) return vck::set_error("vc_loss_backward: null cfg");
  return vck::loss_backward(*cfg, cmds, params, targets, ws, upstream, dcmds, dparams, stream);
}
int vc_loss_metrics(const vc_loss_cfg* cfg, const vc_metrics_cfg* mcfg, const float* targets, const float* ws, int T, int64_t* counts,
                    void* stream) {
  if (!cfg || !mcfg) return vck::set_error("vc_loss_metrics: null cfg");
  return vck::loss_metrics(*cfg, *mcfg, targets, ws, T, counts, stream);
}

size_t vc_clip_adam_scratch_floats(void) { return vck::clip_adam_scratch_floats(); }
int vc_clip_adam_step(const vc_adam_tensor* tensors, int num_tensors, double beta1, double beta2, double eps, double max_norm,
                      int64_t step, float* scratch, float* total_norm_out, void* stream) {
  return vck::clip_adam_step(tensors, num_tensors, beta1, beta2, eps, max_norm, step, scratch, total_norm_out, stream);
}

int vc_head_small_fwd(const float* x, int64_t R, int H, const float* W, const float* b, int C, float* out, void* stream) {
  return vck::head_small_fwd(x, R, H, W, b, C, out, stream);
}
int vc_head_small_bwd(const float* dout, const float* x, int64_t R, int H, const float* W, int C, float* dx,
                      int accumulate_dx, float* dW, float* db, void* stream) {
  return vck::head_small_bwd(dout, x, R, H, W, C, dx, accumulate_dx, dW, db, stream);
}
int vc_linear_rows_fwd(const float* x, const vc_bf16* x_hi, const vc_bf16* x_lo, int64_t ldx, int M, const float* W, const float* bias, int N,
                       int K, int act, const float* residual, int64_t ld_res, float* out_f32, int64_t ldo, vc_bf16* out_hi, vc_bf16* out_lo,
                       int64_t ldo_split, void* stream) {
  return vck::linear_rows_fwd(x, x_hi, x_lo, ldx, M, W, bias, N, K, act, residual, ld_res, out_f32, ldo, out_hi, out_lo, ldo_split, stream);
}
int vc_frames_u8_normalize(const uint8_t* src, int64_t n, float mean, float std, float* dst, void* stream) {
  return vck::frames_u8_normalize(src, n, mean, std, dst, stream);
}
int vc_frames_rgb_u8_ingest(const uint8_t* src, int64_t n, int Hin, int Win, int Hout, int Wout, const int* kk_h, const int* bounds_h, int ks_h,
                            const int* kk_v, const int* bounds_v, int ks_v, uint8_t* tmp, float mean, float std, float* dst, void* stream) {
  return vck::frames_rgb_u8_ingest(src, n, Hin, Win, Hout, Wout, kk_h, bounds_h, ks_h, kk_v, bounds_v, ks_v, tmp, mean, std, dst, stream);
}
int vc_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) { return vck::add_f32(a, b, out, n, stream); }
int vc_zero_f32(float* x, int64_t n, void* stream) { return vck::zero_f32(x, n, stream); }
int vc_dropout_mask_debug(vc_drop drop, int64_t n, float* out, void* stream) {
  return vck::dropout_mask_debug(drop, n, out, stream);
}

}  // extern "C"
