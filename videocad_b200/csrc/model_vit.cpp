// Host orchestration of one ViT encoder (forward + backward) over the kernels of kernels.h.
//
// Restates vit_pytorch.ViT as configured by the reference (image_size 224, patch 32, dim 512, depth 6, heads 16,
// dim_head 64, mlp_dim 512, pool 'cls', mlp_head = Identity; /root/reference/model/trajectory_model.py:52-67),
// called at trajectory_model.py:90-100 for the frame encoder (state_embedding_model) and the CAD-image encoder
// (cad_embedding_model).  Math: SURVEY.md Appendix A.1.
//
// HBM layout (all row-major, M = F*(N+1) token rows, Mp = F*N patch rows):
//   residual stream x[0..6]  fp32 [M,512]           (kept for the LayerNorm backward)
//   GEMM A-operands           split-bf16 [M,K]        (written directly by the producing LN / epilogue / attention)
//   qkv                       split-bf16 [M,3072]     (written by the to_qkv GEMM epilogue, read by the attention kernels)
// Dropout sites (site_base + i): 0 embedding; layer l: 1+4l attention probs, 2+4l to_out, 3+4l MLP hidden, 4+4l MLP out.
#include <stdlib.h>
#include "model_common.h"

namespace vck {

namespace {

// MLP hidden activation: the fc1 epilogue stores gelu'(pre) * dropout mask * scale for the backward (VC_ACT_GELU_DSTORE /
// VC_ACT_MUL_AUX) instead of the pre-activation; VC_GELU_DSTORE=0 keeps the pre-activation and recomputes erf + mask in the fc2
// dgrad epilogue (the two give the same gradients to rounding; A/B switch for measurements, read once per process)
bool gelu_dstore() {
  static const bool on = [] { const char* e = getenv("VC_GELU_DSTORE"); return e ? atoi(e) != 0 : true; }();
  return on;
}

constexpr int D = VC_VIT_DIM;        // 512
constexpr int DI = VC_VIT_HEADS * VC_VIT_DHEAD;  // 1024
constexpr int PD = VC_PATCH * VC_PATCH;          // 1024
constexpr float LN_EPS = 1e-5f;

struct VitWs {
  Split pl; float *pmean, *prstd;
  float* e0; float *emean, *erstd; float* e1;
  float* x[VC_VIT_DEPTH + 1];
  struct Layer {
    float *m1, *r1; Split h1; Split qkv; float* lse; Split o;
    float* x2; float *m2, *r2; Split h2; float* pre1; Split ud;
  } l[VC_VIT_DEPTH];
  float *fmean, *frstd;
};

void vit_carve(Arena& a, int F, int S, VitWs& w) {
  const int N = (S / VC_PATCH) * (S / VC_PATCH), n = N + 1;
  const size_t Mp = (size_t)F * N, M = (size_t)F * n;
  w.pl = a.alloc_split(Mp, PD);
  w.pmean = a.alloc<float>(Mp); w.prstd = a.alloc<float>(Mp);
  w.e0 = a.alloc<float>(Mp * D);
  w.emean = a.alloc<float>(Mp); w.erstd = a.alloc<float>(Mp);
  w.e1 = a.alloc<float>(Mp * D);
  for (int i = 0; i <= VC_VIT_DEPTH; ++i) w.x[i] = a.alloc<float>(M * D);
  for (int l = 0; l < VC_VIT_DEPTH; ++l) {
    VitWs::Layer& L = w.l[l];
    L.m1 = a.alloc<float>(M); L.r1 = a.alloc<float>(M);
    L.h1 = a.alloc_split(M, D);
    L.qkv = a.alloc_split(M, 3 * DI);
    L.lse = a.alloc<float>((size_t)F * VC_VIT_HEADS * n);
    L.o = a.alloc_split(M, DI);
    L.x2 = a.alloc<float>(M * D);
    L.m2 = a.alloc<float>(M); L.r2 = a.alloc<float>(M);
    L.h2 = a.alloc_split(M, D);
    L.pre1 = a.alloc<float>(M * VC_VIT_MLP);
    L.ud = a.alloc_split(M, VC_VIT_MLP);
  }
  w.fmean = a.alloc<float>(F); w.frstd = a.alloc<float>(F);
}

struct VitScratch {
  float *dxa, *dxb, *dh, *dud, *dO, *dqkv;
  Split g, dpre, dqkvS, dOS;
};

void vit_scratch_carve(Arena& a, int F, int S, VitScratch& s) {
  const int N = (S / VC_PATCH) * (S / VC_PATCH), n = N + 1;
  const size_t M = (size_t)F * n;
  s.dxa = a.alloc<float>(M * D);
  s.dxb = a.alloc<float>(M * D);
  s.dh = a.alloc<float>(M * D);
  s.dud = a.alloc<float>(M * VC_VIT_MLP);
  s.dO = a.alloc<float>(M * DI);
  s.dqkv = a.alloc<float>(M * 3 * DI);
  s.g = a.alloc_split(M, D);
  s.dpre = a.alloc_split(M, VC_VIT_MLP);
  s.dqkvS = a.alloc_split(M, 3 * DI);
  s.dOS = a.alloc_split(M, DI);
}

int check_call(const vc_vit_call* c) {
  if (!c || !c->w || !c->img || !c->cls_out || !c->ws) return set_error("vit: null argument");
  if (c->F <= 0) return set_error("vit: F must be positive");
  if (c->S % VC_PATCH != 0 || c->S <= 0) return set_error("vit: image size must be a positive multiple of 32");
  const int N = (c->S / VC_PATCH) * (c->S / VC_PATCH);
  if (N + 1 > 50) return set_error("vit: more than 49 patches (image larger than 224x224) exceeds the positional table");
  if (c->passes != 1 && c->passes != 3) return set_error("vit: passes must be 1 or 3");
  return 0;
}

AttnDesc vit_attn_desc(const VitWs::Layer& L, int F, int n, Drop drop) {
  AttnDesc a = {};
  a.q = a.k = a.v = nullptr;  // q/k/v come as split-bf16, straight from the to_qkv GEMM epilogue
  a.q_hi = L.qkv.hi; a.q_lo = L.qkv.lo;
  a.k_hi = L.qkv.hi + DI; a.k_lo = L.qkv.lo + DI;
  a.v_hi = L.qkv.hi + 2 * DI; a.v_lo = L.qkv.lo + 2 * DI;
  a.ldq = a.ldk = a.ldv = 3 * DI;
  a.B = F; a.Tq = n; a.Tk = n; a.nh = VC_VIT_HEADS; a.d = VC_VIT_DHEAD;
  a.mask = VC_MASK_NONE; a.window = 1;
  a.scale = 0.125f;  // dim_head ** -0.5
  a.drop = drop;
  return a;
}

}  // namespace

size_t vit_workspace_bytes(int F, int S) {
  Arena a(nullptr, 0);
  VitWs w;
  vit_carve(a, F, S, w);
  return a.used();
}

size_t vit_scratch_bytes(int F, int S) {
  Arena a(nullptr, 0);
  VitScratch s;
  vit_scratch_carve(a, F, S, s);
  return a.used();
}

int vit_forward(const vc_vit_call* c, stream_t st) {
  VC_TRY(check_call(c));
  const vc_vit_weights& W = *c->w;
  const int F = c->F, S = c->S, N = (S / VC_PATCH) * (S / VC_PATCH), n = N + 1;
  const int Mp = F * N, M = F * n, P = c->passes;
  Arena arena(c->ws, c->ws_bytes);
  VitWs w;
  vit_carve(arena, F, S, w);
  if (!arena.ok()) return set_error("vit_forward: workspace too small");
  const float p = c->dropout_p;
  const uint32_t sb = c->site_base;

  // patch embedding: gather + LN(1024) -> Linear(1024,512) -> LN(512)
  VC_TRY(patch_layernorm_fwd(c->img, F, S, W.pe_ln1.w, W.pe_ln1.b, LN_EPS, w.pl.hi, w.pl.lo, w.pmean, w.prstd, st));
  {
    GemmDesc d;
    gemm_linear_fwd(d, w.pl, wsplit(W.pe, PD), Mp, D, PD, P);
    d.bias = W.pe.b; d.out_f32 = w.e0; d.ldo = D;
    VC_TRY(gemm(d, st));
  }
  VC_TRY(layernorm_fwd(w.e0, D, Mp, D, W.pe_ln2.w, W.pe_ln2.b, LN_EPS, w.e1, D, nullptr, nullptr, 0, w.emean, w.erstd, st));
  VC_TRY(vit_assemble_fwd(w.e1, F, N, D, W.cls, W.pos, site_drop(p, c->training, c->seed, sb + 0, c->seed_dev), w.x[0], st));

  for (int l = 0; l < VC_VIT_DEPTH; ++l) {
    const vc_vit_layer& LW = W.layer[l];
    VitWs::Layer& L = w.l[l];
    const uint32_t s0 = sb + 1 + 4 * l;
    // attention block: x2 = x + drop(to_out(attn(LN(x))))
    VC_TRY(layernorm_fwd(w.x[l], D, M, D, LW.ln1.w, LW.ln1.b, LN_EPS, nullptr, 0, L.h1.hi, L.h1.lo, D, L.m1, L.r1, st));
    {
      GemmDesc d;
      gemm_linear_fwd(d, L.h1, wsplit(LW.qkv, D), M, 3 * DI, D, P);
      d.out_hi = L.qkv.hi; d.out_lo = L.qkv.lo; d.ldo_split = 3 * DI;
      VC_TRY(gemm(d, st));
    }
    {
      AttnDesc a = vit_attn_desc(L, F, n, site_drop(p, c->training, c->seed, s0 + 0, c->seed_dev));
      VC_TRY(attention_fwd(a, L.o.hi, L.o.lo, DI, L.lse, st));
    }
    {
      GemmDesc d;
      gemm_linear_fwd(d, L.o, wsplit(LW.out, DI), M, D, DI, P);
      d.bias = LW.out.b; d.drop = site_drop(p, c->training, c->seed, s0 + 1, c->seed_dev);
      d.residual = w.x[l]; d.ld_res = D; d.out_f32 = L.x2; d.ldo = D;
      VC_TRY(gemm(d, st));
    }
    // MLP block: x3 = x2 + drop(fc2(drop(gelu(fc1(LN(x2))))))
    VC_TRY(layernorm_fwd(L.x2, D, M, D, LW.ln2.w, LW.ln2.b, LN_EPS, nullptr, 0, L.h2.hi, L.h2.lo, D, L.m2, L.r2, st));
    {
      GemmDesc d;
      gemm_linear_fwd(d, L.h2, wsplit(LW.fc1, D), M, VC_VIT_MLP, D, P);
      // L.pre1 receives gelu'(pre) * mask * scale of this dropout site (VC_ACT_GELU_DSTORE) -- the factor the fc2 dgrad epilogue
      // multiplies by -- instead of the pre-activation: erf and the mask exist here anyway, the backward then needs neither
      d.bias = LW.fc1.b; d.preact = L.pre1; d.ld_preact = VC_VIT_MLP; d.act = gelu_dstore() ? VC_ACT_GELU_DSTORE : VC_ACT_GELU;
      d.drop = site_drop(p, c->training, c->seed, s0 + 2, c->seed_dev);
      d.out_hi = L.ud.hi; d.out_lo = L.ud.lo; d.ldo_split = VC_VIT_MLP;
      VC_TRY(gemm(d, st));
    }
    {
      GemmDesc d;
      gemm_linear_fwd(d, L.ud, wsplit(LW.fc2, VC_VIT_MLP), M, D, VC_VIT_MLP, P);
      d.bias = LW.fc2.b; d.drop = site_drop(p, c->training, c->seed, s0 + 3, c->seed_dev);
      d.residual = L.x2; d.ld_res = D; d.out_f32 = w.x[l + 1]; d.ldo = D;
      VC_TRY(gemm(d, st));
    }
  }
  // final LayerNorm, evaluated on the CLS rows only (pool = 'cls')
  VC_TRY(layernorm_fwd(w.x[VC_VIT_DEPTH], (int64_t)n * D, F, D, W.norm.w, W.norm.b, LN_EPS, c->cls_out, D, nullptr, nullptr, 0,
                       w.fmean, w.frstd, st));
  return 0;
}

// Backward of encoder layers l_hi, l_hi - 1, ..., l_lo (0 <= l_lo <= l_hi < depth).  l_hi == depth - 1 also runs the final
// LayerNorm's backward (and needs dcls), l_lo == 0 also the token assembly / patch embedding.  Consecutive ranges called in
// descending order on the same scratch equal one full backward: the running gradient of the residual stream (s.dxa) and the
// masked split operand of the next fc2 (s.g) stay in the scratch between calls.  This lets the host hand the gradients of the
// upper layers to the data-parallel all-reduce while the lower layers are still in their backward.
int vit_backward_layers(const vc_vit_call* c, const float* dcls, void* scratch, size_t scratch_bytes, int l_hi, int l_lo, stream_t st) {
  VC_TRY(check_call(c));
  if (!scratch) return set_error("vit_backward: null argument");
  if (l_lo < 0 || l_hi >= VC_VIT_DEPTH || l_lo > l_hi) return set_error("vit_backward_layers: bad layer range");
  if (l_hi == VC_VIT_DEPTH - 1 && !dcls) return set_error("vit_backward: the top range needs dcls");
  const vc_vit_weights& W = *c->w;
  const int F = c->F, S = c->S, N = (S / VC_PATCH) * (S / VC_PATCH), n = N + 1;
  const int Mp = F * N, M = F * n, P = c->passes;
  Arena arena(c->ws, c->ws_bytes);
  VitWs w;
  vit_carve(arena, F, S, w);
  if (!arena.ok()) return set_error("vit_backward: workspace too small");
  Arena sa(scratch, scratch_bytes);
  VitScratch s;
  vit_scratch_carve(sa, F, S, s);
  if (!sa.ok()) return set_error("vit_backward: scratch too small");
  const float p = c->dropout_p;
  const uint32_t sb = c->site_base;
  const int ax2 = c->aux_streams > 0 ? c->aux_streams : 2, ax3 = ax2 + 1;  // this encoder's pair of auxiliary streams

  if (l_hi == VC_VIT_DEPTH - 1) {
    // final LN (CLS rows only): every other row of d x[6] is zero
    VC_TRY(zero_f32(s.dxa, (int64_t)M * D, st));
    VC_TRY(layernorm_bwd(dcls, D, w.x[VC_VIT_DEPTH], (int64_t)n * D, w.fmean, w.frstd, W.norm.w, F, D, nullptr, 0, s.dxa,
                         (int64_t)n * D, W.norm.dw, W.norm.db, st));
  }
  float* cur = s.dxa;  // d x[l + 1] at the top of every layer iteration (the two buffers swap twice per layer)
  float* other = s.dxb;
  for (int l = l_hi; l >= l_lo; --l) {
    const vc_vit_layer& LW = W.layer[l];
    VitWs::Layer& L = w.l[l];
    const uint32_t s0 = sb + 1 + 4 * l;
    // ---- MLP block.  s.g = split(d x[l+1] * mask(fc2 site)): for l < depth-1 it was produced, together with the fc2 bias
    // gradient, by the fused LayerNorm backward at the end of the previous iteration.
    if (l == VC_VIT_DEPTH - 1)
      VC_TRY(act_dropout_bwd(cur, D, M, D, VC_ACT_NONE, nullptr, 0, nullptr, 0, site_drop(p, c->training, c->seed, s0 + 3, c->seed_dev), nullptr,
                             0, s.g.hi, s.g.lo, D, LW.fc2.db, st));
    // The four weight-gradient GEMMs of a layer only feed the optimizer: they run on two auxiliary streams, next to the
    // dgrad / attention / LayerNorm chain, and fill the SMs the chain's partial last waves leave idle.  Each is joined
    // before the scratch operand it reads (s.g, s.dpre, s.dqkvS) is overwritten.
    stream_t side2, side3;
    VC_TRY(stream_fork(st, ax2, &side2));
    VC_TRY(linear_wgrad(s.g, L.ud, M, D, VC_VIT_MLP, LW.fc2.dw, P, side2));
    {
      // d pre1 = (g W2) * mask(MLP hidden site) * gelu'(pre1): activation backward + fc1 bias gradient fused in the epilogue
      GemmDesc d;
      gemm_linear_dgrad(d, s.g, wsplit(LW.fc2, VC_VIT_MLP), M, D, VC_VIT_MLP, P);
      d.act_backward = 1; d.act_aux = L.pre1; d.ld_act_aux = VC_VIT_MLP;
      if (gelu_dstore()) {
        d.act = VC_ACT_MUL_AUX;  // L.pre1 already holds mask * scale * gelu'(pre1)
      } else {
        d.act = VC_ACT_GELU;
        d.drop = site_drop(p, c->training, c->seed, s0 + 2, c->seed_dev);
      }
      d.out_hi = s.dpre.hi; d.out_lo = s.dpre.lo; d.ldo_split = VC_VIT_MLP; d.colsum = LW.fc1.db;
      VC_TRY(gemm(d, st));
    }
    VC_TRY(stream_fork(st, ax3, &side3));
    VC_TRY(linear_wgrad(s.dpre, L.h2, M, VC_VIT_MLP, D, LW.fc1.dw, P, side3));
    {
      GemmDesc d;
      gemm_linear_dgrad(d, s.dpre, wsplit(LW.fc1, D), M, VC_VIT_MLP, D, P);
      d.out_f32 = s.dh; d.ldo = D;
      VC_TRY(gemm(d, st));
    }
    VC_TRY(stream_join(st, ax2));  // fc2 wgrad has read s.g
    // d x2 = d x3 + LN2_bwd(dh); fused: s.g = split(d x2 * mask(to_out site)), to_out bias gradient
    VC_TRY(layernorm_bwd_fused(s.dh, D, L.x2, D, L.m2, L.r2, LW.ln2.w, M, D, cur, D, other, D, LW.ln2.dw, LW.ln2.db,
                               site_drop(p, c->training, c->seed, s0 + 1, c->seed_dev), s.g.hi, s.g.lo, D, LW.out.db, st));
    // ---- attention block (other = d x2)
    VC_TRY(stream_fork(st, ax2, &side2));
    VC_TRY(linear_wgrad(s.g, L.o, M, D, DI, LW.out.dw, P, side2));
    {
      GemmDesc d;
      gemm_linear_dgrad(d, s.g, wsplit(LW.out, DI), M, D, DI, P);
      d.out_hi = s.dOS.hi; d.out_lo = s.dOS.lo; d.ldo_split = DI;  // d(attention output), consumed as split-bf16
      VC_TRY(gemm(d, st));
    }
    {
      AttnDesc a = vit_attn_desc(L, F, n, site_drop(p, c->training, c->seed, s0 + 0, c->seed_dev));
      VC_TRY(attention_bwd_split(a, L.o.hi, L.o.lo, DI, L.lse, nullptr, s.dOS.hi, s.dOS.lo, DI, s.dqkv, s.dqkvS.hi, s.dqkvS.lo, s.dqkvS.hi + DI,
                                 s.dqkvS.lo + DI, s.dqkvS.hi + 2 * DI, s.dqkvS.lo + 2 * DI, 3 * DI, st));
    }
    VC_TRY(stream_fork(st, ax3, &side3));
    VC_TRY(linear_wgrad(s.dqkvS, L.h1, M, 3 * DI, D, LW.qkv.dw, P, side3));
    {
      GemmDesc d;
      gemm_linear_dgrad(d, s.dqkvS, wsplit(LW.qkv, D), M, 3 * DI, D, P);
      d.out_f32 = s.dh; d.ldo = D;
      VC_TRY(gemm(d, st));
    }
    VC_TRY(stream_join(st, ax2));  // to_out wgrad has read s.g
    if (l > 0) {
      // d x[l] = d x2 + LN1_bwd(dh); fused: s.g = split(d x[l] * mask(fc2 site of layer l-1)), fc2 bias gradient of layer l-1
      VC_TRY(layernorm_bwd_fused(s.dh, D, w.x[l], D, L.m1, L.r1, LW.ln1.w, M, D, other, D, cur, D, LW.ln1.dw, LW.ln1.db,
                                 site_drop(p, c->training, c->seed, sb + 1 + 4 * (l - 1) + 3, c->seed_dev), s.g.hi, s.g.lo, D,
                                 W.layer[l - 1].fc2.db, st));
    } else {
      VC_TRY(layernorm_bwd(s.dh, D, w.x[l], D, L.m1, L.r1, LW.ln1.w, M, D, other, D, cur, D, LW.ln1.dw, LW.ln1.db, st));
    }
    VC_TRY(stream_join(st, ax3));  // fc1 / to_qkv wgrads have read s.dpre / s.dqkvS (rewritten by the next layer's dgrad chain)
  }
  if (l_lo > 0) return 0;
  // ---- token assembly + patch embedding (cur = d x[0]); buffers reused: dh -> d e1, dud -> d e0, dO -> d pl
  float* de1 = s.dh;
  float* de0 = s.dud;
  float* dpl = s.dO;
  VC_TRY(vit_assemble_bwd(cur, F, N, D, site_drop(p, c->training, c->seed, sb + 0, c->seed_dev), de1, W.dcls, W.dpos, st));
  if (Mp > 0) {
    VC_TRY(layernorm_bwd(de1, D, w.e0, D, w.emean, w.erstd, W.pe_ln2.w, Mp, D, nullptr, 0, de0, D, W.pe_ln2.dw, W.pe_ln2.db, st));
    VC_TRY(act_dropout_bwd(de0, D, Mp, D, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, s.g.hi, s.g.lo, D, W.pe.db,
                           st));
    VC_TRY(linear_wgrad(s.g, w.pl, Mp, D, PD, W.pe.dw, P, st));
    {
      GemmDesc d;
      gemm_linear_dgrad(d, s.g, wsplit(W.pe, PD), Mp, D, PD, P);
      d.out_f32 = dpl; d.ldo = PD;
      VC_TRY(gemm(d, st));
    }
    VC_TRY(patch_layernorm_bwd_params(c->img, F, S, w.pmean, w.prstd, dpl, W.pe_ln1.dw, W.pe_ln1.db, st));
  }
  return 0;
}

}  // namespace vck

extern "C" {
size_t vc_vit_workspace_bytes(int F, int S) { return vck::vit_workspace_bytes(F, S); }
size_t vc_vit_scratch_bytes(int F, int S) { return vck::vit_scratch_bytes(F, S); }
int vc_vit_forward(const vc_vit_call* c, void* stream) { return vck::vit_forward(c, stream); }
int vc_vit_backward(const vc_vit_call* c, const float* dcls, void* scratch, size_t scratch_bytes, void* stream) {
  return vck::vit_backward_layers(c, dcls, scratch, scratch_bytes, VC_VIT_DEPTH - 1, 0, stream);
}
int vc_vit_backward_layers(const vc_vit_call* c, const float* dcls, void* scratch, size_t scratch_bytes, int l_hi, int l_lo, void* stream) {
  return vck::vit_backward_layers(c, dcls, scratch, scratch_bytes, l_hi, l_lo, stream);
}
}
