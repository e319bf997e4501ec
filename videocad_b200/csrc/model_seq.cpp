// Host orchestration of the action-sequence transformer: token construction, nn.TransformerDecoder (post-norm,
// ReLU, packed in_proj attention), the two action heads, and the full backward.
//
// Restates AutoRegressiveTransformer.forward after the image encoders
// (/root/reference/model/autoregressive_transformer.py:143-218) with torch's TransformerDecoderLayer defaults
// (norm_first=False, activation=relu, built at :54-62).  Math: SURVEY.md Appendix A.2 / A.3.
// Rows are batch-first (r = b*T + t); the reference's seq-first permutes are views and do not change the math.
//
// Branches (autoregressive_transformer.py:152-213):
//   past_actions            : tgt = tanh(embed_action(a)+E), causal self-attention;
//                             memory = tanh(image_projection([ui ; cad])) if past_states else tanh(cad)
//   past_states only        : tgt = ui = tanh(embed_state(vit(frames))+E), memory = tanh(cad), both masks banded
//   neither                 : tgt = memory = tanh(cad), both masks banded
// Dropout sites (site_base + 7*l + i): 0 self-attn probs, 1 dropout1, 2 cross-attn probs, 3 dropout2, 4 FFN hidden, 5 dropout3.
#include <math.h>
#include <vector>
#include <stdlib.h>
#include "model_common.h"

namespace vck {

namespace {

constexpr int VD = VC_VIT_DIM;
constexpr float LN_EPS = 1e-5f;

struct Dims {
  int B, T, R, H, Ff, L, nh, dh, NP, NC, nv;
  bool mem_has_ui, past_actions, past_states;
  // memory sources concatenated before image_projection: [ui][cad][multiview]; offsets in units of H
  int nsrc, cad_off, mv_off;
};

Dims dims_of(const vc_seq_call* c) {
  Dims d;
  d.B = c->B; d.T = c->T; d.R = c->B * c->T; d.H = c->H; d.Ff = c->Ff; d.L = c->w ? c->w->num_layers : 0;
  d.nh = c->nhead; d.dh = c->nhead > 0 ? c->H / c->nhead : 0; d.NP = c->num_param_out; d.NC = c->num_cmd;
  d.past_actions = c->past_actions != 0;
  d.past_states = c->past_states != 0;
  d.mem_has_ui = d.past_actions && d.past_states;
  d.nv = c->num_views > 0 ? c->num_views : 0;
  d.nsrc = (d.mem_has_ui ? 1 : 0) + 1 + (d.nv > 0 ? 1 : 0);
  d.cad_off = d.mem_has_ui ? 1 : 0;
  d.mv_off = d.cad_off + 1;
  return d;
}

struct SeqWs {
  Split scls, ccls;
  float* cad_tok;
  float* cat; Split catS;   // [R,3H]  sources of image_projection: [ui][cad token][multiview token], the latter two broadcast over T
  Split mvS; float* mv_tok; // multiview: split copy of the view embeddings [B, nv*512], token [B,H]
  float* ui; Split uiS;     // [R,H]   (past_states && !mem_has_ui)
  float* mem; Split memS;   // [R,H]
  float* act; Split actS;   // [R,H]   action tokens (past_actions)
  struct Layer {
    float* qkv; float* sa_lse; Split a; float* y1; float *m1, *r1; float* x1; Split x1S;
    float* q2; float* kv2; float* ca_lse; Split c; float* y2; float *m2, *r2; float* x2; Split x2S;
    Split f; float* y3; float *m3, *r3; float* x3; Split x3S;
  };
  std::vector<Layer> l;
};

void seq_carve(Arena& a, int B, int T, int H, int Ff, int L, int nh, int nv, SeqWs& w) {
  const size_t R = (size_t)B * T;
  w.scls = a.alloc_split(R, VD);
  w.ccls = a.alloc_split(B, VD);
  w.cad_tok = a.alloc<float>((size_t)B * H);
  w.cat = a.alloc<float>(R * 3 * H); w.catS = a.alloc_split(R, 3 * H);
  w.mvS = a.alloc_split(B, (int64_t)(nv > 0 ? nv : 1) * VD); w.mv_tok = a.alloc<float>((size_t)B * H);
  w.ui = a.alloc<float>(R * H); w.uiS = a.alloc_split(R, H);
  w.mem = a.alloc<float>(R * H); w.memS = a.alloc_split(R, H);
  w.act = a.alloc<float>(R * H); w.actS = a.alloc_split(R, H);
  w.l.resize(L);
  for (int i = 0; i < L; ++i) {
    SeqWs::Layer& Y = w.l[i];
    Y.qkv = a.alloc<float>(R * 3 * H); Y.sa_lse = a.alloc<float>((size_t)B * nh * T); Y.a = a.alloc_split(R, H);
    Y.y1 = a.alloc<float>(R * H); Y.m1 = a.alloc<float>(R); Y.r1 = a.alloc<float>(R);
    Y.x1 = a.alloc<float>(R * H); Y.x1S = a.alloc_split(R, H);
    Y.q2 = a.alloc<float>(R * H); Y.kv2 = a.alloc<float>(R * 2 * H); Y.ca_lse = a.alloc<float>((size_t)B * nh * T);
    Y.c = a.alloc_split(R, H);
    Y.y2 = a.alloc<float>(R * H); Y.m2 = a.alloc<float>(R); Y.r2 = a.alloc<float>(R);
    Y.x2 = a.alloc<float>(R * H); Y.x2S = a.alloc_split(R, H);
    Y.f = a.alloc_split(R, Ff);
    Y.y3 = a.alloc<float>(R * H); Y.m3 = a.alloc<float>(R); Y.r3 = a.alloc<float>(R);
    Y.x3 = a.alloc<float>(R * H); Y.x3S = a.alloc_split(R, H);
  }
}

struct SeqScratch {
  float *A, *Bf, *Y, *dF, *dAtt, *dqkv, *dmem, *dcat, *dcadtok, *dmvtok;
  Split gB2;
  Split gH, dpreF, dqkvS, dparS, gB;
  Split gH_ca, gH_sa, dqkvS_sa;  // separate buffers per use: the weight-gradient GEMMs read them on an auxiliary stream
};

void seq_scratch_carve(Arena& a, int B, int T, int H, int Ff, int NP, int nv, SeqScratch& s) {
  (void)nv;
  const size_t R = (size_t)B * T;
  s.A = a.alloc<float>(R * H); s.Bf = a.alloc<float>(R * H); s.Y = a.alloc<float>(R * H);
  s.dF = a.alloc<float>(R * Ff); s.dAtt = a.alloc<float>(R * H);
  s.dqkv = a.alloc<float>(R * 3 * H); s.dmem = a.alloc<float>(R * H);
  s.dcat = a.alloc<float>(R * 3 * H); s.dcadtok = a.alloc<float>((size_t)B * H); s.dmvtok = a.alloc<float>((size_t)B * H);
  s.gB2 = a.alloc_split(B, H);
  s.gH = a.alloc_split(R, H); s.dpreF = a.alloc_split(R, Ff); s.dqkvS = a.alloc_split(R, 3 * H);
  s.dparS = a.alloc_split(R, NP); s.gB = a.alloc_split(B, H);
  s.gH_ca = a.alloc_split(R, H); s.gH_sa = a.alloc_split(R, H); s.dqkvS_sa = a.alloc_split(R, 3 * H);
}

int check_call(const vc_seq_call* c) {
  if (!c || !c->w || !c->cad_cls || !c->ws || !c->cmds || !c->params) return set_error("seq: null argument");
  if (c->B <= 0 || c->T <= 0) return set_error("seq: B and T must be positive");
  if (c->H % 128 != 0 || c->H > 1024) return set_error("seq: hidden_size must be a multiple of 128 and <= 1024");
  if (c->Ff % 8 != 0) return set_error("seq: dim_feedforward must be a multiple of 8");
  if (c->nhead <= 0 || c->H % c->nhead != 0) return set_error("seq: hidden_size must be divisible by nhead");
  if ((c->H / c->nhead) % 4 != 0 || c->H / c->nhead > 256) return set_error("seq: head dim must be a multiple of 4 and <= 256");
  if (c->window < 1) return set_error("seq: window_size must be > 0");
  if (c->num_param_out % 8 != 0) return set_error("seq: parameter head width must be a multiple of 8");
  if (c->num_cmd > 8) return set_error("seq: more than 8 command classes unsupported");
  if (c->past_states && !c->state_cls) return set_error("seq: past_states needs state_cls");
  if (c->past_actions && !c->actions) return set_error("seq: past_actions needs actions");
  if (c->passes != 1 && c->passes != 3) return set_error("seq: passes must be 1 or 3");
  if (c->w->num_layers < 1 || !c->w->layers) return set_error("seq: need at least one decoder layer");
  if (c->num_views > 0 && (!c->mv_cls || !c->w->embed_multiview.w)) return set_error("seq: num_views > 0 needs mv_cls and embed_multiview");
  if (c->num_views > 0 && c->past_states && !c->past_actions)
    return set_error("seq: multiview with past_states only is shape-inconsistent in the reference as well (image_projection width)");
  return 0;
}

AttnDesc self_attn_desc(const vc_seq_call* c, const Dims& d, const SeqWs::Layer& Y, Drop drop) {
  AttnDesc a = {};
  a.q = Y.qkv; a.k = Y.qkv + d.H; a.v = Y.qkv + 2 * d.H;
  a.ldq = a.ldk = a.ldv = 3 * d.H;
  a.B = d.B; a.Tq = d.T; a.Tk = d.T; a.nh = d.nh; a.d = d.dh;
  a.mask = d.past_actions ? VC_MASK_CAUSAL : VC_MASK_WINDOW;
  a.window = c->window;
  a.scale = 1.0f / sqrtf((float)d.dh);
  a.drop = drop;
  return a;
}
AttnDesc cross_attn_desc(const vc_seq_call* c, const Dims& d, const SeqWs::Layer& Y, Drop drop) {
  AttnDesc a = {};
  a.q = Y.q2; a.k = Y.kv2; a.v = Y.kv2 + d.H;
  a.ldq = d.H; a.ldk = a.ldv = 2 * d.H;
  a.B = d.B; a.Tq = d.T; a.Tk = d.T; a.nh = d.nh; a.d = d.dh;
  a.mask = VC_MASK_WINDOW;
  a.window = c->window;
  a.scale = 1.0f / sqrtf((float)d.dh);
  a.drop = drop;
  return a;
}

static inline Split cols(const Split& s, int64_t col0) { return mk_split(s.hi + col0, s.lo + col0, s.ld); }

}  // namespace

size_t seq_workspace_bytes(int B, int T, int H, int Ff, int L, int nh, int nv) {
  Arena a(nullptr, 0);
  SeqWs w;
  seq_carve(a, B, T, H, Ff, L, nh, nv, w);
  return a.used();
}
size_t seq_scratch_bytes(int B, int T, int H, int Ff, int NP, int nv) {
  Arena a(nullptr, 0);
  SeqScratch s;
  seq_scratch_carve(a, B, T, H, Ff, NP, nv, s);
  return a.used();
}

int seq_forward(const vc_seq_call* c, stream_t st) {
  VC_TRY(check_call(c));
  const vc_seq_weights& W = *c->w;
  const Dims d = dims_of(c);
  const int B = d.B, T = d.T, R = d.R, H = d.H, Ff = d.Ff, P = c->passes;
  Arena arena(c->ws, c->ws_bytes);
  SeqWs w;
  seq_carve(arena, B, T, H, Ff, d.L, d.nh, d.nv, w);
  if (!arena.ok()) return set_error("seq_forward: workspace too small");
  const float p = c->dropout_p;
  const float* E = W.timestep_emb;  // rows 0..T-1 are the positions arange(T)

  // ---- frame tokens: ui = tanh(embed_state(cls) + E[t])
  float* ui = nullptr; Split uiS = mk_split(nullptr, nullptr, 0); int64_t ld_ui = 0;
  if (d.past_states) {
    VC_TRY(split_f32(c->state_cls, VD, R, VD, w.scls.hi, w.scls.lo, VD, st));
    if (d.mem_has_ui) { ld_ui = (int64_t)d.nsrc * H; ui = w.cat; uiS = mk_split(w.catS.hi, w.catS.lo, ld_ui); } else { ui = w.ui; uiS = w.uiS; ld_ui = H; }
    GemmDesc g;
    gemm_linear_fwd(g, w.scls, wsplit(W.embed_state, VD), R, H, VD, P);
    g.bias = W.embed_state.b;
    if (E) { g.rowadd = E; g.ld_rowadd = H; g.rowadd_div = 1; g.rowadd_mod = T; }
    g.act = VC_ACT_TANH;
    g.out_f32 = ui; g.ldo = ld_ui; g.out_hi = uiS.hi; g.out_lo = uiS.lo; g.ldo_split = uiS.ld;
    VC_TRY(gemm(g, st));
  }
  // ---- CAD token (static over T)
  VC_TRY(split_f32(c->cad_cls, VD, B, VD, w.ccls.hi, w.ccls.lo, VD, st));
  {
    GemmDesc g;
    gemm_linear_fwd(g, w.ccls, wsplit(W.embed_image, VD), B, H, VD, P);
    g.bias = W.embed_image.b;
    g.act = d.nsrc > 1 ? VC_ACT_NONE : VC_ACT_TANH;
    g.out_f32 = w.cad_tok; g.ldo = H;
    VC_TRY(gemm(g, st));
  }
  const int64_t ldc = (int64_t)d.nsrc * H;
  if (d.nv > 0) {
    // multiview token (trajectory_model.py:77-87, autoregressive_transformer.py:167-170): Linear over the concatenated view
    // embeddings, identical for every time step
    const int Kmv = d.nv * VD;
    VC_TRY(split_f32(c->mv_cls, Kmv, B, Kmv, w.mvS.hi, w.mvS.lo, Kmv, st));
    GemmDesc g;
    gemm_linear_fwd(g, w.mvS, wsplit(W.embed_multiview, Kmv), B, H, Kmv, P);
    g.bias = W.embed_multiview.b; g.out_f32 = w.mv_tok; g.ldo = H;
    VC_TRY(gemm(g, st));
    VC_TRY(broadcast_rows(w.mv_tok, H, R, H, T, w.cat + (int64_t)d.mv_off * H, ldc, w.catS.hi + (int64_t)d.mv_off * H,
                          w.catS.lo + (int64_t)d.mv_off * H, ldc, st));
  }
  if (d.nsrc > 1) {
    VC_TRY(broadcast_rows(w.cad_tok, H, R, H, T, w.cat + (int64_t)d.cad_off * H, ldc, w.catS.hi + (int64_t)d.cad_off * H,
                          w.catS.lo + (int64_t)d.cad_off * H, ldc, st));
    GemmDesc g;
    gemm_linear_fwd(g, mk_split(w.catS.hi, w.catS.lo, ldc), wsplit(W.image_proj, ldc), R, H, (int)ldc, P);
    g.bias = W.image_proj.b; g.act = VC_ACT_TANH;
    g.out_f32 = w.mem; g.ldo = H; g.out_hi = w.memS.hi; g.out_lo = w.memS.lo; g.ldo_split = H;
    VC_TRY(gemm(g, st));
  } else {
    VC_TRY(broadcast_rows(w.cad_tok, H, R, H, T, w.mem, H, w.memS.hi, w.memS.lo, H, st));
  }
  // ---- target tokens
  const float* x_in; Split x_inS; int64_t ld_x;
  if (d.past_actions) {
    VC_TRY(embed_action_fwd(c->actions, R, c->act_dim, H, W.embed_action_w, W.embed_action_b, E, T, w.act, w.actS.hi, w.actS.lo, st));
    x_in = w.act; x_inS = w.actS; ld_x = H;
  } else if (d.past_states) {
    x_in = ui; x_inS = uiS; ld_x = ld_ui;
  } else {
    x_in = w.mem; x_inS = w.memS; ld_x = H;
  }

  // The cross-attention K/V projections of ALL layers depend only on the memory tokens: issue them on an auxiliary stream,
  // next to the first self-attention block, instead of inside the per-layer chain.
  {
    stream_t side;
    VC_TRY(stream_fork(st, 0, &side));
    for (int l = 0; l < d.L; ++l) {
      const vc_dec_layer& LW = W.layers[l];
      GemmDesc g;
      gemm_linear_fwd(g, w.memS, wsplit(LW.ca_in, H, H), R, 2 * H, H, P);
      g.bias = LW.ca_in.b ? LW.ca_in.b + H : nullptr; g.out_f32 = w.l[l].kv2; g.ldo = 2 * H;
      VC_TRY(gemm(g, side));
    }
  }
  for (int l = 0; l < d.L; ++l) {
    const vc_dec_layer& LW = W.layers[l];
    SeqWs::Layer& Y = w.l[l];
    const uint32_t s0 = c->site_base + 7 * l;
    // self-attention block: x1 = LN1(x + drop(out_proj(SA(x))))
    {
      GemmDesc g;
      gemm_linear_fwd(g, x_inS, wsplit(LW.sa_in, H), R, 3 * H, H, P);
      g.bias = LW.sa_in.b; g.out_f32 = Y.qkv; g.ldo = 3 * H;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(attention_fwd(self_attn_desc(c, d, Y, site_drop(p, c->training, c->seed, s0 + 0, c->seed_dev)), Y.a.hi, Y.a.lo, H, Y.sa_lse, st));
    {
      GemmDesc g;
      gemm_linear_fwd(g, Y.a, wsplit(LW.sa_out, H), R, H, H, P);
      g.bias = LW.sa_out.b; g.drop = site_drop(p, c->training, c->seed, s0 + 1, c->seed_dev);
      g.residual = x_in; g.ld_res = ld_x; g.out_f32 = Y.y1; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(layernorm_fwd(Y.y1, H, R, H, LW.n1.w, LW.n1.b, LN_EPS, Y.x1, H, Y.x1S.hi, Y.x1S.lo, H, Y.m1, Y.r1, st));
    // cross-attention block: x2 = LN2(x1 + drop(out_proj(CA(x1, mem))))
    {
      GemmDesc g;
      gemm_linear_fwd(g, Y.x1S, wsplit(LW.ca_in, H, 0), R, H, H, P);
      g.bias = LW.ca_in.b; g.out_f32 = Y.q2; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    if (l == 0) VC_TRY(stream_join(st, 0));  // K/V projections of all layers are complete from here on
    VC_TRY(attention_fwd(cross_attn_desc(c, d, Y, site_drop(p, c->training, c->seed, s0 + 2, c->seed_dev)), Y.c.hi, Y.c.lo, H, Y.ca_lse, st));
    {
      GemmDesc g;
      gemm_linear_fwd(g, Y.c, wsplit(LW.ca_out, H), R, H, H, P);
      g.bias = LW.ca_out.b; g.drop = site_drop(p, c->training, c->seed, s0 + 3, c->seed_dev);
      g.residual = Y.x1; g.ld_res = H; g.out_f32 = Y.y2; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(layernorm_fwd(Y.y2, H, R, H, LW.n2.w, LW.n2.b, LN_EPS, Y.x2, H, Y.x2S.hi, Y.x2S.lo, H, Y.m2, Y.r2, st));
    // feed-forward block: x3 = LN3(x2 + drop(linear2(drop(relu(linear1(x2))))))
    {
      GemmDesc g;
      gemm_linear_fwd(g, Y.x2S, wsplit(LW.lin1, H), R, Ff, H, P);
      g.bias = LW.lin1.b; g.act = VC_ACT_RELU; g.drop = site_drop(p, c->training, c->seed, s0 + 4, c->seed_dev);
      g.out_hi = Y.f.hi; g.out_lo = Y.f.lo; g.ldo_split = Ff;
      VC_TRY(gemm(g, st));
    }
    {
      GemmDesc g;
      gemm_linear_fwd(g, Y.f, wsplit(LW.lin2, Ff), R, H, Ff, P);
      g.bias = LW.lin2.b; g.drop = site_drop(p, c->training, c->seed, s0 + 5, c->seed_dev);
      g.residual = Y.x2; g.ld_res = H; g.out_f32 = Y.y3; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(layernorm_fwd(Y.y3, H, R, H, LW.n3.w, LW.n3.b, LN_EPS, Y.x3, H, Y.x3S.hi, Y.x3S.lo, H, Y.m3, Y.r3, st));
    x_in = Y.x3; x_inS = Y.x3S; ld_x = H;
  }
  // ---- heads
  {
    stream_t side;
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(head_small_fwd(x_in, R, H, W.head_cmd_w, W.head_cmd_b, d.NC, c->cmds, side));
  }
  {
    GemmDesc g;
    gemm_linear_fwd(g, x_inS, wsplit(W.head_params, H), R, d.NP, H, P);
    g.bias = W.head_params.b; g.out_f32 = c->params; g.ldo = d.NP;
    VC_TRY(gemm(g, st));
  }
  VC_TRY(stream_join(st, 0));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Incremental decoding (autoregressive rollout with action feedback, autoregressive_transformer.py:222-275).
//
// The reference re-runs the whole forward on the growing prefix at every step.  The forward is causal, so position t depends on
// positions <= t only, and everything position t needs from the past is (i) the self-attention keys/values of every layer at
// positions < t and (ii) the cross-attention keys/values of the memory tokens, which do not depend on the actions at all.
// seq_forward() over the full length (any actions) leaves exactly these in its workspace: Y.kv2 = cross-attention K|V of all
// positions, Y.qkv = self-attention q|k|v rows.  seq_decode_step(t) then pushes ONE token per sequence through the decoder:
// its q|k|v row is written straight into row (b, t) of Y.qkv -- the key/value cache -- and the two attention calls read the
// first t + 1 rows of the cache / the window of Y.kv2 through the batch strides of vc_attn_desc.  Steps must be issued in
// order t = 0, 1, ... after one seq_forward() on the same workspace.  Eval mode only (no dropout), past_actions required.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct StepWs {
  float* act; Split actS;
  Split a, c, f;
  float *y, *x1, *x2, *x3, *q2, *ff;
  Split x1S, x2S, x3S;
  float *lse, *mean, *rstd;
};
void step_carve(Arena& a, int B, int H, int Ff, int nh, StepWs& s) {
  s.act = a.alloc<float>((size_t)B * H); s.actS = a.alloc_split(B, H);
  s.a = a.alloc_split(B, H); s.c = a.alloc_split(B, H); s.f = a.alloc_split(B, Ff);
  s.y = a.alloc<float>((size_t)B * H);
  s.x1 = a.alloc<float>((size_t)B * H); s.x2 = a.alloc<float>((size_t)B * H); s.x3 = a.alloc<float>((size_t)B * H);
  s.q2 = a.alloc<float>((size_t)B * H); s.ff = a.alloc<float>((size_t)B * Ff);
  s.x1S = a.alloc_split(B, H); s.x2S = a.alloc_split(B, H); s.x3S = a.alloc_split(B, H);
  s.lse = a.alloc<float>((size_t)B * nh); s.mean = a.alloc<float>(B); s.rstd = a.alloc<float>(B);
}
}  // namespace

size_t seq_decode_scratch_bytes(int B, int H, int Ff, int nh) {
  Arena a(nullptr, 0);
  StepWs s;
  step_carve(a, B, H, Ff, nh, s);
  return a.used();
}

int seq_decode_step(const vc_seq_call* c, int t, const float* actions_t, void* scratch, size_t scratch_bytes, float* cmds_t,
                    float* params_t, stream_t st) {
  VC_TRY(check_call(c));
  if (!actions_t || !scratch || !cmds_t || !params_t) return set_error("seq_decode_step: null argument");
  if (!c->past_actions) return set_error("seq_decode_step: enable_past_actions is required (the target tokens are the fed-back actions)");
  if (c->training) return set_error("seq_decode_step: eval mode only");
  if (t < 0 || t >= c->T) return set_error("seq_decode_step: step out of range");
  const vc_seq_weights& W = *c->w;
  const Dims d = dims_of(c);
  const int B = d.B, T = d.T, H = d.H, Ff = d.Ff, P = c->passes;
  Arena arena(c->ws, c->ws_bytes);
  SeqWs w;
  seq_carve(arena, B, T, H, Ff, d.L, d.nh, d.nv, w);
  if (!arena.ok()) return set_error("seq_decode_step: workspace too small");
  Arena sa(scratch, scratch_bytes);
  StepWs s;
  step_carve(sa, B, H, Ff, d.nh, s);
  if (!sa.ok()) return set_error("seq_decode_step: scratch too small");

  // One Linear of the step: x (fp32 and/or split) -> out.  Up to 16 sequences: the exact-fp32 row kernel (the N x K weights are
  // spread over all SMs); larger batches: the tensor-core GEMM on the split operands, as in the full forward.
  const bool rows_path = B <= 16 && H % 128 == 0 && Ff % 128 == 0;
  auto linear = [&](const float* xf, const Split& xs, const vc_linear& Lw, int64_t row0, int N, int K, int act, const float* res,
                    float* out, int64_t ldo, const Split* outS) -> int {
    const float* bias = Lw.b ? Lw.b + row0 : nullptr;
    if (rows_path) {
      return linear_rows_fwd(xf, xf ? nullptr : xs.hi, xf ? nullptr : xs.lo, xf ? (int64_t)K : xs.ld, B, Lw.w + row0 * K, bias, N, K, act, res, H,
                             out, ldo, outS ? outS->hi : nullptr, outS ? outS->lo : nullptr, outS ? outS->ld : 0, st);
    }
    GemmDesc g;
    gemm_linear_fwd(g, xs, wsplit(Lw, K, row0), B, N, K, P);
    g.bias = bias; g.act = act; g.residual = res; g.ld_res = H; g.out_f32 = out; g.ldo = ldo;
    if (outS) { g.out_hi = outS->hi; g.out_lo = outS->lo; g.ldo_split = outS->ld; }
    return gemm(g, st);
  };

  // token of step t: tanh(W_a a_t + b + E[t])   (embed_action_fwd indexes E by row % T: one row per sequence -> T = 1, E + t*H)
  const float* E = W.timestep_emb ? W.timestep_emb + (int64_t)t * H : nullptr;
  VC_TRY(embed_action_fwd(actions_t, B, c->act_dim, H, W.embed_action_w, W.embed_action_b, E, 1, s.act, s.actS.hi, s.actS.lo, st));
  const float* x_in = s.act; Split x_inS = s.actS;
  const float scale = 1.0f / sqrtf((float)d.dh);
  for (int l = 0; l < d.L; ++l) {
    const vc_dec_layer& LW = W.layers[l];
    SeqWs::Layer& Y = w.l[l];
    // q | k | v of the new token -> row (b, t) of the cache
    VC_TRY(linear(x_in, x_inS, LW.sa_in, 0, 3 * H, H, VC_ACT_NONE, nullptr, Y.qkv + (int64_t)t * 3 * H, (int64_t)T * 3 * H, nullptr));
    {  // self-attention of the new token over cache rows [0, t]: every cached key is in its causal past
      AttnDesc a = {};
      a.q = Y.qkv + (int64_t)t * 3 * H; a.ldq = 3 * H; a.bsq = (int64_t)T * 3 * H;
      a.k = Y.qkv + H; a.v = Y.qkv + 2 * H; a.ldk = a.ldv = 3 * H; a.bsk = a.bsv = (int64_t)T * 3 * H;
      a.B = B; a.Tq = 1; a.Tk = t + 1; a.nh = d.nh; a.d = d.dh;
      a.mask = VC_MASK_NONE; a.window = 1; a.scale = scale; a.drop = no_drop();
      VC_TRY(attention_fwd(a, s.a.hi, s.a.lo, H, s.lse, st));
    }
    VC_TRY(linear(nullptr, s.a, LW.sa_out, 0, H, H, VC_ACT_NONE, x_in, s.y, H, nullptr));
    VC_TRY(layernorm_fwd(s.y, H, B, H, LW.n1.w, LW.n1.b, LN_EPS, s.x1, H, s.x1S.hi, s.x1S.lo, H, s.mean, s.rstd, st));
    VC_TRY(linear(s.x1, s.x1S, LW.ca_in, 0, H, H, VC_ACT_NONE, nullptr, s.q2, H, nullptr));
    {  // cross-attention over the memory window (t - W, t]
      const int jlo = t - c->window + 1 > 0 ? t - c->window + 1 : 0;
      AttnDesc a = {};
      a.q = s.q2; a.ldq = H; a.bsq = H;
      a.k = Y.kv2 + (int64_t)jlo * 2 * H; a.v = a.k + H; a.ldk = a.ldv = 2 * H; a.bsk = a.bsv = (int64_t)T * 2 * H;
      a.B = B; a.Tq = 1; a.Tk = t - jlo + 1; a.nh = d.nh; a.d = d.dh;
      a.mask = VC_MASK_NONE; a.window = 1; a.scale = scale; a.drop = no_drop();
      VC_TRY(attention_fwd(a, s.c.hi, s.c.lo, H, s.lse, st));
    }
    VC_TRY(linear(nullptr, s.c, LW.ca_out, 0, H, H, VC_ACT_NONE, s.x1, s.y, H, nullptr));
    VC_TRY(layernorm_fwd(s.y, H, B, H, LW.n2.w, LW.n2.b, LN_EPS, s.x2, H, s.x2S.hi, s.x2S.lo, H, s.mean, s.rstd, st));
    VC_TRY(linear(s.x2, s.x2S, LW.lin1, 0, Ff, H, VC_ACT_RELU, nullptr, rows_path ? s.ff : nullptr, Ff, rows_path ? nullptr : &s.f));
    VC_TRY(linear(rows_path ? s.ff : nullptr, s.f, LW.lin2, 0, H, Ff, VC_ACT_NONE, s.x2, s.y, H, nullptr));
    // the layer output ping-pongs between the (x3, x3S) and (act, actS) buffers: it must not overwrite this layer's own input,
    // which the residual connections above read
    float* xo = (x_in == s.act) ? s.x3 : s.act;
    const Split xoS = (x_in == s.act) ? s.x3S : s.actS;
    VC_TRY(layernorm_fwd(s.y, H, B, H, LW.n3.w, LW.n3.b, LN_EPS, xo, H, xoS.hi, xoS.lo, H, s.mean, s.rstd, st));
    x_in = xo; x_inS = xoS;
  }
  VC_TRY(head_small_fwd(x_in, B, H, W.head_cmd_w, W.head_cmd_b, d.NC, cmds_t, st));
  VC_TRY(linear(x_in, x_inS, W.head_params, 0, d.NP, H, VC_ACT_NONE, nullptr, params_t, d.NP, nullptr));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same step for a handful of sequences (B <= 16), bound by reading every weight once: 8 launches per layer (decode.cu) --
//   K1 q|k|v of the new token (input rows built on load: token embedding for layer 0, LayerNorm3 of the previous layer
//      otherwise) -> cache row (b, t)            K2 self-attention over cache rows [0, t], keys split over 8 CTAs per head
//   K3 out_proj(.) + x                            K4 cross-attention query = W_q LayerNorm1(.)
//   K5 cross-attention over the memory window     K6 out_proj(.) + x1
//   K7 relu(linear1(LayerNorm2(.)))               K8 linear2(.) + x2
// then the parameter head (LayerNorm3 on load) writes row (b, t) of params_all and dec_select_kernel finishes the step on the
// device: command head, both argmaxes, action mask, normalisation -> actions_io (the next step's input), *t_dev += 1.
// Nothing depends on the host: the launch sequence is identical for every position (the position is read from *t_dev), so one
// captured CUDA graph replays for all T steps.  Same contract as seq_decode_step otherwise (one seq_forward first; eval mode).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int DEC_SPLITS = 8;  // self-attention keys of one (sequence, head) are split over this many CTAs
struct DevStepWs {
  float *xn0, *xn1, *xn2, *y1, *y2, *y3, *q2, *f, *sa, *ca;
  float *part_o, *part_ml;
  unsigned int *counters, *done;
};
void dev_step_carve(Arena& a, int B, int H, int Ff, int nh, DevStepWs& s) {
  const size_t BH = (size_t)B * H;
  s.counters = a.alloc<unsigned int>((size_t)B * nh);  // zero before the first step, left zero by every kernel
  s.done = a.alloc<unsigned int>(64);                  // likewise
  s.xn0 = a.alloc<float>(BH); s.xn1 = a.alloc<float>(BH); s.xn2 = a.alloc<float>(BH);
  s.y1 = a.alloc<float>(BH); s.y2 = a.alloc<float>(BH); s.y3 = a.alloc<float>(BH);
  s.q2 = a.alloc<float>(BH); s.f = a.alloc<float>((size_t)B * Ff);
  s.sa = a.alloc<float>(BH); s.ca = a.alloc<float>(BH);
  s.part_o = a.alloc<float>(BH * DEC_SPLITS); s.part_ml = a.alloc<float>((size_t)B * nh * DEC_SPLITS * 2);
}
}  // namespace

int seq_decode_dev_supported(const vc_seq_call* c) {
  if (check_call(c)) return 0;
  const int dh = c->H / c->nhead;
  return c->past_actions && !c->training && c->B <= 16 && c->H % 128 == 0 && c->H <= 1024 && c->Ff % 128 == 0 && c->Ff <= 1024 &&
         (dh == 64 || dh == 128 || dh == 256) && c->num_cmd <= 5 && c->num_param_out % 1000 == 0 && c->num_param_out / 1000 <= 6 &&
         c->act_dim == 1 + c->num_param_out / 1000;
}

size_t seq_decode_dev_scratch_bytes(int B, int H, int Ff, int nh) {
  Arena a(nullptr, 0);
  DevStepWs s;
  dev_step_carve(a, B, H, Ff, nh, s);
  return a.used();
}

int seq_decode_step_dev(const vc_seq_call* c, int* t_dev, float* actions_io, void* scratch, size_t scratch_bytes, float* cmds_all,
                        float* params_all, stream_t st) {
  VC_TRY(check_call(c));
  if (!t_dev || !actions_io || !scratch || !cmds_all || !params_all) return set_error("seq_decode_step_dev: null argument");
  if (!seq_decode_dev_supported(c)) return set_error("seq_decode_step_dev: unsupported configuration (use vc_seq_decode_step)");
  const vc_seq_weights& W = *c->w;
  const Dims d = dims_of(c);
  const int B = d.B, T = d.T, H = d.H, Ff = d.Ff;
  Arena arena(c->ws, c->ws_bytes);
  SeqWs w;
  seq_carve(arena, B, T, H, Ff, d.L, d.nh, d.nv, w);
  if (!arena.ok()) return set_error("seq_decode_step_dev: workspace too small");
  Arena sa(scratch, scratch_bytes);
  DevStepWs s;
  dev_step_carve(sa, B, H, Ff, d.nh, s);
  if (!sa.ok()) return set_error("seq_decode_step_dev: scratch too small");
  const float scale = 1.0f / sqrtf((float)d.dh);

  auto gemv = [&](DecGemv& g, const vc_linear& Lw, int64_t row0, int N, int K, float* out, int64_t row_stride, int64_t t_stride) -> int {
    g.M = B; g.N = N; g.K = K;
    g.W = Lw.w + row0 * K; g.bias = Lw.b ? Lw.b + row0 : nullptr;
    g.out = out; g.out_row_stride = row_stride; g.out_t_stride = t_stride; g.t_ptr = t_dev;
    return dec_gemv(g, st);
  };
  for (int l = 0; l < d.L; ++l) {
    const vc_dec_layer& LW = W.layers[l];
    SeqWs::Layer& Y = w.l[l];
    {  // K1
      DecGemv g = {};
      if (l == 0) {
        g.in_mode = VC_DEC_IN_EMBED; g.actions = actions_io; g.act_dim = c->act_dim;
        g.emb_W = W.embed_action_w; g.emb_b = W.embed_action_b; g.emb_E = W.timestep_emb;
      } else {
        g.in_mode = VC_DEC_IN_LN; g.x = s.y3; g.ldx = H; g.gamma = W.layers[l - 1].n3.w; g.beta = W.layers[l - 1].n3.b;
      }
      g.x_out = s.xn0;
      VC_TRY(gemv(g, LW.sa_in, 0, 3 * H, H, Y.qkv, (int64_t)T * 3 * H, 3 * H));
    }
    {  // K2
      DecAttn a = {};
      a.q = Y.qkv; a.q_bstride = (int64_t)T * 3 * H; a.q_tstride = 3 * H;
      a.k = Y.qkv + H; a.v = Y.qkv + 2 * H; a.kv_bstride = (int64_t)T * 3 * H; a.kv_rstride = 3 * H;
      a.nh = d.nh; a.dh = d.dh; a.nsplit = DEC_SPLITS; a.window = 0; a.scale = scale; a.t_ptr = t_dev;
      a.out = s.sa; a.ld_out = H; a.part_o = s.part_o; a.part_ml = s.part_ml; a.counters = s.counters;
      VC_TRY(dec_attn(a, B, st));
    }
    {  // K3
      DecGemv g = {};
      g.in_mode = VC_DEC_IN_PLAIN; g.x = s.sa; g.ldx = H;
      g.residual = s.xn0; g.ld_res = H;
      VC_TRY(gemv(g, LW.sa_out, 0, H, H, s.y1, H, 0));
    }
    {  // K4
      DecGemv g = {};
      g.in_mode = VC_DEC_IN_LN; g.x = s.y1; g.ldx = H; g.gamma = LW.n1.w; g.beta = LW.n1.b; g.x_out = s.xn1;
      VC_TRY(gemv(g, LW.ca_in, 0, H, H, s.q2, H, 0));
    }
    {  // K5
      DecAttn a = {};
      a.q = s.q2; a.q_bstride = H; a.q_tstride = 0;
      a.k = Y.kv2; a.v = Y.kv2 + H; a.kv_bstride = (int64_t)T * 2 * H; a.kv_rstride = 2 * H;
      a.nh = d.nh; a.dh = d.dh; a.nsplit = 1; a.window = c->window; a.scale = scale; a.t_ptr = t_dev;
      a.out = s.ca; a.ld_out = H;
      VC_TRY(dec_attn(a, B, st));
    }
    {  // K6
      DecGemv g = {};
      g.in_mode = VC_DEC_IN_PLAIN; g.x = s.ca; g.ldx = H;
      g.residual = s.xn1; g.ld_res = H;
      VC_TRY(gemv(g, LW.ca_out, 0, H, H, s.y2, H, 0));
    }
    {  // K7
      DecGemv g = {};
      g.in_mode = VC_DEC_IN_LN; g.x = s.y2; g.ldx = H; g.gamma = LW.n2.w; g.beta = LW.n2.b; g.x_out = s.xn2;
      g.act = VC_ACT_RELU;
      VC_TRY(gemv(g, LW.lin1, 0, Ff, H, s.f, Ff, 0));
    }
    {  // K8
      DecGemv g = {};
      g.in_mode = VC_DEC_IN_PLAIN; g.x = s.f; g.ldx = Ff;
      g.residual = s.xn2; g.ld_res = H;
      VC_TRY(gemv(g, LW.lin2, 0, H, Ff, s.y3, H, 0));
    }
  }
  const vc_dec_layer& last = W.layers[d.L - 1];
  {
    DecGemv g = {};
    g.in_mode = VC_DEC_IN_LN; g.x = s.y3; g.ldx = H; g.gamma = last.n3.w; g.beta = last.n3.b;
    VC_TRY(gemv(g, W.head_params, 0, d.NP, H, params_all, (int64_t)T * d.NP, d.NP));
  }
  {
    DecSelect a = {};
    a.y = s.y3; a.gamma = last.n3.w; a.beta = last.n3.b;
    a.Wc = W.head_cmd_w; a.bc = W.head_cmd_b; a.NC = d.NC;
    a.H = H; a.B = B; a.NPAR = d.NP / 1000; a.NV = 1000; a.T = T;
    a.cmds_all = cmds_all; a.params_all = params_all; a.action_next = actions_io;
    a.t_ptr = t_dev; a.done_ctr = s.done;
    VC_TRY(dec_select(a, st));
  }
  return 0;
}

int seq_backward(const vc_seq_call* c, const float* dcmds, const float* dparams, float* d_state_cls, float* d_cad_cls,
                 float* d_mv_cls, void* scratch, size_t scratch_bytes, stream_t st) {
  VC_TRY(check_call(c));
  if (!dcmds || !dparams || !d_cad_cls || !scratch) return set_error("seq_backward: null argument");
  const vc_seq_weights& W = *c->w;
  const Dims d = dims_of(c);
  if (d.past_states && !d_state_cls) return set_error("seq_backward: past_states needs d_state_cls");
  if (d.nv > 0 && !d_mv_cls) return set_error("seq_backward: num_views > 0 needs d_mv_cls");
  const int64_t ldc = (int64_t)d.nsrc * d.H;
  const int B = d.B, T = d.T, R = d.R, H = d.H, Ff = d.Ff, P = c->passes;
  Arena arena(c->ws, c->ws_bytes);
  SeqWs w;
  seq_carve(arena, B, T, H, Ff, d.L, d.nh, d.nv, w);
  if (!arena.ok()) return set_error("seq_backward: workspace too small");
  Arena sa(scratch, scratch_bytes);
  SeqScratch s;
  seq_scratch_carve(sa, B, T, H, Ff, d.NP, d.nv, s);
  if (!sa.ok()) return set_error("seq_backward: scratch too small");
  const float p = c->dropout_p;

  float* ui = nullptr; int64_t ld_ui = 0;
  if (d.past_states) { if (d.mem_has_ui) { ui = w.cat; ld_ui = ldc; } else { ui = w.ui; ld_ui = H; } }
  const SeqWs::Layer& last = w.l[d.L - 1];

  // ---- heads: A = d x_last
  VC_TRY(act_dropout_bwd(dparams, d.NP, R, d.NP, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, s.dparS.hi, s.dparS.lo,
                         d.NP, W.head_params.db, st));
  // weight gradients never feed the activation-gradient chain: they go to an auxiliary stream (joined once per layer)
  stream_t side;
  VC_TRY(stream_fork(st, 0, &side));
  VC_TRY(linear_wgrad(s.dparS, last.x3S, R, d.NP, H, W.head_params.dw, P, side));
  {
    // d x_last = dparams [R, 6000] W [6000, H]: K = 6000 on R / 128 x H / 64 tiles is a chain of 94 k-blocks on 16 CTAs (C1: 37 us at
    // the head of the backward's critical path, profiles/r02ab_ncu_full_in_step_summary.txt); split-K spreads it over the chip
    // (atomic accumulation into the zeroed s.A, like the weight-gradient GEMMs).  VC_HEAD_DGRAD_SPLITK=1 restores the single pass.
    static const int head_splitk = [] { const char* e = getenv("VC_HEAD_DGRAD_SPLITK"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : v; }();
    GemmDesc g;
    gemm_linear_dgrad(g, s.dparS, wsplit(W.head_params, H), R, d.NP, H, P);
    g.out_f32 = s.A; g.ldo = H;
    if (head_splitk > 1 && d.NP >= 64 * 4 * head_splitk) {
      VC_TRY(zero_f32(s.A, (int64_t)R * H, st));
      g.splitk = head_splitk;
    }
    VC_TRY(gemm(g, st));
  }
  VC_TRY(head_small_bwd(dcmds, last.x3, R, H, W.head_cmd_w, d.NC, s.A, 1, W.d_head_cmd_w, W.d_head_cmd_b, st));

  bool dmem_init = false;
  for (int l = d.L - 1; l >= 0; --l) {
    const vc_dec_layer& LW = W.layers[l];
    const SeqWs::Layer& Y = w.l[l];
    const uint32_t s0 = c->site_base + 7 * l;
    const Split x_inS = (l > 0) ? w.l[l - 1].x3S
                                : (d.past_actions ? w.actS : (d.past_states ? (d.mem_has_ui ? mk_split(w.catS.hi, w.catS.lo, ldc) : w.uiS) : w.memS));
    // ---- feed-forward block
    // LayerNorm backward with the fused second output: dropout-masked gradient of the sublayer output as split-bf16 operand
    // of its dgrad/wgrad GEMMs + the bias gradient (column sums); s.Y keeps the unmasked d y for the residual path
    VC_TRY(layernorm_bwd_fused(s.A, H, Y.y3, H, Y.m3, Y.r3, LW.n3.w, R, H, nullptr, 0, s.Y, H, LW.n3.dw, LW.n3.db,
                               site_drop(p, c->training, c->seed, s0 + 5, c->seed_dev), s.gH.hi, s.gH.lo, H, LW.lin2.db, st));
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(s.gH, Y.f, R, H, Ff, LW.lin2.dw, P, side));
    {
      // d pre = (g W2) * mask(FFN hidden site) * relu'(f): fused activation backward + linear1 bias gradient
      GemmDesc g;
      gemm_linear_dgrad(g, s.gH, wsplit(LW.lin2, Ff), R, H, Ff, P);
      g.act_backward = 1; g.act = VC_ACT_RELU; g.act_aux_hi = Y.f.hi; g.ld_act_aux_hi = Ff;
      g.drop = site_drop(p, c->training, c->seed, s0 + 4, c->seed_dev);
      g.out_hi = s.dpreF.hi; g.out_lo = s.dpreF.lo; g.ldo_split = Ff; g.colsum = LW.lin1.db;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(s.dpreF, Y.x2S, R, Ff, H, LW.lin1.dw, P, side));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.dpreF, wsplit(LW.lin1, H), R, Ff, H, P);
      g.residual = s.Y; g.ld_res = H; g.out_f32 = s.Bf; g.ldo = H;  // d x2 = d y3 + dpre W1
      VC_TRY(gemm(g, st));
    }
    // ---- cross-attention block
    // LayerNorm backward with the fused second output: dropout-masked gradient of the sublayer output as split-bf16 operand
    // of its dgrad/wgrad GEMMs + the bias gradient (column sums); s.Y keeps the unmasked d y for the residual path
    VC_TRY(layernorm_bwd_fused(s.Bf, H, Y.y2, H, Y.m2, Y.r2, LW.n2.w, R, H, nullptr, 0, s.Y, H, LW.n2.dw, LW.n2.db,
                               site_drop(p, c->training, c->seed, s0 + 3, c->seed_dev), s.gH_ca.hi, s.gH_ca.lo, H, LW.ca_out.db, st));
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(s.gH_ca, Y.c, R, H, H, LW.ca_out.dw, P, side));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.gH_ca, wsplit(LW.ca_out, H), R, H, H, P);
      g.out_f32 = s.dAtt; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    // dq | dk | dv delivered as the split-bf16 operand [R, 3H] of the in_proj dgrad/wgrad GEMMs, in_proj bias gradient included
    VC_TRY(attention_bwd_split_bias(cross_attn_desc(c, d, Y, site_drop(p, c->training, c->seed, s0 + 2, c->seed_dev)), Y.c.hi, Y.c.lo, H,
                                    Y.ca_lse, s.dAtt, H, s.dqkv, s.dqkvS.hi, s.dqkvS.lo, s.dqkvS.hi + H, s.dqkvS.lo + H,
                                    s.dqkvS.hi + 2 * H, s.dqkvS.lo + 2 * H, 3 * H, LW.ca_in.db, LW.ca_in.db ? LW.ca_in.db + H : nullptr,
                                    LW.ca_in.db ? LW.ca_in.db + 2 * H : nullptr, st));
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(cols(s.dqkvS, 0), Y.x1S, R, H, H, LW.ca_in.dw, P, side));
    VC_TRY(linear_wgrad(cols(s.dqkvS, H), w.memS, R, 2 * H, H, LW.ca_in.dw + (size_t)H * H, P, side));
    {
      GemmDesc g;  // d mem accumulates over the layers; only needed after the last layer -> auxiliary stream as well
      gemm_linear_dgrad(g, cols(s.dqkvS, H), wsplit(LW.ca_in, H, H), R, 2 * H, H, P);
      if (dmem_init) { g.residual = s.dmem; g.ld_res = H; }
      g.out_f32 = s.dmem; g.ldo = H;
      VC_TRY(gemm(g, side));
      dmem_init = true;
    }
    {
      GemmDesc g;
      gemm_linear_dgrad(g, cols(s.dqkvS, 0), wsplit(LW.ca_in, H, 0), R, H, H, P);
      g.residual = s.Y; g.ld_res = H; g.out_f32 = s.A; g.ldo = H;  // d x1 = d y2 + dq Wq
      VC_TRY(gemm(g, st));
    }
    // ---- self-attention block
    // LayerNorm backward with the fused second output: dropout-masked gradient of the sublayer output as split-bf16 operand
    // of its dgrad/wgrad GEMMs + the bias gradient (column sums); s.Y keeps the unmasked d y for the residual path
    VC_TRY(layernorm_bwd_fused(s.A, H, Y.y1, H, Y.m1, Y.r1, LW.n1.w, R, H, nullptr, 0, s.Y, H, LW.n1.dw, LW.n1.db,
                               site_drop(p, c->training, c->seed, s0 + 1, c->seed_dev), s.gH_sa.hi, s.gH_sa.lo, H, LW.sa_out.db, st));
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(s.gH_sa, Y.a, R, H, H, LW.sa_out.dw, P, side));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.gH_sa, wsplit(LW.sa_out, H), R, H, H, P);
      g.out_f32 = s.dAtt; g.ldo = H;
      VC_TRY(gemm(g, st));
    }
    VC_TRY(attention_bwd_split_bias(self_attn_desc(c, d, Y, site_drop(p, c->training, c->seed, s0 + 0, c->seed_dev)), Y.a.hi, Y.a.lo, H,
                                    Y.sa_lse, s.dAtt, H, s.dqkv, s.dqkvS_sa.hi, s.dqkvS_sa.lo, s.dqkvS_sa.hi + H, s.dqkvS_sa.lo + H,
                                    s.dqkvS_sa.hi + 2 * H, s.dqkvS_sa.lo + 2 * H, 3 * H, LW.sa_in.db, LW.sa_in.db ? LW.sa_in.db + H : nullptr,
                                    LW.sa_in.db ? LW.sa_in.db + 2 * H : nullptr, st));
    VC_TRY(stream_fork(st, 0, &side));
    VC_TRY(linear_wgrad(s.dqkvS_sa, x_inS, R, 3 * H, H, LW.sa_in.dw, P, side));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.dqkvS_sa, wsplit(LW.sa_in, H), R, 3 * H, H, P);
      g.residual = s.Y; g.ld_res = H; g.out_f32 = s.A; g.ldo = H;  // d x_in = d y1 + dqkv W_in
      VC_TRY(gemm(g, st));
    }
    VC_TRY(stream_join(st, 0));  // the scratch operands of this layer's weight-gradient GEMMs may be overwritten from here on
  }

  // ---- token construction backward: s.A = d tgt, s.dmem = d memory
  float* dE = W.d_timestep_emb;  // may be null
  const float* dui = nullptr; int64_t ld_dui = 0;
  if (d.past_actions) {
    VC_TRY(embed_action_bwd(s.A, w.act, c->actions, R, c->act_dim, H, T, W.d_embed_action_w, W.d_embed_action_b,
                            W.timestep_emb ? dE : nullptr, st));
  } else if (d.past_states) {
    dui = s.A; ld_dui = H;
  } else {
    VC_TRY(add_f32(s.dmem, s.A, s.dmem, (int64_t)R * H, st));
  }
  VC_TRY(zero_f32(s.dcadtok, (int64_t)B * H, st));
  if (d.nsrc > 1) {
    // mem = tanh(image_projection([ui ; cad ; multiview]))
    VC_TRY(act_dropout_bwd(s.dmem, H, R, H, VC_ACT_TANH, w.mem, H, nullptr, 0, no_drop(), nullptr, 0, s.gH.hi, s.gH.lo, H,
                           W.image_proj.db, st));
    VC_TRY(linear_wgrad(s.gH, mk_split(w.catS.hi, w.catS.lo, ldc), R, H, (int)ldc, W.image_proj.dw, P, st));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.gH, wsplit(W.image_proj, ldc), R, H, (int)ldc, P);
      g.out_f32 = s.dcat; g.ldo = ldc;
      VC_TRY(gemm(g, st));
    }
    if (d.mem_has_ui) { dui = s.dcat; ld_dui = ldc; }
    VC_TRY(row_reduce_mod(s.dcat + (int64_t)d.cad_off * H, ldc, R, H, T, B, s.dcadtok, st));
    if (d.nv > 0) {
      const int Kmv = d.nv * VD;
      VC_TRY(zero_f32(s.dmvtok, (int64_t)B * H, st));
      VC_TRY(row_reduce_mod(s.dcat + (int64_t)d.mv_off * H, ldc, R, H, T, B, s.dmvtok, st));
      VC_TRY(act_dropout_bwd(s.dmvtok, H, B, H, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, s.gB2.hi, s.gB2.lo, H,
                             W.embed_multiview.db, st));
      VC_TRY(linear_wgrad(s.gB2, w.mvS, B, H, Kmv, W.embed_multiview.dw, P, st));
      GemmDesc g;
      gemm_linear_dgrad(g, s.gB2, wsplit(W.embed_multiview, Kmv), B, H, Kmv, P);
      g.out_f32 = d_mv_cls; g.ldo = Kmv;
      VC_TRY(gemm(g, st));
    }
  } else {
    VC_TRY(row_reduce_mod(s.dmem, H, R, H, T, B, s.dcadtok, st));
  }
  if (d.past_states) {
    // ui = tanh(embed_state(cls) + E[t])
    VC_TRY(act_dropout_bwd(dui, ld_dui, R, H, VC_ACT_TANH, ui, ld_ui, nullptr, 0, no_drop(), s.Y, H, s.gH.hi, s.gH.lo, H,
                           W.embed_state.db, st));
    if (W.timestep_emb && dE) VC_TRY(row_reduce_mod(s.Y, H, R, H, 1, T, dE, st));
    VC_TRY(linear_wgrad(s.gH, w.scls, R, H, VD, W.embed_state.dw, P, st));
    {
      GemmDesc g;
      gemm_linear_dgrad(g, s.gH, wsplit(W.embed_state, VD), R, H, VD, P);
      g.out_f32 = d_state_cls; g.ldo = VD;
      VC_TRY(gemm(g, st));
    }
  }
  // cad token: embed_image (with tanh folded in when there is no projection)
  VC_TRY(act_dropout_bwd(s.dcadtok, H, B, H, d.nsrc > 1 ? VC_ACT_NONE : VC_ACT_TANH, w.cad_tok, H, nullptr, 0, no_drop(), nullptr, 0,
                         s.gB.hi, s.gB.lo, H, W.embed_image.db, st));
  VC_TRY(linear_wgrad(s.gB, w.ccls, B, H, VD, W.embed_image.dw, P, st));
  {
    GemmDesc g;
    gemm_linear_dgrad(g, s.gB, wsplit(W.embed_image, VD), B, H, VD, P);
    g.out_f32 = d_cad_cls; g.ldo = VD;
    VC_TRY(gemm(g, st));
  }
  return 0;
}

}  // namespace vck

extern "C" {
size_t vc_seq_workspace_bytes(int B, int T, int H, int Ff, int num_layers, int nhead, int num_param_out, int num_views) {
  (void)num_param_out;
  return vck::seq_workspace_bytes(B, T, H, Ff, num_layers, nhead, num_views);
}
size_t vc_seq_scratch_bytes(int B, int T, int H, int Ff, int num_param_out, int num_views) {
  return vck::seq_scratch_bytes(B, T, H, Ff, num_param_out, num_views);
}
int vc_seq_forward(const vc_seq_call* c, void* stream) { return vck::seq_forward(c, stream); }
size_t vc_seq_decode_scratch_bytes(int B, int H, int Ff, int nhead) { return vck::seq_decode_scratch_bytes(B, H, Ff, nhead); }
int vc_seq_decode_step(const vc_seq_call* c, int t, const float* actions_t, void* scratch, size_t scratch_bytes, float* cmds_t,
                       float* params_t, void* stream) {
  return vck::seq_decode_step(c, t, actions_t, scratch, scratch_bytes, cmds_t, params_t, stream);
}
int vc_seq_decode_dev_supported(const vc_seq_call* c) { return vck::seq_decode_dev_supported(c); }
size_t vc_seq_decode_dev_scratch_bytes(int B, int H, int Ff, int nhead) { return vck::seq_decode_dev_scratch_bytes(B, H, Ff, nhead); }
int vc_seq_decode_step_dev(const vc_seq_call* c, int* t_dev, float* actions_io, void* scratch, size_t scratch_bytes, float* cmds_all,
                           float* params_all, void* stream) {
  return vck::seq_decode_step_dev(c, t_dev, actions_io, scratch, scratch_bytes, cmds_all, params_all, stream);
}
int vc_seq_backward(const vc_seq_call* c, const float* dcmds, const float* dparams, float* d_state_cls, float* d_cad_cls,
                    float* d_mv_cls, void* scratch, size_t scratch_bytes, void* stream) {
  return vck::seq_backward(c, dcmds, dparams, d_state_cls, d_cad_cls, d_mv_cls, scratch, scratch_bytes, stream);
}
}
