"""Functional CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Restates, with plain torch tensor ops on a flat ``state_dict`` (key schema = SURVEY.md App. B):

  * ``vit_forward``      vit_pytorch.ViT as configured at /root/reference/model/trajectory_model.py:54-67
                         (third-party, absent from the tree: see oracle/shims/vit_pytorch -- PARITY
                         UNPINNED for that dependency; call sites trajectory_model.py:90-100)
  * ``decoder_layer``    torch.nn.TransformerDecoderLayer post-norm/ReLU path as built at
                         /root/reference/model/autoregressive_transformer.py:54-62
  * ``forward``          AutoRegressiveTransformer.forward, autoregressive_transformer.py:121-220
  * ``apply_action_mask`` / ``normalize_actions``   autoregressive_transformer.py:91-118
  * ``rollout``          sequential_inference with the intended action-feedback semantics,
                         autoregressive_transformer.py:222-275 (the shipped ``action=True`` path raises
                         IndexError; this follows SURVEY.md 3.3 / App. D.1: masks applied on [B,1,*])

Everything runs in the dtype of the tensors handed in (fp32, or fp64 for a "truth" run).  Dropout is
the identity (eval mode): training-mode parity is statistical and tested separately.

Pinned against the real reference by tests/test_oracle_vs_reference.py (build container only) and by
tests/golden/*.npz (outputs of the unmodified reference; generator: oracle/make_golden.py).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Optional

import torch
import torch.nn.functional as F

VIT_DIM = 512
VIT_DEPTH = 6
VIT_HEADS = 16
VIT_DHEAD = 64
VIT_MLP = 512
PATCH = 32
LN_EPS = 1e-5

DEFAULTS = dict(
    act_dim=7, hidden_size=256, max_ep_len=1000, enable_past_actions=False, enable_past_states=False,
    enable_timestep_embedding=False, num_classes=5, num_params=6, num_params_values=1000,
    num_decoder_layers=8, dim_feedforward=512, nhead=4, dropout=0.1, num_views=0, window_size=1,
)


def full_cfg(cfg: dict) -> dict:
    out = dict(DEFAULTS)
    out.update({k: v for k, v in cfg.items() if k in DEFAULTS})
    return out


# --------------------------------------------------------------------------------------------
# deterministic weights that do not depend on module construction order
# --------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """Shapes of every LIVE parameter of AutoRegressiveTransformer."""
    c = full_cfg(cfg)
    H, Ff, L = c["hidden_size"], c["dim_feedforward"], c["num_decoder_layers"]
    shapes: Dict[str, tuple] = {}
    for vit in ("state_embedding_model.", "cad_embedding_model."):
        shapes[vit + "pos_embedding"] = (1, 50, VIT_DIM)
        shapes[vit + "cls_token"] = (1, 1, VIT_DIM)
        shapes[vit + "to_patch_embedding.1.weight"] = (PATCH * PATCH,)
        shapes[vit + "to_patch_embedding.1.bias"] = (PATCH * PATCH,)
        shapes[vit + "to_patch_embedding.2.weight"] = (VIT_DIM, PATCH * PATCH)
        shapes[vit + "to_patch_embedding.2.bias"] = (VIT_DIM,)
        shapes[vit + "to_patch_embedding.3.weight"] = (VIT_DIM,)
        shapes[vit + "to_patch_embedding.3.bias"] = (VIT_DIM,)
        for l in range(VIT_DEPTH):
            p = f"{vit}transformer.layers.{l}."
            shapes[p + "0.norm.weight"] = (VIT_DIM,)
            shapes[p + "0.norm.bias"] = (VIT_DIM,)
            shapes[p + "0.to_qkv.weight"] = (3 * VIT_HEADS * VIT_DHEAD, VIT_DIM)
            shapes[p + "0.to_out.0.weight"] = (VIT_DIM, VIT_HEADS * VIT_DHEAD)
            shapes[p + "0.to_out.0.bias"] = (VIT_DIM,)
            shapes[p + "1.net.0.weight"] = (VIT_DIM,)
            shapes[p + "1.net.0.bias"] = (VIT_DIM,)
            shapes[p + "1.net.1.weight"] = (VIT_MLP, VIT_DIM)
            shapes[p + "1.net.1.bias"] = (VIT_MLP,)
            shapes[p + "1.net.4.weight"] = (VIT_DIM, VIT_MLP)
            shapes[p + "1.net.4.bias"] = (VIT_DIM,)
        shapes[vit + "transformer.norm.weight"] = (VIT_DIM,)
        shapes[vit + "transformer.norm.bias"] = (VIT_DIM,)
    num_inputs = 1 + (1 if c["enable_past_states"] else 0) + (1 if c["num_views"] > 0 else 0)
    shapes["embed_state.weight"] = (H, VIT_DIM)
    shapes["embed_state.bias"] = (H,)
    shapes["embed_image.weight"] = (H, VIT_DIM)
    shapes["embed_image.bias"] = (H,)
    shapes["image_projection.weight"] = (H, H * num_inputs)
    shapes["image_projection.bias"] = (H,)
    shapes["embed_action.weight"] = (H, c["act_dim"])
    shapes["embed_action.bias"] = (H,)
    if c["enable_timestep_embedding"]:
        shapes["timestep_embedding.weight"] = (c["max_ep_len"], H)
    if c["num_views"] > 0:
        shapes["embed_multiview.weight"] = (H, VIT_DIM * c["num_views"])
        shapes["embed_multiview.bias"] = (H,)
    for l in range(L):
        p = f"transformer_decoder.layers.{l}."
        for att in ("self_attn.", "multihead_attn."):
            shapes[p + att + "in_proj_weight"] = (3 * H, H)
            shapes[p + att + "in_proj_bias"] = (3 * H,)
            shapes[p + att + "out_proj.weight"] = (H, H)
            shapes[p + att + "out_proj.bias"] = (H,)
        shapes[p + "linear1.weight"] = (Ff, H)
        shapes[p + "linear1.bias"] = (Ff,)
        shapes[p + "linear2.weight"] = (H, Ff)
        shapes[p + "linear2.bias"] = (H,)
        for n in ("norm1.", "norm2.", "norm3."):
            shapes[p + n + "weight"] = (H,)
            shapes[p + n + "bias"] = (H,)
    shapes["predict_action_class_0_4.weight"] = (c["num_classes"], H)
    shapes["predict_action_class_0_4.bias"] = (c["num_classes"],)
    shapes["predict_action_class_0_999.weight"] = (c["num_params"] * c["num_params_values"], H)
    shapes["predict_action_class_0_999.bias"] = (c["num_params"] * c["num_params_values"],)
    return shapes


def _is_norm_key(key: str) -> bool:
    parts = key.split(".")
    if "norm" in parts[-2] or parts[-2] in ("norm1", "norm2", "norm3"):
        return True
    # vit: to_patch_embedding.{1,3} and ff net.0 are LayerNorms
    if "to_patch_embedding" in key and parts[-2] in ("1", "3"):
        return True
    if ".net.0." in key:
        return True
    return False


def seeded_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, construction-order-independent weights: every tensor comes from its own
    generator seeded by (seed, crc32(key)).  LayerNorm affine params and all biases are non-trivial
    on purpose (default init would hide gamma/beta bugs)."""
    sd = {}
    for key, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        if _is_norm_key(key):
            t = torch.randn(shape, generator=g) * 0.1
            if key.endswith("weight"):
                t = t + 1.0
        elif key.endswith("pos_embedding") or key.endswith("cls_token") or key.startswith("timestep_embedding"):
            t = torch.randn(shape, generator=g)
        elif len(shape) == 2:
            t = torch.randn(shape, generator=g) * (0.6 / math.sqrt(shape[1]))
        else:  # biases
            t = torch.randn(shape, generator=g) * 0.05
        sd[key] = t.to(dtype)
    return sd


from videocad_b200.synthetic import synthetic_batch  # noqa: E402,F401  (shared synthetic-input generator)


def synthetic_views(B: int, num_views: int, S: int, seed: int = 4321) -> torch.Tensor:
    """Synthetic multiview CAD renderings [B, num_views, 1, S, S] in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, num_views, 1, S, S, generator=g).clamp_(-1, 1)


def normalize_actions(actions: torch.Tensor) -> torch.Tensor:
    """autoregressive_transformer.py:115-118 (out-of-place here)."""
    out = actions.clone()
    out[:, :, 0] = out[:, :, 0] / 4.0
    out[:, :, 1:] = out[:, :, 1:] / 1000.0
    return out


def model_inputs_from_batch(batch: dict) -> dict:
    """trainer.py:507-517: model sees frames[:, :-1], normalised actions[:, :-1]; targets actions[:, 1:]."""
    return {
        "frames": batch["frames"][:, :-1],
        "actions": normalize_actions(batch["actions"][:, :-1]),
        "cad_image": batch["cad_image"],
    }


# --------------------------------------------------------------------------------------------
# ViT
# --------------------------------------------------------------------------------------------
def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], LN_EPS)


def patchify(img: torch.Tensor) -> torch.Tensor:
    """'b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p1=p2=32."""
    b, c, hh, ww = img.shape
    h, w = hh // PATCH, ww // PATCH
    x = img.reshape(b, c, h, PATCH, w, PATCH).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(b, h * w, PATCH * PATCH * c)


def _drop(x, p):
    return F.dropout(x, p, training=True) if p > 0 else x


def vit_forward(sd: dict, prefix: str, img: torch.Tensor, inter: Optional[dict] = None, dropout_p: float = 0.0) -> torch.Tensor:
    """[F,1,S,S] -> CLS embedding [F,512].  SURVEY.md App. A.1.  dropout_p > 0 = training mode (statistical only)."""
    dp = dropout_p
    x = patchify(img)
    x = _ln(x, sd, prefix + "to_patch_embedding.1.")
    x = x @ sd[prefix + "to_patch_embedding.2.weight"].T + sd[prefix + "to_patch_embedding.2.bias"]
    x = _ln(x, sd, prefix + "to_patch_embedding.3.")
    b, n, _ = x.shape
    cls = sd[prefix + "cls_token"].reshape(1, 1, VIT_DIM).expand(b, 1, VIT_DIM)
    pos = sd[prefix + "pos_embedding"].reshape(-1, VIT_DIM)
    x = _drop(torch.cat([cls, x], dim=1) + pos[: n + 1], dp)
    if inter is not None:
        inter[prefix + "tokens"] = x
    for l in range(VIT_DEPTH):
        p = f"{prefix}transformer.layers.{l}."
        h = _ln(x, sd, p + "0.norm.")
        qkv = h @ sd[p + "0.to_qkv.weight"].T
        q, k, v = qkv.chunk(3, dim=-1)
        q, k, v = (t.reshape(b, n + 1, VIT_HEADS, VIT_DHEAD).permute(0, 2, 1, 3) for t in (q, k, v))
        a = _drop(torch.softmax((q @ k.transpose(-1, -2)) * (VIT_DHEAD ** -0.5), dim=-1), dp)
        o = (a @ v).permute(0, 2, 1, 3).reshape(b, n + 1, VIT_HEADS * VIT_DHEAD)
        x = _drop(o @ sd[p + "0.to_out.0.weight"].T + sd[p + "0.to_out.0.bias"], dp) + x
        h = _ln(x, sd, p + "1.net.0.")
        u = _drop(F.gelu(h @ sd[p + "1.net.1.weight"].T + sd[p + "1.net.1.bias"]), dp)
        x = _drop(u @ sd[p + "1.net.4.weight"].T + sd[p + "1.net.4.bias"], dp) + x
        if inter is not None:
            inter[f"{prefix}layer{l}"] = x
    x = _ln(x, sd, prefix + "transformer.norm.")
    return x[:, 0]


# --------------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------------
def build_masks(T: int, W: int, dtype, device=None):
    """autoregressive_transformer.py:180-188."""
    rows = torch.arange(T, device=device)[:, None]
    cols = torch.arange(T, device=device)[None, :]
    neg = torch.full((T, T), float("-inf"), dtype=dtype, device=device)
    zero = torch.zeros((T, T), dtype=dtype, device=device)
    causal = torch.where(cols <= rows, zero, neg)
    window = torch.where((cols > rows - W) & (cols <= rows), zero, neg)
    return causal, window


def _mha(xq, xkv, sd, prefix, nh, mask, dp=0.0):
    """nn.MultiheadAttention (packed in_proj), batch-first restatement: [B,T,H]."""
    B, Tq, H = xq.shape
    Tk = xkv.shape[1]
    d = H // nh
    w, bias = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = xq @ w[:H].T + bias[:H]
    k = xkv @ w[H:2 * H].T + bias[H:2 * H]
    v = xkv @ w[2 * H:].T + bias[2 * H:]
    q = q.reshape(B, Tq, nh, d).permute(0, 2, 1, 3)
    k = k.reshape(B, Tk, nh, d).permute(0, 2, 1, 3)
    v = v.reshape(B, Tk, nh, d).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d) + mask
    a = _drop(torch.softmax(s, dim=-1), dp)
    o = (a @ v).permute(0, 2, 1, 3).reshape(B, Tq, H)
    return o @ sd[prefix + "out_proj.weight"].T + sd[prefix + "out_proj.bias"]


def decoder_layer(sd, prefix, x, mem, nh, tgt_mask, mem_mask, dp=0.0):
    x = _ln(x + _drop(_mha(x, x, sd, prefix + "self_attn.", nh, tgt_mask, dp), dp), sd, prefix + "norm1.")
    x = _ln(x + _drop(_mha(x, mem, sd, prefix + "multihead_attn.", nh, mem_mask, dp), dp), sd, prefix + "norm2.")
    ff = _drop(torch.relu(x @ sd[prefix + "linear1.weight"].T + sd[prefix + "linear1.bias"]), dp)
    ff = _drop(ff @ sd[prefix + "linear2.weight"].T + sd[prefix + "linear2.bias"], dp)
    return _ln(x + ff, sd, prefix + "norm3.")


# --------------------------------------------------------------------------------------------
# full forward
# --------------------------------------------------------------------------------------------
def forward(sd: dict, cfg: dict, inputs: dict, inter: Optional[dict] = None, dropout_p: float = 0.0):
    """AutoRegressiveTransformer.forward.  Returns (cmds [B,T,5], params [B,T,6,1000]).
    dropout_p = 0 is eval mode (bit-level parity runs); > 0 applies dropout at the reference's call sites (used only
    by the timed CPU port of the training step)."""
    c = full_cfg(cfg)
    dp = dropout_p
    H, nh, L, W = c["hidden_size"], c["nhead"], c["num_decoder_layers"], c["window_size"]
    frames, actions, cad = inputs["frames"], inputs["actions"], inputs["cad_image"]
    B, T = actions.shape[0], actions.shape[1]
    dtype = cad.dtype
    if c["enable_timestep_embedding"]:
        E = sd["timestep_embedding.weight"][:T]
    else:
        E = torch.zeros(T, H, dtype=dtype, device=cad.device)

    images = []
    ui = None
    if c["enable_past_states"]:
        emb = vit_forward(sd, "state_embedding_model.", frames.reshape(-1, *frames.shape[2:]), inter, dp)
        if inter is not None:
            inter["state_cls"] = emb
        ui = torch.tanh((emb @ sd["embed_state.weight"].T + sd["embed_state.bias"]).reshape(B, T, H) + E)
        if c["enable_past_actions"]:
            images.append(ui)
    cad_emb = vit_forward(sd, "cad_embedding_model.", cad, inter, dp)
    if inter is not None:
        inter["cad_cls"] = cad_emb
    cad_tok = (cad_emb @ sd["embed_image.weight"].T + sd["embed_image.bias"]).unsqueeze(1).expand(B, T, H)
    images.append(cad_tok)
    mv = inputs.get("multiview_images", None)
    if mv is not None and c["num_views"] > 0:
        nv = mv.shape[1]
        mv_emb = vit_forward(sd, "cad_embedding_model.", mv.reshape(-1, *mv.shape[2:]), None, dp).reshape(B, nv * VIT_DIM)
        mv_tok = mv_emb @ sd["embed_multiview.weight"].T + sd["embed_multiview.bias"]
        images.append(mv_tok.unsqueeze(1).expand(B, T, H))
    mem = torch.cat(images, dim=-1)
    if len(images) > 1:
        mem = mem @ sd["image_projection.weight"].T + sd["image_projection.bias"]
    mem = torch.tanh(mem)
    tgt_act = torch.tanh(actions.to(dtype) @ sd["embed_action.weight"].T + sd["embed_action.bias"] + E)
    causal, window = build_masks(T, W, dtype, cad.device)
    if c["enable_past_actions"]:
        x, tgt_mask = tgt_act, causal
    elif c["enable_past_states"]:
        x, tgt_mask = ui, window
    else:
        x, tgt_mask = mem, window
    if inter is not None:
        inter["tgt"], inter["memory"] = x, mem
    for l in range(L):
        x = decoder_layer(sd, f"transformer_decoder.layers.{l}.", x, mem, nh, tgt_mask, window, dp)
        if inter is not None:
            inter[f"dec{l}"] = x
    cmds = x @ sd["predict_action_class_0_4.weight"].T + sd["predict_action_class_0_4.bias"]
    params = x @ sd["predict_action_class_0_999.weight"].T + sd["predict_action_class_0_999.bias"]
    return cmds, params.reshape(B, T, c["num_params"], c["num_params_values"])


# --------------------------------------------------------------------------------------------
# rollout
# --------------------------------------------------------------------------------------------
ACTION_MASK = torch.tensor(
    [[1, 1, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 0]]
).float()


def apply_action_mask(cmd_pred: torch.Tensor, param_pred: torch.Tensor) -> torch.Tensor:
    """autoregressive_transformer.py:91-108 on [B,S] / [B,S,6] integer tensors."""
    mask = ACTION_MASK.to(param_pred.device)[cmd_pred]
    out = param_pred.clone()
    out[mask == 0] = -1
    out[:, :, 3] = torch.where((out[:, :, 2] >= 200) & (out[:, :, 2] < 250), out[:, :, 3], -1)
    return out


def rollout(sd: dict, cfg: dict, frames: torch.Tensor, cad: torch.Tensor, action: bool = True):
    """sequential_inference (autoregressive_transformer.py:222-275) by full recompute -- O(T^2)."""
    B, T = frames.shape[:2]
    acts = torch.zeros(B, 1, 7, dtype=frames.dtype, device=frames.device)
    out_c, out_p = [], []
    for t in range(T):
        a_in = acts if action else torch.zeros(B, t + 1, 7, dtype=frames.dtype, device=frames.device)
        cmd, par = forward(sd, cfg, {"frames": frames[:, : t + 1], "actions": a_in, "cad_image": cad})
        out_c.append(cmd[:, -1])
        out_p.append(par[:, -1])
        if action:
            cp = cmd[:, -1].argmax(-1)
            pp = par[:, -1].argmax(-1)
            nxt = apply_action_mask(cp.unsqueeze(1), pp.unsqueeze(1)).to(frames.dtype)
            nxt = torch.cat([cp.reshape(B, 1, 1).to(frames.dtype), nxt], dim=2)
            acts = torch.cat([acts, normalize_actions(nxt)], dim=1)
    return torch.stack(out_c, 1), torch.stack(out_p, 1)
