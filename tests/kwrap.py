"""Thin per-kernel wrappers over the C ABI for the unit parity tests (tensors in, tensors out)."""
import ctypes as C
import math

import torch

from videocad_b200 import lib as L


def bf16_pair(shape, device="cuda"):
    return (torch.empty(shape, dtype=torch.bfloat16, device=device), torch.empty(shape, dtype=torch.bfloat16, device=device))


def join(pair):
    return pair[0].float() + pair[1].float()


def ref_split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def layernorm_fwd(x, gamma, beta, eps=1e-5, want_f32=True, want_split=True):
    rows, Cc = x.shape
    y = torch.empty_like(x) if want_f32 else None
    ys = bf16_pair(x.shape) if want_split else (None, None)
    mean = torch.empty(rows, device=x.device)
    rstd = torch.empty(rows, device=x.device)
    L.check(L.load().vc_layernorm_fwd(L.ptr(x), x.stride(0), rows, Cc, L.ptr(gamma), L.ptr(beta), eps, L.ptr(y), Cc,
                                     L.ptr(ys[0]), L.ptr(ys[1]), Cc, L.ptr(mean), L.ptr(rstd), L.cur_stream()))
    return y, ys, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None):
    rows, Cc = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, device=x.device)
    db = torch.zeros(Cc, device=x.device)
    L.check(L.load().vc_layernorm_bwd(L.ptr(dy), dy.stride(0), L.ptr(x), x.stride(0), L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                     rows, Cc, L.ptr(dres), Cc if dres is not None else 0, L.ptr(dx), Cc, L.ptr(dg),
                                     L.ptr(db), L.cur_stream()))
    return dx, dg, db


def layernorm_bwd_fused(dy, x, mean, rstd, gamma, dres, drop):
    rows, Cc = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, device=x.device)
    db = torch.zeros(Cc, device=x.device)
    gs = bf16_pair(x.shape)
    cs = torch.zeros(Cc, device=x.device)
    L.check(L.load().vc_layernorm_bwd_fused(L.ptr(dy), dy.stride(0), L.ptr(x), x.stride(0), L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                           rows, Cc, L.ptr(dres), Cc, L.ptr(dx), Cc, L.ptr(dg), L.ptr(db), drop, L.ptr(gs[0]),
                                           L.ptr(gs[1]), Cc, L.ptr(cs), L.cur_stream()))
    return dx, dg, db, gs, cs


def patch_layernorm_fwd(img, gamma, beta, eps=1e-5):
    F, _, S, _ = img.shape
    N = (S // 32) ** 2
    ys = bf16_pair((F * N, 1024))
    mean = torch.empty(F * N, device=img.device)
    rstd = torch.empty(F * N, device=img.device)
    L.check(L.load().vc_patch_layernorm_fwd(L.ptr(img), F, S, L.ptr(gamma), L.ptr(beta), eps, L.ptr(ys[0]), L.ptr(ys[1]),
                                           L.ptr(mean), L.ptr(rstd), L.cur_stream()))
    return ys, mean, rstd


def patch_layernorm_bwd_params(img, mean, rstd, dy):
    F, _, S, _ = img.shape
    dg = torch.zeros(1024, device=img.device)
    db = torch.zeros(1024, device=img.device)
    L.check(L.load().vc_patch_layernorm_bwd_params(L.ptr(img), F, S, L.ptr(mean), L.ptr(rstd), L.ptr(dy), L.ptr(dg), L.ptr(db),
                                                  L.cur_stream()))
    return dg, db


def dropout_mask(drop, n):
    out = torch.empty(n, device="cuda")
    L.check(L.load().vc_dropout_mask_debug(drop, n, L.ptr(out), L.cur_stream()))
    return out


def vit_assemble_fwd(e, F, N, Cc, cls, pos, drop):
    x = torch.empty(F * (N + 1), Cc, device=e.device)
    L.check(L.load().vc_vit_assemble_fwd(L.ptr(e), F, N, Cc, L.ptr(cls), L.ptr(pos), drop, L.ptr(x), L.cur_stream()))
    return x


def vit_assemble_bwd(dx, F, N, Cc, drop):
    de = torch.empty(F * N, Cc, device=dx.device)
    dcls = torch.zeros(Cc, device=dx.device)
    dpos = torch.zeros(N + 1, Cc, device=dx.device)
    L.check(L.load().vc_vit_assemble_bwd(L.ptr(dx), F, N, Cc, drop, L.ptr(de), L.ptr(dcls), L.ptr(dpos), L.cur_stream()))
    return de, dcls, dpos


def attn_desc(q, k, v, B, Tq, Tk, nh, d, mask=L.MASK_NONE, window=1, scale=None, drop=None):
    a = L.AttnDesc()
    a.q, a.k, a.v = L.ptr(q), L.ptr(k), L.ptr(v)
    a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
    a.B, a.Tq, a.Tk, a.nh, a.d = B, Tq, Tk, nh, d
    a.mask, a.window = mask, window
    a.scale = scale if scale is not None else 1.0 / math.sqrt(d)
    a.drop = drop if drop is not None else L.make_drop()
    return a


def attention_fwd(a, B, Tq, nh, d):
    o = bf16_pair((B * Tq, nh * d))
    lse = torch.empty(B, nh, Tq, device="cuda")
    L.check(L.load().vc_attention_fwd(C.byref(a), L.ptr(o[0]), L.ptr(o[1]), nh * d, L.ptr(lse), L.cur_stream()))
    return o, lse


def attention_bwd(a, o, lse, dout, B, Tq, Tk, nh, d):
    dq = torch.full((B * Tq, nh * d), float("nan"), device="cuda")
    dk = torch.full((B * Tk, nh * d), float("nan"), device="cuda")
    dv = torch.full((B * Tk, nh * d), float("nan"), device="cuda")
    L.check(L.load().vc_attention_bwd(C.byref(a), L.ptr(o[0]), L.ptr(o[1]), nh * d, L.ptr(lse), L.ptr(dout), dout.stride(0),
                                     L.ptr(dq), nh * d, L.ptr(dk), nh * d, L.ptr(dv), nh * d, L.cur_stream()))
    return dq, dk, dv


def attn_desc_split(qs, ks, vs, B, T, nh, d, scale=None, drop=None):
    """q/k/v given as (hi, lo) pairs of bf16 views with a common row stride"""
    a = L.AttnDesc()
    a.q_hi, a.q_lo, a.k_hi, a.k_lo, a.v_hi, a.v_lo = (L.ptr(t) for t in (qs[0], qs[1], ks[0], ks[1], vs[0], vs[1]))
    a.ldq, a.ldk, a.ldv = qs[0].stride(0), ks[0].stride(0), vs[0].stride(0)
    a.B, a.Tq, a.Tk, a.nh, a.d = B, T, T, nh, d
    a.mask, a.window = L.MASK_NONE, 1
    a.scale = scale if scale is not None else 1.0 / math.sqrt(d)
    a.drop = drop if drop is not None else L.make_drop()
    return a


def attention_bwd_split(a, o, lse, dout_split, B, T, nh, d):
    W = nh * d
    dq, dk, dv = bf16_pair((B * T, W)), bf16_pair((B * T, W)), bf16_pair((B * T, W))
    scratch = torch.empty(3 * B * T * W, device="cuda")
    L.check(L.load().vc_attention_bwd_split(C.byref(a), L.ptr(o[0]), L.ptr(o[1]), W, L.ptr(lse), None, L.ptr(dout_split[0]),
                                           L.ptr(dout_split[1]), dout_split[0].stride(0), L.ptr(scratch), L.ptr(dq[0]), L.ptr(dq[1]),
                                           L.ptr(dk[0]), L.ptr(dk[1]), L.ptr(dv[0]), L.ptr(dv[1]), W, L.cur_stream()))
    return dq, dk, dv


def attention_bwd_split_bias(a, o, lse, dout, B, T, nh, d):
    """dq | dk | dv as ONE split-bf16 [B*T, 3W] operand (the decoder's in_proj layout) + the packed bias gradient [3W]"""
    W = nh * d
    g = bf16_pair((B * T, 3 * W))
    db = torch.zeros(3 * W, device="cuda")
    scratch = torch.empty(3 * B * T * W, device="cuda")
    L.check(L.load().vc_attention_bwd_split_bias(C.byref(a), L.ptr(o[0]), L.ptr(o[1]), W, L.ptr(lse), L.ptr(dout), dout.stride(0),
                                                L.ptr(scratch), L.ptr(g[0]), L.ptr(g[1]), L.ptr(g[0][:, W:]), L.ptr(g[1][:, W:]),
                                                L.ptr(g[0][:, 2 * W:]), L.ptr(g[1][:, 2 * W:]), 3 * W, L.ptr(db), L.ptr(db[W:]),
                                                L.ptr(db[2 * W:]), L.cur_stream()))
    return g, db


def act_dropout_bwd(dy, act, aux=None, aux_hi=None, drop=None, want_colsum=True):
    M, N = dy.shape
    g = torch.empty_like(dy)
    gs = bf16_pair(dy.shape)
    cs = torch.zeros(N, device=dy.device) if want_colsum else None
    L.check(L.load().vc_act_dropout_bwd(L.ptr(dy), dy.stride(0), M, N, act, L.ptr(aux), N, L.ptr(aux_hi), N,
                                       drop if drop is not None else L.make_drop(), L.ptr(g), N, L.ptr(gs[0]), L.ptr(gs[1]), N,
                                       L.ptr(cs), L.cur_stream()))
    return g, gs, cs
