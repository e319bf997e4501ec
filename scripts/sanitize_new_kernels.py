#!/usr/bin/env python
"""Small invocations of the kernels added in this round, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_new_kernels.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kwrap as K  # noqa: E402
from videocad_b200 import lib as L  # noqa: E402
from videocad_b200.loss import compute_loss_and_metrics_fused  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
lib = L.load()
# short-sequence decoder attention, forward + fused backward (both head-dim variants, both masks)
for (B, T, nh, d, mask, window) in [(3, 8, 4, 128, L.MASK_CAUSAL, 1), (2, 32, 2, 256, L.MASK_WINDOW, 10), (2, 5, 4, 64, L.MASK_WINDOW, 2)]:
    H = nh * d
    x = torch.randn(B * T, 3 * H, device="cuda", generator=g)
    a = K.attn_desc(x[:, :H], x[:, H:2 * H], x[:, 2 * H:], B, T, T, nh, d, mask=mask, window=window, drop=L.make_drop(0.1, 4, 9))
    o, lse = K.attention_fwd(a, B, T, nh, d)
    K.attention_bwd_split_bias(a, o, lse, torch.randn(B * T, H, device="cuda", generator=g), B, T, nh, d)
# LayerNorm backward with the cp.async ring (fused and plain, C = 256 and 512, ragged row counts)
for rows, Cc in [(37, 256), (1000, 512), (5, 512)]:
    xx = torch.randn(rows, Cc, device="cuda", generator=g)
    gam, bet = torch.randn(Cc, device="cuda", generator=g), torch.randn(Cc, device="cuda", generator=g)
    y, ys, mean, rstd = K.layernorm_fwd(xx, gam, bet)
    dy, dres = torch.randn(rows, Cc, device="cuda", generator=g), torch.randn(rows, Cc, device="cuda", generator=g)
    K.layernorm_bwd_fused(dy, xx, mean, rstd, gam, dres, L.make_drop(0.1, 2, 3))
    K.layernorm_bwd(dy, xx, mean, rstd, gam, dres)
# row-kernel Linear, frame ingestion, metrics
for M, N, Kd in [(8, 96, 256), (13, 40, 128)]:
    xx, Wt = torch.randn(M, Kd, device="cuda", generator=g), torch.randn(N, Kd, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    L.check(lib.vc_linear_rows_fwd(xx.data_ptr(), None, None, Kd, M, Wt.data_ptr(), None, N, Kd, L.ACT_RELU, None, 0, out.data_ptr(), N, None, None, 0,
                                  L.cur_stream()))
u8 = torch.randint(0, 256, (733,), dtype=torch.uint8, device="cuda", generator=g)
dst = torch.empty(733, device="cuda")
L.check(lib.vc_frames_u8_normalize(u8.data_ptr(), 733, 0.5, 0.5, dst.data_ptr(), L.cur_stream()))
cm, pa = torch.randn(3, 4, 5, device="cuda", generator=g), torch.randn(3, 4, 6, 1000, device="cuda", generator=g)
tg = torch.cat([torch.randint(0, 5, (3, 4, 1), device="cuda", generator=g), torch.randint(-1, 1000, (3, 4, 6), device="cuda", generator=g)], -1).float()
compute_loss_and_metrics_fused((cm, pa), tg)
# single-CTA GEMM with the epilogue operand prefetch (residual + dropout, 64- and 128-wide tiles) and the 256 x 128 pair tile
for M, N, Kd in [(256, 512, 128), (300, 136, 64), (2560, 384, 64)]:
    A_, B_ = torch.randn(M, Kd, device="cuda", generator=g), torch.randn(N, Kd, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    L.gemm(L.split(A_), L.split(B_), M, N, Kd, bias=torch.randn(N, device="cuda", generator=g), drop=L.make_drop(0.1, 5, 1),
           residual=torch.randn(M, N, device="cuda", generator=g), out_f32=out)
torch.cuda.synchronize()
print("sanitize_new_kernels: done")
