#!/usr/bin/env python
"""Per-weight gradient error of the CUDA path against the fp64 oracle at the C3 per-GPU model shape (diagnostic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import torch_oracle as to  # noqa: E402
from videocad_b200 import AutoRegressiveTransformer  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    B, S = 4, 224
    cfg = dict(hidden_size=1024, nhead=4, num_decoder_layers=L, dim_feedforward=1024, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, dropout=0.0, vit_dropout=0.0, encoder="vit", **cfg)
    sd = to.seeded_state_dict(cfg, 0)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().train()
    batch = to.synthetic_batch(B, T + 1, S, seed=1234)
    inp = {k: v.cuda() for k, v in to.model_inputs_from_batch(batch).items()}
    g = torch.Generator().manual_seed(5)
    wc, wp = torch.randn((B, T, 5), generator=g).cuda(), (torch.randn((B, T, 6, 1000), generator=g) * 0.05).cuda()
    sdd = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    ((oc * wc.double()).sum() + (op * wp.double()).sum()).backward()
    # fp32 torch run of the same oracle: how far does plain fp32 arithmetic land from fp64?
    sd32 = {k: v.float().cuda().requires_grad_(True) for k, v in sd.items()}
    c32, p32 = to.forward(sd32, cfg, inp)
    ((c32 * wc).sum() + (p32 * wp).sum()).backward()
    cmds, params = m(inp)
    ((cmds * wc).sum() + (params * wp).sum()).backward()
    print(f"L={L} T={T}: logits max|d| ours {(cmds.double() - oc).abs().max().item():.3e} / {(params.double() - op).abs().max().item():.3e}; "
          f"torch fp32 {(c32.double() - oc).abs().max().item():.3e} / {(p32.double() - op).abs().max().item():.3e}")
    rows = []
    for k, p in m.named_weights():
        ref = sdd[k].grad
        if ref is None or p.grad is None:
            continue
        den = ref.abs().max().item() + 1e-6
        rows.append(((p.grad.double() - ref).abs().max().item() / den, (sd32[k].grad.double() - ref).abs().max().item() / den, den, k))
    rows.sort(reverse=True)
    print("rel err ours | rel err torch fp32 | max|ref| | key")
    for r in rows[:25]:
        print(f"{r[0]:.3e}  {r[1]:.3e}  {r[2]:.3e}  {r[3]}")


if __name__ == "__main__":
    main()
