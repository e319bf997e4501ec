#!/usr/bin/env python
"""Rollout throughput at BASELINE.json configs[4] (C4): 186-step autoregressive rollout of the H = 1024 model, 8 sequences per GPU
(64 over 8 GPUs: sequences are independent, no communication), eval mode, 224 x 224 frames.

    python scripts/rollout_bench.py [--batch 8] [--steps 186] [--iters 2]

Reports frames/s of AutoRegressiveTransformer.sequential_inference for action=False (one pass) and action=True (argmax feedback:
every frame is encoded ONCE, one full-length pass builds the cross-attention keys/values, then one key/value-cached decode step per
position -- the reference re-runs the whole forward, all encoder passes included, on the growing prefix at every step: 17 577 encoder
passes per sample instead of 187).  bench.py reports the same measurement under its `rollout` key."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocad_b200 import AutoRegressiveTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=186)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--only-feedback", action="store_true", help="skip the action=False single pass")
    ap.add_argument("--calls", type=int, default=0, help="> 0: exactly this many untimed-warm-up-free calls (profiling under ncu)")
    args = ap.parse_args()
    cfg = dict(hidden_size=1024, nhead=4, num_decoder_layers=8, dim_feedforward=1024, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    torch.manual_seed(0)
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", **cfg).cuda().eval()
    B, T, S = args.batch, args.steps, 224
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = torch.randn(B, T, 1, S, S, device="cuda", generator=g).clamp_(-1, 1)
    cad = torch.randn(B, 1, S, S, device="cuda", generator=g).clamp_(-1, 1)
    out = {}
    for action in ((True,) if args.only_feedback else (False, True)):
        ts = []
        for it in range(args.calls if args.calls > 0 else args.iters + 2):  # the first two calls run eagerly / capture the CUDA graphs
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cmds, params = m.sequential_inference(frames, cad, action=action)
            e1.record()
            torch.cuda.synchronize()
            if it > 1 or args.calls > 0:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2] if ts else float('nan')
        assert cmds.shape == (B, T, 5) and params.shape == (B, T, 6, 1000) and torch.isfinite(params).all()
        out["action_feedback" if action else "single_pass"] = dict(ms_per_rollout=ms, frames_per_s=B * T / (ms / 1e3))
    print(json.dumps(dict(metric="rollout frames/sec", n_gpus=1, config=dict(workload=f"c4: {B} sequences x {T} steps, H=1024, 224x224, eval"),
                          **out)))


if __name__ == "__main__":
    main()
