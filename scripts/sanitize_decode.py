#!/usr/bin/env python
"""Small rollout through the device-resident decode step (run under compute-sanitizer: memcheck / racecheck / synccheck)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VIDEOCAD_B200_GRAPHS", "0")
from oracle import torch_oracle as to  # noqa: E402
from videocad_b200 import AutoRegressiveTransformer  # noqa: E402

for H, nh, B, T in ((256, 4, 2, 5), (512, 4, 9, 4)):
    cfg = dict(hidden_size=H, nhead=nh, num_decoder_layers=2, dim_feedforward=256, window_size=2,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", **cfg)
    m.load_state_dict(to.seeded_state_dict(cfg, 0), strict=False)
    m = m.cuda().eval()
    inp = {k: v.cuda() for k, v in to.model_inputs_from_batch(to.synthetic_batch(B, T + 1, 64)).items()}
    with torch.no_grad():
        c, p = m.sequential_inference(inp["frames"], inp["cad_image"], action=True)
    torch.cuda.synchronize()
    print("ok", H, B, float(c.abs().sum()), float(p.abs().sum()))
