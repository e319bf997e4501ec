"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE; run in the build container only).

For every case: seeded weights (oracle.torch_oracle.seeded_state_dict -- construction-order independent) are loaded
into the real reference model imported from /root/reference (with the vit_pytorch/timm shims of oracle/shims), the
seeded synthetic batch of SURVEY.md 8(d) is pushed through `model.eval()(inputs)` and through autograd for the
scalar loss  sum(cmds * wc) + sum(params * wp);  outputs, a few full gradient tensors and all gradient norms are
stored.  The fixtures pin oracle/torch_oracle.py (CPU tests) and the CUDA path (GPU tests) to the reference.

    python -m oracle.make_golden
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import torch_oracle as to  # noqa: E402
from oracle.reference_model import build_reference_model  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

BASE = dict(act_dim=7, state_dim=1644, max_length=None, num_classes=5, encoder="vit", normalize=True, num_views=0,
            model_name="autoregressive")

CASES = {
    # model_configs/autoregressive_transformer.json["default_params"] with encoder -> vit: frames unused
    "c0_shipped": dict(cfg=dict(hidden_size=256, nhead=4, num_decoder_layers=8, dim_feedforward=256, enable_past_actions=False),
                       B=2, T=2, S=64),
    # C0 variant that exercises the frame encoder (past states + actions + timestep embedding)
    "c0_states_actions": dict(cfg=dict(hidden_size=256, nhead=4, num_decoder_layers=8, dim_feedforward=256, window_size=3,
                                       enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
                              B=2, T=2, S=64),
    # past-states-only branch (banded self-attention on the frame tokens)
    "states_only": dict(cfg=dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=2,
                                 enable_past_actions=False, enable_past_states=True, enable_timestep_embedding=True),
                        B=2, T=3, S=64),
    # full-resolution frames (50 tokens per image), transformer_experiments.json-style model at reduced width/depth
    "fullres_small": dict(cfg=dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=512, window_size=10,
                                   enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
                          B=1, T=3, S=224),
    # BASELINE configs[1] model (C1: transformer_experiments.json "cad_past_10_actions_and_states_timestep_embedding" with
    # hidden_size = dim_feedforward = 512, 8 layers, window 10) at full resolution, reduced batch
    "c1_model": dict(cfg=dict(hidden_size=512, nhead=4, num_decoder_layers=8, dim_feedforward=512, window_size=10,
                              enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
                     B=1, T=3, S=224),
    # multiview conditioning: two extra views through the CAD encoder, embed_multiview, 3-source image_projection
    "multiview": dict(cfg=dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, window_size=2, num_views=2,
                               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
                      B=2, T=3, S=64),
}

FULL_GRADS = [
    "cad_embedding_model.cls_token", "cad_embedding_model.transformer.norm.weight",
    "cad_embedding_model.to_patch_embedding.1.bias", "cad_embedding_model.transformer.layers.0.0.to_out.0.bias",
    "state_embedding_model.transformer.layers.5.1.net.1.bias", "state_embedding_model.to_patch_embedding.3.weight",
    "embed_action.weight", "embed_state.bias", "embed_image.bias", "image_projection.bias",
    "transformer_decoder.layers.0.self_attn.in_proj_bias", "transformer_decoder.layers.1.multihead_attn.in_proj_bias",
    "transformer_decoder.layers.1.norm2.weight", "transformer_decoder.layers.0.linear1.bias",
    "predict_action_class_0_4.weight", "predict_action_class_0_4.bias", "embed_multiview.bias",
]


def loss_weights(cmds_shape, params_shape, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(cmds_shape, generator=g), torch.randn(params_shape, generator=g) * 0.05


def run_case(name, case):
    cfg = dict(BASE, **case["cfg"])
    model, _ = build_reference_model(cfg)
    sd = to.seeded_state_dict(cfg, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    live_missing = [k for k in missing if not (k.startswith("transformer.") or k.startswith("embed_timestep") or
                                               k.startswith("embed_ln") or k.startswith("predict_action."))]
    assert not live_missing, live_missing
    model.eval()
    batch = to.synthetic_batch(case["B"], case["T"] + 1, case["S"], seed=1234)
    inp = to.model_inputs_from_batch(batch)
    inp["timesteps"] = torch.zeros(case["B"], 1, dtype=torch.long)
    nv = cfg.get("num_views", 0)
    if nv > 0:
        inp["multiview_images"] = to.synthetic_views(case["B"], nv, case["S"], seed=4321)
    cmds, params = model(inp)
    wc, wp = loss_weights(cmds.shape, params.shape)
    loss = (cmds * wc).sum() + (params * wp).sum()
    loss.backward()
    out = {"cmds": cmds.detach().numpy(), "params": params.detach().numpy(), "loss": np.float64(loss.item())}
    norms = {}
    for k, p in model.named_parameters():
        if k in sd and p.grad is not None:
            norms[k] = float(p.grad.double().norm().item())
            if k in FULL_GRADS:
                out["grad::" + k] = p.grad.detach().numpy()
    out["grad_norms_json"] = np.frombuffer(json.dumps(norms).encode(), dtype=np.uint8)
    out["weight_checksum"] = np.float64(sum(v.double().abs().sum().item() for v in sd.values()))
    out["meta_json"] = np.frombuffer(json.dumps(dict(cfg=case["cfg"], B=case["B"], T=case["T"], S=case["S"], weight_seed=0,
                                                     batch_seed=1234, loss_seed=5, torch=torch.__version__)).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "cmds", tuple(cmds.shape), "loss %.6f" % loss.item(), "grads", len(norms))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1:]  # optional case names: generate only those (the committed fixtures of the others stay untouched)
    for n, c in CASES.items():
        if not only or n in only:
            run_case(n, c)
