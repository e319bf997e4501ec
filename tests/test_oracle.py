"""CPU tests of the oracle: pinned against the golden fixtures (outputs of the unmodified reference) and, where
/root/reference exists (the build container), against the reference itself."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as to
from oracle import reference_model as rm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c0_shipped", "c0_states_actions", "states_only", "fullres_small", "multiview", "c1_model"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    norms = json.loads(bytes(z["grad_norms_json"]).decode())
    return z, meta, norms


def loss_weights(cs, ps, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(cs, generator=g), torch.randn(ps, generator=g) * 0.05


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    z, meta, norms = load_case(name)
    cfg = meta["cfg"]
    sd = {k: v.requires_grad_(True) for k, v in to.seeded_state_dict(cfg, meta["weight_seed"]).items()}
    chk = sum(v.detach().double().abs().sum().item() for v in sd.values())
    assert abs(chk - float(z["weight_checksum"])) < 1e-6 * chk, "seeded weights are not reproducible on this machine"
    batch = to.synthetic_batch(meta["B"], meta["T"] + 1, meta["S"], seed=meta["batch_seed"])
    inp = to.model_inputs_from_batch(batch)
    if cfg.get("num_views", 0) > 0:
        inp["multiview_images"] = to.synthetic_views(meta["B"], cfg["num_views"], meta["S"], seed=4321)
    cmds, params = to.forward(sd, cfg, inp)
    assert (cmds.detach() - torch.from_numpy(z["cmds"])).abs().max() < 2e-5
    assert (params.detach() - torch.from_numpy(z["params"])).abs().max() < 2e-5
    wc, wp = loss_weights(cmds.shape, params.shape, meta["loss_seed"])
    ((cmds * wc).sum() + (params * wp).sum()).backward()
    for k, ref_norm in norms.items():
        g = sd[k].grad
        assert g is not None, k
        assert abs(g.double().norm().item() - ref_norm) < 1e-3 * ref_norm + 1e-7, k
        if ("grad::" + k) in z.files:
            ref = torch.from_numpy(z["grad::" + k])
            assert (g - ref).abs().max() < 2e-4 * ref.abs().max() + 1e-7, k
    unused = [k for k in sd if k not in norms]
    for k in unused:
        assert sd[k].grad is None or sd[k].grad.abs().max() == 0


@pytest.mark.skipif(not rm.available(), reason="/root/reference not present on this machine")
@pytest.mark.parametrize("mode", [(True, True, 0), (False, True, 0), (False, False, 0), (True, False, 0), (True, True, 2), (False, False, 3)])
def test_oracle_vs_live_reference(mode):
    pa, ps, nv = mode
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, window_size=2, encoder="vit", num_views=nv,
               enable_past_actions=pa, enable_past_states=ps, enable_timestep_embedding=pa or ps)
    model, _ = rm.build_reference_model(cfg)
    sd = to.seeded_state_dict(cfg, 3)
    model.load_state_dict(sd, strict=False)
    model.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(2, 5, 64, seed=9))
    if nv > 0:
        inp["multiview_images"] = to.synthetic_views(2, nv, 64)
    with torch.no_grad():
        rc, rp = model(dict(inp, timesteps=torch.zeros(2, 1)))
        oc, op = to.forward(sd, cfg, inp)
    assert (rc - oc).abs().max() < 5e-6 and (rp - op).abs().max() < 5e-6


def test_state_dict_schema_matches_reference_or_appendix_b():
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, enable_past_actions=True,
               enable_past_states=True, enable_timestep_embedding=True)
    shapes = to.param_shapes(cfg)
    if rm.available():
        model, _ = rm.build_reference_model(dict(cfg, encoder="vit"))
        ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        dead = ("transformer.", "embed_timestep", "embed_ln", "predict_action.")
        live = {k: v for k, v in ref.items() if not k.startswith(dead)}
        assert live == shapes
    assert shapes["transformer_decoder.layers.1.multihead_attn.in_proj_weight"] == (384, 128)
    assert shapes["cad_embedding_model.transformer.layers.5.0.to_qkv.weight"] == (3072, 512)


def test_prefix_invariance_and_rollout_semantics():
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, window_size=2,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    sd = to.seeded_state_dict(cfg, 1)
    inp = to.model_inputs_from_batch(to.synthetic_batch(2, 6, 64, seed=4))
    with torch.no_grad():
        c, p = to.forward(sd, cfg, inp)
        k = 3
        pc, pp = to.forward(sd, cfg, {"frames": inp["frames"][:, :k], "actions": inp["actions"][:, :k], "cad_image": inp["cad_image"]})
        assert (pc - c[:, :k]).abs().max() < 1e-5 and (pp - p[:, :k]).abs().max() < 1e-5
        rc, rp = to.rollout(sd, cfg, inp["frames"], inp["cad_image"], action=False)
        zc, zp = to.forward(sd, cfg, dict(inp, actions=torch.zeros_like(inp["actions"])))
        assert (rc - zc).abs().max() < 1e-5 and (rp - zp).abs().max() < 1e-5


def test_action_mask_rules():
    cmd = torch.tensor([[0, 1, 1, 2, 3, 4]])
    par = torch.tensor([[[5, 6, 7, 8, 9, 10], [5, 6, 210, 8, 9, 10], [5, 6, 300, 8, 9, 10], [5, 6, 7, 8, 9, 10],
                         [5, 6, 7, 8, 9, 10], [5, 6, 7, 8, 9, 10]]])
    out = to.apply_action_mask(cmd, par)
    assert out[0, 0].tolist() == [5, 6, -1, -1, -1, -1]
    assert out[0, 1].tolist() == [-1, -1, 210, 8, -1, -1]      # param[3] kept only if 200 <= param[2] < 250
    assert out[0, 2].tolist() == [-1, -1, 300, -1, -1, -1]
    assert out[0, 3].tolist() == [-1, -1, -1, -1, 9, -1]
    assert out[0, 4].tolist() == [-1, -1, -1, -1, -1, 10]
    assert out[0, 5].tolist() == [-1] * 6


@pytest.mark.skipif(not rm.available(), reason="/root/reference not present on this machine")
def test_loss_port_matches_reference_trainer(tmp_path, monkeypatch):
    import shutil

    from videocad_b200.loss import compute_loss

    shutil.copy(os.path.join(rm.REFERENCE_ROOT, "class_weights.json"), tmp_path / "class_weights.json")
    monkeypatch.chdir(tmp_path)
    tr = rm.import_trainer()

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

    pk = {"loader": [], "sampler": None}
    t = tr.MultiClassesTrainer(pk, pk, pk, Dummy(), {"lr": 1e-5, "use_mse": True, "experiment_name": "x"}, "cpu", 0)
    torch.manual_seed(0)
    B, T = 3, 7
    tgt = to.synthetic_batch(B, T + 1, 32)["actions"][:, 1:]
    cmds = torch.randn(B, T, 5, requires_grad=True)
    params = torch.randn(B, T, 6, 1000, requires_grad=True)
    with torch.no_grad():  # make some predictions land inside the tolerance window (those rows are dropped)
        for b in range(B):
            for tt in range(0, T, 2):
                for i in range(6):
                    v = int(tgt[b, tt, 1 + i].item())
                    if v >= 0:
                        params[b, tt, i, min(v + 1, 999)] += 20
    with torch.no_grad():  # and make some commands right, so that the parameter counters (which need a right command) move
        for b in range(B):
            for tt in range(1, T, 2):
                c = int(tgt[b, tt, 0].item())
                if c >= 0:
                    cmds[b, tt, c] += 10
    # out-of-range labels: both window ends are clamped into [0, 999] by the reference (trainer.py:880-905)
    from videocad_b200.loss import flexible_cross_entropy as fce
    g3 = torch.Generator().manual_seed(12)
    lg = torch.randn(8, 1000, generator=g3)
    for tol in (2, 50, 500):
        tg3 = torch.tensor([1000, 1700, -3, -7, 999, 0, 998, -1])
        want = t.flexible_cross_entropy(lg, tg3, 1000, tolerance=tol, above=[True])
        assert abs(fce(lg, tg3, tol).item() - want.item()) < 1e-5 * abs(want.item()), tol
    lr, metrics_ref = t.compute_loss((cmds, params), tgt)
    lm = compute_loss((cmds, params), tgt)
    assert abs(lr.item() - lm.item()) < 1e-5 * abs(lr.item())
    # the metrics half (trainer.py:968-1061) through the native counters (CPU twin of vc_loss_metrics), T > top-k window too
    from oracle import build_emu
    from videocad_b200 import lib as L
    from videocad_b200.loss import compute_loss_and_metrics_fused, metrics_from_counts

    emu = L.load(build_emu.build(), require_cuda_build=False)
    _, counts = compute_loss_and_metrics_fused((cmds.detach(), params.detach()), tgt, _lib=emu)
    got = metrics_from_counts(counts)
    assert got["total_predictions"] > 0 and got["correct_predictions"] > 0 and sum(got["param_corrects"]) > 0
    for k, v in metrics_ref.items():
        assert got[k] == v, (k, got[k], v)
    B2, T2 = 2, 34  # longer than k = 30: the *_topk counters differ from the totals
    tgt2 = to.synthetic_batch(B2, T2 + 1, 32, seed=9)["actions"][:, 1:]
    g2 = torch.Generator().manual_seed(4)
    cm2, pa2 = torch.randn(B2, T2, 5, generator=g2), torch.randn(B2, T2, 6, 1000, generator=g2)
    for b in range(B2):
        for tt in range(T2):
            c = int(tgt2[b, tt, 0].item())
            if c >= 0 and tt % 3:
                cm2[b, tt, c] += 10
            for i in range(6):
                v = int(tgt2[b, tt, 1 + i].item())
                if v >= 0 and (tt + i) % 2:
                    pa2[b, tt, i, min(v + 1, 999)] += 20
    _, m2_ref = t.compute_loss((cm2, pa2), tgt2)
    _, c2 = compute_loss_and_metrics_fused((cm2, pa2), tgt2, _lib=emu)
    got2 = metrics_from_counts(c2)
    assert got2["cmd_counts_topk"] < got2["total_predictions"] and got2["param_correct_topk"] > 0
    for k, v in m2_ref.items():
        assert got2[k] == v, (k, got2[k], v)
    gr = torch.autograd.grad(lr, [cmds, params])
    gm = torch.autograd.grad(lm, [cmds, params])
    assert (gr[0] - gm[0]).abs().max() < 1e-6 and (gr[1] - gm[1]).abs().max() < 1e-6
