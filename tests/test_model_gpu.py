"""Parity of the full CUDA path (through the drop-in nn.Module and the C ABI) against the oracle and the committed
golden fixtures produced by the unmodified reference.  Run on the B200 box: python -m pytest tests -m gpu"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_oracle as to  # noqa: E402
from videocad_b200 import AutoRegressiveTransformer  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c0_shipped", "c0_states_actions", "states_only", "fullres_small", "multiview", "c1_model"]

# fp tolerance of the parity mode (3-pass split-bf16 GEMMs, fp32 everything else) against the reference's fp32
# forward: BASELINE.json asks for 1e-3 max-abs on the logits; measured ~2e-5, asserted at 2e-4.
LOGIT_TOL = 2e-4
GRAD_REL_TOL = 2e-3


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    norms = json.loads(bytes(z["grad_norms_json"]).decode())
    return z, meta, norms


def build(cfg, device="cuda", dropout=0.1, seed=0, **kw):
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, dropout=dropout, vit_dropout=dropout, encoder="vit", **cfg, **kw)
    sd = to.seeded_state_dict(cfg, seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(device), sd


def loss_weights(cs, ps, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(cs, generator=g), torch.randn(ps, generator=g) * 0.05


def cuda_inputs(B, T, S, seed=1234):
    batch = to.synthetic_batch(B, T + 1, S, seed=seed)
    inp = to.model_inputs_from_batch(batch)
    return {k: v.cuda() for k, v in inp.items()}, batch


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_vs_reference_golden(name):
    z, meta, norms = load_case(name)
    m, sd = build(meta["cfg"])
    m.eval()
    inp, _ = cuda_inputs(meta["B"], meta["T"], meta["S"], meta["batch_seed"])
    if meta["cfg"].get("num_views", 0) > 0:
        inp["multiview_images"] = to.synthetic_views(meta["B"], meta["cfg"]["num_views"], meta["S"], seed=4321).cuda()
    cmds, params = m(inp)
    dc = (cmds.cpu() - torch.from_numpy(z["cmds"])).abs().max().item()
    dp = (params.cpu() - torch.from_numpy(z["params"])).abs().max().item()
    assert dc < LOGIT_TOL and dp < LOGIT_TOL, f"{name}: max|d| cmds {dc:.3e} params {dp:.3e}"
    wc, wp = loss_weights(cmds.shape, params.shape, meta["loss_seed"])
    loss = (cmds * wc.cuda()).sum() + (params * wp.cuda()).sum()
    assert abs(loss.item() - float(z["loss"])) < 5e-3
    loss.backward()
    worst = 0.0
    for k, p in m.named_weights():
        if k not in norms:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, f"{k} should be unused"
            continue
        assert p.grad is not None, k
        n = p.grad.double().norm().item()
        rel = abs(n - norms[k]) / (norms[k] + 1e-6)
        assert rel < GRAD_REL_TOL, f"{name} {k}: grad norm {n:.6e} vs reference {norms[k]:.6e}"
        if ("grad::" + k) in z.files:
            ref = torch.from_numpy(z["grad::" + k])
            err = (p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
            worst = max(worst, err)
            assert err < GRAD_REL_TOL, f"{name} {k}: grad max rel err {err:.3e}"
    print(f"{name}: logits {dc:.2e}/{dp:.2e}, worst full-grad rel err {worst:.2e}")


def test_c1_shape_vs_oracle_fp64():
    """BASELINE config C1 model (H=512, 8 layers, window 10) at a reduced batch, against the fp64 oracle on the GPU."""
    cfg = dict(hidden_size=512, nhead=4, num_decoder_layers=8, dim_feedforward=512, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build(cfg)
    m.eval()
    inp, _ = cuda_inputs(3, 8, 224)
    with torch.no_grad():
        cmds, params = m(inp)
        sdd = {k: v.double().cuda() for k, v in sd.items()}
        oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    dc, dp = (cmds.double() - oc).abs().max().item(), (params.double() - op).abs().max().item()
    print(f"C1-shape logits max|d|: cmds {dc:.3e} params {dp:.3e}")
    assert dc < LOGIT_TOL and dp < LOGIT_TOL


def test_c2_full_shape_forward_backward_vs_oracle():
    """BASELINE config C2 at its full shape: model_configs/vid_pretrained.json "base_model" (H = 256, window 3, defaults nhead = 4,
    8 layers, Ff = 512) with the frame encoder enabled, batch 8, 16-frame context, 224 x 224 (128 frames: the 2-SM GEMM path with
    25 row blocks; decoder head dim 64, T = 16 short-sequence attention): logits and every gradient against the fp64 oracle."""
    cfg = dict(hidden_size=256, nhead=4, num_decoder_layers=8, dim_feedforward=512, window_size=3,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build(cfg, dropout=0.0)
    m.train()
    inp, _ = cuda_inputs(8, 16, 224)
    wc, wp = loss_weights((8, 16, 5), (8, 16, 6, 1000))
    for it in range(3):  # eager, graph capture, graph replay
        m.zero_grad(set_to_none=True)
        cmds, params = m(inp)
        ((cmds * wc.cuda()).sum() + (params * wp.cuda()).sum()).backward()
    sdd = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    ((oc * wc.cuda().double()).sum() + (op * wp.cuda().double()).sum()).backward()
    dc, dp = (cmds.double() - oc).abs().max().item(), (params.double() - op).abs().max().item()
    print(f"C2 logits max|d|: cmds {dc:.3e} params {dp:.3e}")
    assert dc < LOGIT_TOL and dp < LOGIT_TOL
    worst = 0.0
    for k, p in m.named_weights():
        ref = sdd[k].grad
        if ref is None:
            continue
        err = (p.grad.double() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        worst = max(worst, err)
        assert err < GRAD_REL_TOL, f"{k}: {err:.3e}"
    print(f"C2 worst relative gradient error {worst:.3e}")


def test_c3_width_small_batch_forward_backward_vs_oracle():
    """BASELINE config C3 model width (H = Ff = 1024, head dim 256) at a small batch: logits and a few gradients against
    the fp64 oracle (covers the d = 256 attention kernels, the 64-wide GEMM tiles and the auxiliary-stream fork/join)."""
    cfg = dict(hidden_size=1024, nhead=4, num_decoder_layers=3, dim_feedforward=1024, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build(cfg, dropout=0.0)
    m.train()  # dropout p = 0: training path without randomness
    inp, _ = cuda_inputs(2, 12, 64)
    wc, wp = loss_weights((2, 12, 5), (2, 12, 6, 1000))
    for it in range(3):  # eager, graph capture, graph replay
        m.zero_grad(set_to_none=True)
        cmds, params = m(inp)
        ((cmds * wc.cuda()).sum() + (params * wp.cuda()).sum()).backward()
    sdd = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    ((oc * wc.cuda().double()).sum() + (op * wp.cuda().double()).sum()).backward()
    assert (cmds.double() - oc).abs().max() < LOGIT_TOL and (params.double() - op).abs().max() < LOGIT_TOL
    for k, p in m.named_weights():
        ref = sdd[k].grad
        if ref is None:
            continue
        err = (p.grad.double() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        assert err < GRAD_REL_TOL, f"{k}: {err:.3e}"


def test_bf16_single_pass_mode_is_looser_but_close():
    cfg = dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=3,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build(cfg, precision="bf16")
    m.eval()
    inp, _ = cuda_inputs(2, 4, 224)
    with torch.no_grad():
        cmds, params = m(inp)
        oc, op = to.forward({k: v.cuda() for k, v in sd.items()}, cfg, inp)
    d = max((cmds - oc).abs().max().item(), (params - op).abs().max().item())
    assert 1e-4 < d < 0.1, d


# H = 256 (head dim 64), <= 16 sequences: the device-resident decode step (vc_seq_decode_step_dev; 9 sequences: its 16-row
# instantiation); 18 sequences: the per-kernel step's tensor-core path; H = 128 (head dim 32): the per-kernel step's row kernels
@pytest.mark.parametrize("B,T,H", [(2, 6, 256), (2, 19, 256), (9, 7, 256), (18, 5, 256), (2, 7, 128)])
def test_prefix_invariance_and_rollout(B, T, H):
    cfg = dict(hidden_size=H, nhead=4, num_decoder_layers=3, dim_feedforward=256, window_size=2,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build(cfg)
    m.eval()
    S = 64
    inp, _ = cuda_inputs(B, T, S)
    with torch.no_grad():
        full_c, full_p = m(inp)
        k = 4
        pre = {"frames": inp["frames"][:, :k], "actions": inp["actions"][:, :k], "cad_image": inp["cad_image"]}
        pc, pp = m(pre)
        assert (pc - full_c[:, :k]).abs().max() < 1e-5 and (pp - full_p[:, :k]).abs().max() < 1e-5
        # rollout without feedback == one forward with zero actions
        rc, rp = m.sequential_inference(inp["frames"], inp["cad_image"], action=False)
        zc, zp = m({"frames": inp["frames"], "actions": torch.zeros_like(inp["actions"]), "cad_image": inp["cad_image"]})
        assert (rc - zc).abs().max() < 1e-6 and (rp - zp).abs().max() < 1e-6
        # rollout with action feedback against the oracle's O(T^2) recompute
        oc, op = to.rollout(sd, cfg, inp["frames"].cpu(), inp["cad_image"].cpu(), action=True)
        for rep in range(3):  # eager, graph capture, graph replay
            ac, ap = m.sequential_inference(inp["frames"], inp["cad_image"], action=True)
            assert (ac.cpu().argmax(-1) == oc.argmax(-1)).all()
            assert (ac.cpu() - oc).abs().max() < LOGIT_TOL and (ap.cpu() - op).abs().max() < LOGIT_TOL, rep


def test_training_mode_dropout_is_reproducible_and_trains():
    cfg = dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=3,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    from videocad_b200.loss import compute_loss

    m, sd = build(cfg, dropout=0.1)
    m.train()
    inp, batch = cuda_inputs(4, 4, 64)
    tgt = batch["actions"][:, 1:].cuda()
    torch.manual_seed(7)
    c1, p1 = m(inp)
    torch.manual_seed(7)
    c2, p2 = m(inp)
    assert torch.equal(c1, c2) and torch.equal(p1, p2), "same torch seed -> same dropout masks"
    c3, _ = m(inp)
    assert not torch.equal(c1, c3), "a new seed is drawn on every forward"
    m.eval()
    ce, _ = m(inp)
    assert 1e-3 < (ce - c1).abs().max().item() < 5.0
    # a few Adam steps on one batch must reduce the reference loss
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = compute_loss(m(inp), tgt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert np.mean(losses[-3:]) < np.mean(losses[:3]), losses


def test_cuda_graph_replay_matches_eager():
    """2nd use of a shape captures the native call sequence into a CUDA graph, later uses replay it: outputs and
    gradients must be identical to the eager launches, in eval mode and (same seed) in training mode."""
    cfg = dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=3,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, _ = build(cfg, dropout=0.1)
    inp, _ = cuda_inputs(3, 4, 64)
    wc, wp = loss_weights((3, 4, 5), (3, 4, 6, 1000))
    for mode in ("eval", "train"):
        m.train(mode == "train")
        m._draw_seed = lambda: 987654321
        outs, grads = [], []
        for it in range(4):  # eager, capture, replay, replay
            m.zero_grad(set_to_none=True)
            c, p = m(inp)
            ((c * wc.cuda()).sum() + (p * wp.cuda()).sum()).backward()
            outs.append((c.detach().clone(), p.detach().clone()))
            grads.append({k: v.grad.clone() for k, v in m.named_weights() if v.grad is not None})
        for it in range(1, 4):
            assert torch.equal(outs[it][0], outs[0][0]) and torch.equal(outs[it][1], outs[0][1]), (mode, it)
            for k in grads[0]:
                d = (grads[it][k] - grads[0][k]).abs().max().item()
                assert d <= 1e-4 * grads[0][k].abs().max().item() + 1e-9, (mode, it, k, d)  # split-K atomics reorder sums
    # weights changed by an optimizer step must be picked up by the replayed graph
    m.eval()
    with torch.no_grad():
        before, _ = m(inp)
        W = dict(m.named_weights())
        W["embed_action.weight"].data.add_(0.05)
        W["cad_embedding_model.transformer.layers.0.1.net.1.weight"].data.mul_(1.1)
        after, _ = m(inp)
    assert (after - before).abs().max() > 1e-4


def test_uint8_frame_ingestion_equals_normalised_float_frames():
    """uint8 grey-level frames / CAD image handed to the module are normalised on the device exactly as
    the reference's loader does on the CPU (ToTensor + Normalize([0.5], [0.5]), main.py:103-110): identical logits and gradients,
    on the eager, captured and replayed paths."""
    cfg = dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=3,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, _ = build(cfg, dropout=0.0)
    m.eval()
    inp, _ = cuda_inputs(3, 4, 64)
    g = torch.Generator().manual_seed(11)
    f8 = torch.randint(0, 256, tuple(inp["frames"].shape), dtype=torch.uint8, generator=g).cuda()
    c8 = torch.randint(0, 256, tuple(inp["cad_image"].shape), dtype=torch.uint8, generator=g).cuda()
    norm = lambda u: u.cpu().to(torch.float32).div(255).sub_(0.5).div_(0.5).cuda()  # the loader's CPU arithmetic (true division)
    wc, wp = loss_weights((3, 4, 5), (3, 4, 6, 1000))
    res = []
    for frames, cad in ((f8, c8), (norm(f8), norm(c8))):
        for it in range(3):
            m.zero_grad(set_to_none=True)
            c, p = m({"frames": frames, "actions": inp["actions"], "cad_image": cad})
            ((c * wc.cuda()).sum() + (p * wp.cuda()).sum()).backward()
            res.append((c.detach().clone(), p.detach().clone(),
                        dict(m.named_weights())["state_embedding_model.to_patch_embedding.2.weight"].grad.clone()))
    for r in res[1:]:
        assert torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1])
        assert (r[2] - res[0][2]).abs().max() <= 1e-4 * res[0][2].abs().max() + 1e-9  # split-K atomics reorder sums


def test_full_size_c1_batch_properties():
    """BASELINE C1 at full size (B=32, T=8, 224x224, H=512): size-independent properties."""
    cfg = dict(hidden_size=512, nhead=4, num_decoder_layers=8, dim_feedforward=512, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, _ = build(cfg)
    m.eval()
    inp, _ = cuda_inputs(32, 8, 224)
    with torch.no_grad():
        c, p = m(inp)
        assert c.shape == (32, 8, 5) and p.shape == (32, 8, 6, 1000)
        assert torch.isfinite(c).all() and torch.isfinite(p).all()
        # batch independence: any sub-batch gives the same rows
        sub = {k: v[5:9] for k, v in inp.items()}
        sc, sp = m(sub)
        assert (sc - c[5:9]).abs().max() < 1e-5 and (sp - p[5:9]).abs().max() < 1e-5
        # prefix invariance at full size
        pre = {"frames": inp["frames"][:, :3], "actions": inp["actions"][:, :3], "cad_image": inp["cad_image"]}
        pc, pp = m(pre)
        assert (pc - c[:, :3]).abs().max() < 1e-5 and (pp - p[:, :3]).abs().max() < 1e-5


def test_errors_match_reference_behaviour():
    with pytest.raises(ValueError):
        AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=256, encoder="resnet")
    with pytest.raises(AssertionError):
        AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=256, window_size=0)
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=256, enable_past_actions=True, enable_past_states=True).cuda()
    bad = {"frames": torch.zeros(1, 2, 1, 256, 256).cuda(), "actions": torch.zeros(1, 2, 7).cuda(), "cad_image": torch.zeros(1, 1, 64, 64).cuda()}
    with pytest.raises(ValueError):
        m(bad)  # > 224x224 exceeds the positional table (the reference raises a shape error here too)


def _grad_error(g, ref):
    """(max, 99.5th percentile) of |g - ref| / max|ref| over the elements of one weight's gradient."""
    d = ((g.double() - ref).abs() / (ref.abs().max() + 1e-6)).flatten()
    k = max(1, int(0.995 * d.numel()))
    return d.max().item(), d.kthvalue(k).values.item()


def test_c3_per_gpu_shape_forward_backward_vs_oracle():
    """BASELINE config C3 at its per-GPU model shape and a reduced batch: H = Ff = 1024 (head dim 256), 8 decoder layers, window 10,
    T = 32, 224 x 224 frames, batch 4 (128 frames + 4 CAD images; T = 32 short-sequence attention with d = 256, the F = 128-image
    encoder GEMMs, 8-layer workspace carve): logits and EVERY gradient against the fp64 oracle run on the GPU, on the eager,
    CUDA-graph capture and replay paths.

    Gradient tolerance at this size.  The decoder's FFN has 8 x 128 x 1024 ~ 1M ReLU inputs; a pre-activation that the fp64
    oracle puts within the forward rounding error of zero (ours: ~1e-6 absolute there, logits within 3e-5) may get the other sign,
    and relu' is discontinuous: the corresponding row of linear1.weight / entry of linear1.bias is then off by a whole term
    (measured: one such element in layers 1 and 4, 3e-2 / 9e-2 of the largest entry, profiles/r02b_diag_c3.txt; plain torch fp32
    flips less often only because its forward error is smaller), and every gradient upstream of that layer moves by ~1e-3
    (largest measured: 2.0e-3).  So: every gradient within 5e-3 of the oracle relative to its largest entry -- a wrong kernel
    is off by O(1) -- except the linear1 tensors, where 99.5 % of the elements must meet the same bound and the few rows / entries
    behind a flipped relu' at most 0.2."""
    cfg = dict(hidden_size=1024, nhead=4, num_decoder_layers=8, dim_feedforward=1024, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    B, T, S = 4, 32, 224
    m, sd = build(cfg, dropout=0.0)
    m.train()  # dropout p = 0: the training path (workspaces, gradient arenas) without randomness
    inp, _ = cuda_inputs(B, T, S)
    wc, wp = loss_weights((B, T, 5), (B, T, 6, 1000))
    sdd = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    ((oc * wc.cuda().double()).sum() + (op * wp.cuda().double()).sum()).backward()
    ref = {k: v.grad for k, v in sdd.items()}
    oc, op = oc.detach(), op.detach()
    del sdd
    for it in range(3):  # eager, graph capture, graph replay
        m.zero_grad(set_to_none=True)
        cmds, params = m(inp)
        ((cmds * wc.cuda()).sum() + (params * wp.cuda()).sum()).backward()
        dc, dp = (cmds.double() - oc).abs().max().item(), (params.double() - op).abs().max().item()
        assert dc < LOGIT_TOL and dp < LOGIT_TOL, (it, dc, dp)
        worst, worst_q, relu_rows = 0.0, 0.0, 0
        for k, p in m.named_weights():
            if ref[k] is None:
                continue
            emax, eq = _grad_error(p.grad, ref[k])
            worst, worst_q = max(worst, emax), max(worst_q, eq)
            if ".linear1." in k:  # the direct victims of a flipped relu': bounded, and confined to a few rows
                assert emax < 0.2 and eq < 5e-3, f"pass {it} {k}: max {emax:.3e}, 99.5th percentile {eq:.3e}"
                relu_rows += int(emax > 5e-3)
            else:
                assert emax < 5e-3, f"pass {it} {k}: {emax:.3e}"
        assert relu_rows <= 8, "relu' flips in at most a few layers"
    print(f"C3 shape: logits max|d| cmds {dc:.3e} params {dp:.3e}; gradient error max {worst:.3e}, 99.5th percentile {worst_q:.3e}")


def _fed_back_actions(c, p):
    """[B,T,5], [B,T,6,1000] logits -> the raw action rows the rollout feeds back ([B,T,7]; masked parameters are -1)."""
    cp, pp = c.argmax(-1), p.argmax(-1)
    return torch.cat([cp.unsqueeze(-1), to.apply_action_mask(cp, pp)], dim=-1)


def test_c4_rollout_shape_incremental_decode_vs_oracle():
    """BASELINE config C4 model and horizon: H = Ff = 1024, 8 layers, window 10, T = 186 steps of action feedback, 2 sequences
    (64 x 64 frames keep the oracle's O(T^2) recompute affordable): the incremental key/value-cached decode must reproduce the
    oracle (fp64 recompute of the whole prefix at every step): same fed-back action sequence, logits within the tolerance, on the
    eager, capture and replay paths.  The two rollouts may only part where the oracle's own argmax is decided by less than the
    logit tolerance (a rounding-level tie); everything up to such a step is compared, and it must be a tie."""
    cfg = dict(hidden_size=1024, nhead=4, num_decoder_layers=8, dim_feedforward=1024, window_size=10,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    B, T, S = 2, 186, 64
    m, sd = build(cfg)
    m.eval()
    inp, _ = cuda_inputs(B, T, S)
    with torch.no_grad():
        sdd = {k: v.double().cuda() for k, v in sd.items()}
        oc, op = to.rollout(sdd, cfg, inp["frames"].double(), inp["cad_image"].double(), action=True)
        oc, op = oc.float().cpu(), op.float().cpu()
        want = _fed_back_actions(oc, op)
        for rep in range(3):  # eager, graph capture, graph replay
            ac, ap = m.sequential_inference(inp["frames"], inp["cad_image"], action=True)
            ac, ap = ac.cpu(), ap.cpu()
            assert ac.shape == (B, T, 5) and ap.shape == (B, T, 6, 1000)
            got = _fed_back_actions(ac, ap)
            differs = (got != want).any(-1).any(0).nonzero()
            n = int(differs[0]) if differs.numel() else T  # steps [0, n) fed identical actions forward; step n saw identical inputs
            if n < T:
                gaps = torch.cat([oc[:, n].topk(2, -1).values, op[:, n].reshape(B * 6, -1).topk(2, -1).values])
                assert (gaps[:, 0] - gaps[:, 1]).min() < LOGIT_TOL, f"rollouts part at step {n} without a rounding-level tie"
            assert n >= 8, f"first tie already at step {n}: pick another seed"
            k = min(n + 1, T)
            dc, dp = (ac[:, :k] - oc[:, :k]).abs().max().item(), (ap[:, :k] - op[:, :k]).abs().max().item()
            assert dc < LOGIT_TOL and dp < LOGIT_TOL, (rep, dc, dp)
        print(f"C4 rollout: identical fed-back actions for {n} of {T} steps; logits max|d| {dc:.2e} / {dp:.2e} over {k} steps")
        # teacher-forced cross-check over ALL 186 positions: one full-length forward with the oracle's own fed-back actions
        acts = torch.zeros(B, T, 7)
        cp, pp = oc.argmax(-1), op.argmax(-1)
        nxt = to.apply_action_mask(cp, pp).float()
        acts[:, 1:] = to.normalize_actions(torch.cat([cp.unsqueeze(-1).float(), nxt], dim=-1))[:, :-1]
        fc, fp = m({"frames": inp["frames"], "actions": acts.cuda(), "cad_image": inp["cad_image"]})
        assert (fc.cpu() - oc).abs().max() < LOGIT_TOL and (fp.cpu() - op).abs().max() < LOGIT_TOL
