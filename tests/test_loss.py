"""Fused training loss (vc_loss_forward / vc_loss_backward) against the torch restatement of the reference's
MultiClassesTrainer.compute_loss (videocad_b200/loss.py, itself pinned against the reference in tests/test_oracle.py)."""
import pytest
import torch

from videocad_b200 import lib as L
from videocad_b200.loss import (ABOVE, TOLERANCE, TOLERANCES, TOPK, compute_loss, compute_loss_and_metrics_fused, compute_loss_fused,
                                metrics_from_counts)


def _case(R, seed, device, mode="mixed"):
    g = torch.Generator().manual_seed(seed)
    B, T = R // 4, 4
    cmds = torch.randn(B, T, 5, generator=g) * 2
    params = torch.randn(B, T, 6, 1000, generator=g) * 3
    tgt = torch.empty(B, T, 7)
    tgt[..., 0] = torch.randint(0, 5, (B, T), generator=g).float()
    tgt[..., 1:] = torch.randint(0, 1000, (B, T, 6), generator=g).float()
    if mode != "dense":
        tgt[torch.rand(B, T, 7, generator=g) < 0.2] = -1.0          # ignored entries
        # some rows whose argmax already lies inside the window (they are dropped from the mean)
        for b in range(B):
            i = b % 6
            t = int(tgt[b, 0, 1 + i].item())
            if t >= 0:
                params[b, 0, i, min(t + 1, 999)] = 50.0
    if mode == "empty":
        tgt[..., 1] = -1.0                                             # parameter 0: nothing selected -> contributes 0
    if mode == "outrange":
        # out-of-range labels: the reference clamps both window ends into [0, 999] (trainer.py:880-905), so the window is
        # {999} for a target >= 1000 and starts at class 0 for a negative one (other than the ignore index -1)
        tgt[0, 0, 3] = 1000.0
        tgt[1, 1, 4] = 1700.0
        tgt[2, 2, 5] = -3.0      # tolerance 500: window [0, 496]
        tgt[3, 3, 1] = -7.0      # tolerance 2:   window {0}
    return cmds.to(device), params.to(device), tgt.to(device)


def _check(cmds, params, tgt, lib=None):
    c1, p1 = cmds.clone().requires_grad_(True), params.clone().requires_grad_(True)
    ref = compute_loss((c1, p1), tgt)
    (ref * 1.7).backward()
    c2, p2 = cmds.clone().requires_grad_(True), params.clone().requires_grad_(True)
    got = compute_loss_fused((c2, p2), tgt, _lib=lib)
    (got * 1.7).backward()
    assert torch.isfinite(got)
    assert abs(got.item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (got.item(), ref.item())
    for a, b, name in ((c2.grad, c1.grad, "dcmds"), (p2.grad, p1.grad, "dparams")):
        assert (a - b).abs().max().item() < 2e-6 + 2e-5 * b.abs().max().item(), name


def _metrics_torch(cmds, params, tgt):
    """restatement of the metrics half of MultiClassesTrainer.compute_loss (trainer.py:968-1061; pinned against the live reference
    in tests/test_oracle.py::test_loss_port_matches_reference_trainer)"""
    a = tgt.long()
    ac, ap = a[..., 0], a[..., 1:]
    pc, pp = cmds.argmax(-1), params.argmax(-1)
    cmd_mask = ac != -1
    cmd_ok = cmd_mask & (pc == ac)
    pmask = cmd_mask.unsqueeze(-1) & (ap != -1)
    diff = pp - ap
    ok = torch.stack([((diff[..., i] >= 0) & (diff[..., i] < TOLERANCES[i])) if ABOVE[i] else (diff[..., i].abs() < TOLERANCE) for i in range(6)], -1)
    pok = pmask & cmd_ok.unsqueeze(-1) & ok
    m = {"correct_predictions": int(cmd_ok.sum() + pok.sum()), "total_predictions": int(cmd_mask.sum() + pmask.sum()),
         "cmd_corrects": [int(((ac == k) & (pc == ac)).sum()) for k in range(5)], "cmd_counts": [int((ac == k).sum()) for k in range(5)],
         "param_corrects": [int(pok[..., i].sum()) for i in range(6)], "param_counts": [int(pmask[..., i].sum()) for i in range(6)],
         "cmd_correct_topk": int(cmd_ok[:, :TOPK].sum()), "cmd_counts_topk": int(cmd_mask[:, :TOPK].sum()),
         "param_correct_topk": int(pok[:, :TOPK].sum()), "param_counts_topk": int(pmask[:, :TOPK].sum())}
    return m


def _check_metrics(cmds, params, tgt, lib=None):
    dev = cmds.device
    cmds, params, tgt = cmds.cpu(), params.cpu(), tgt.cpu()  # the loops below index element by element
    with torch.no_grad():  # some right commands and right parameters, so that every counter moves
        a = tgt.long()
        for b in range(cmds.shape[0]):
            for t in range(cmds.shape[1]):
                if a[b, t, 0] >= 0 and (b + t) % 2:
                    cmds[b, t, a[b, t, 0]] += 20
                for i in range(6):
                    if a[b, t, 1 + i] >= 0 and (b + t + i) % 3 == 0:
                        params[b, t, i, min(int(a[b, t, 1 + i]) + (1 if ABOVE[i] else -1), 999)] += 40
    cmds, params, tgt = cmds.to(dev), params.to(dev), tgt.to(dev)
    loss, counts = compute_loss_and_metrics_fused((cmds, params), tgt, _lib=lib)
    ref_loss = compute_loss_fused((cmds, params), tgt, _lib=lib)
    assert torch.equal(loss, ref_loss)
    got, want = metrics_from_counts(counts), _metrics_torch(cmds.cpu(), params.cpu(), tgt.cpu())
    assert want["correct_predictions"] > 0 and sum(want["param_corrects"]) > 0
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)


@pytest.mark.parametrize("mode", ["dense", "mixed"])
def test_fused_metrics_cpu_restatement(mode):
    from oracle import build_emu

    lib = L.load(build_emu.build(), require_cuda_build=False)
    _check_metrics(*_case(32, 11, "cpu", mode), lib=lib)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dense", "mixed"])
def test_fused_metrics_gpu(mode):
    _check_metrics(*_case(256, 13, "cuda", mode))
    # sequences longer than the top-k window
    g = torch.Generator().manual_seed(2)
    B, T = 4, 40
    cm, pa = torch.randn(B, T, 5, generator=g).cuda(), torch.randn(B, T, 6, 1000, generator=g).cuda()
    tg = torch.cat([torch.randint(0, 5, (B, T, 1), generator=g), torch.randint(-1, 1000, (B, T, 6), generator=g)], -1).float().cuda()
    _check_metrics(cm, pa, tg)


@pytest.mark.parametrize("mode", ["dense", "mixed", "empty", "outrange"])
def test_fused_loss_cpu_restatement(mode):
    from oracle import build_emu

    lib = L.load(build_emu.build(), require_cuda_build=False)
    _check(*_case(32, 3, "cpu", mode), lib=lib)


def test_fused_loss_refuses_cpu_tensors_without_the_test_hook():
    c, p, t = _case(8, 1, "cpu")
    with pytest.raises(RuntimeError):
        compute_loss_fused((c, p), t)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dense", "mixed", "empty", "outrange"])
def test_fused_loss_gpu(mode):
    _check(*_case(256, 5, "cuda", mode))


@pytest.mark.gpu
def test_fused_loss_gpu_is_deterministic():
    c, p, t = _case(256, 7, "cuda")
    a = compute_loss_fused((c, p), t)
    b = compute_loss_fused((c, p), t)
    assert torch.equal(a, b)
