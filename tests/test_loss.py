"""Fused training loss (vc_loss_forward / vc_loss_backward) against the torch restatement of the reference's
MultiClassesTrainer.compute_loss (videocad_b200/loss.py, itself pinned against the reference in tests/test_oracle.py)."""
import pytest
import torch

from videocad_b200 import lib as L
from videocad_b200.loss import compute_loss, compute_loss_fused


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
    if mode == "nan":
        tgt[0, 0, 3] = 1000.0                                          # out-of-range target: count == 0 -> NaN term, dropped
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


@pytest.mark.parametrize("mode", ["dense", "mixed", "empty", "nan"])
def test_fused_loss_cpu_restatement(mode):
    from oracle import build_emu

    lib = L.load(build_emu.build(), require_cuda_build=False)
    _check(*_case(32, 3, "cpu", mode), lib=lib)


def test_fused_loss_refuses_cpu_tensors_without_the_test_hook():
    c, p, t = _case(8, 1, "cpu")
    with pytest.raises(RuntimeError):
        compute_loss_fused((c, p), t)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dense", "mixed", "empty", "nan"])
def test_fused_loss_gpu(mode):
    _check(*_case(256, 5, "cuda", mode))


@pytest.mark.gpu
def test_fused_loss_gpu_is_deterministic():
    c, p, t = _case(256, 7, "cuda")
    a = compute_loss_fused((c, p), t)
    b = compute_loss_fused((c, p), t)
    assert torch.equal(a, b)
