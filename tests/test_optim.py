"""Fused clip_grad_norm_ + Adam (vc_clip_adam_step) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam, the two
calls of the reference trainer (trainer.py:493-494)."""
import pytest
import torch

from videocad_b200 import lib as L
from videocad_b200.optim import ClipAdam


def _run(device, lib=None, steps=4, lrs=(1e-3, 5e-4, 1e-3), big_grads=True):
    g = torch.Generator().manual_seed(0)
    sizes = (4096 + 64, 1000003, 77)
    ref_p = [torch.randn(n, generator=g).to(device).requires_grad_(True) for n in sizes]
    got_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref_opt = torch.optim.Adam([dict(params=[p], lr=lr) for p, lr in zip(ref_p, lrs)])
    got_opt = ClipAdam([dict(params=[p], lr=lr) for p, lr in zip(got_p, lrs)], max_norm=1.0, _lib=lib)
    for it in range(steps):
        scale = (10.0 if big_grads else 1e-4) * (it + 1)
        grads = [(torch.randn(n, generator=g) * scale).to(device) for n in sizes]
        for p, q, gr in zip(ref_p, got_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        ref_opt.step()
        got_opt.step()
        assert abs(got_opt.last_grad_norm.item() - norm.item()) <= 5e-5 * norm.item()  # torch accumulates 1e6 squares in fp32
        for p, q in zip(ref_p, got_p):
            assert (q.grad - p.grad).abs().max() <= 1e-4 * p.grad.abs().max() + 1e-12      # clipped in place, like torch
            assert (q - p).abs().max().item() <= 6e-7, it  # 1 ulp at |p| ~ 4;                                # updates are O(lr) = 1e-3
            sp, sq = ref_opt.state[p], got_opt.state[q]
            assert (sq["exp_avg"] - sp["exp_avg"]).abs().max() <= 1e-4 * sp["exp_avg"].abs().max() + 1e-12
            assert (sq["exp_avg_sq"] - sp["exp_avg_sq"]).abs().max() <= 2e-4 * sp["exp_avg_sq"].abs().max() + 1e-20
    # optimizer state round-trips with torch.optim.Adam's format
    sd = got_opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and sd["state"][0]["step"].item() == steps


@pytest.mark.parametrize("big_grads", [True, False])
def test_clip_adam_cpu_restatement(big_grads):
    from oracle import build_emu

    lib = L.load(build_emu.build(), require_cuda_build=False)
    _run("cpu", lib=lib, big_grads=big_grads)


def test_clip_adam_refuses_cpu_tensors_without_the_test_hook():
    p = torch.zeros(8, requires_grad=True)
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError):
        ClipAdam([p]).step()


@pytest.mark.gpu
@pytest.mark.parametrize("big_grads", [True, False])
def test_clip_adam_gpu(big_grads):
    _run("cuda", big_grads=big_grads)
