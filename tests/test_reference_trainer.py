"""The drop-in behind the UNMODIFIED reference trainer (/root/reference/trainer.py, or its staged copy oracle/_ref/).

`create_trainer(...)._process_batch(batch)` (trainer.py:480-496, 1384) is run twice on the same seeded batch and weights:
once around the reference's own model, once around `videocad_b200.AutoRegressiveTransformer` wrapped the way
experiment.py:92-93 wraps it (`torch.compile(dynamic=False)`); loss, clipped gradients and the weights after the Adam step
must agree.  On CPU the drop-in's kernels are the emulation library (host logic only); the `gpu` variant runs the CUDA library.
"""
import pytest
import torch

from oracle import reference_model as rm
from oracle import torch_oracle as to

pytestmark = pytest.mark.skipif(not rm.available(), reason="reference sources neither under /root/reference nor staged in oracle/_ref")

CFG = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, window_size=2,
           enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)


def _reference_step(batch, cfg=CFG):
    from oracle import ref_trainer as rt

    model = rt.build_reference_model(cfg, "cpu", seed=0)
    model.eval()  # _process_batch does not touch the mode; eval removes dropout so that the two runs are comparable
    trainer = rt.make_trainer(model, "cpu", lr=1e-3)
    loss, metrics = trainer._process_batch(batch)
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return loss.item(), metrics, grads, after


def _dropin_step(batch, device, emu=None, cfg=CFG, compile_it=True):
    from oracle import ref_trainer as rt
    from videocad_b200 import ModelFactory

    model, _ = ModelFactory().create_model("autoregressive", dict(cfg, state_dim=1644, act_dim=7, encoder="vit"), device,
                                           state_dict=to.seeded_state_dict(cfg, 0))
    if emu is not None:
        model._use_library_for_tests(emu)
    model.eval()
    wrapped = torch.compile(model, dynamic=False) if compile_it else model  # experiment.py:92-93
    trainer = rt.make_trainer(wrapped, device, lr=1e-3)
    loss, metrics = trainer._process_batch(batch)
    grads = {k: w.grad.detach().cpu().clone() for k, w in model.named_weights() if w.grad is not None}
    after = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    return loss.item(), metrics, grads, after


def _compare(ref, got):
    (l0, m0, g0, a0), (l1, m1, g1, a1) = ref, got
    assert abs(l0 - l1) < 2e-4 * max(1.0, abs(l0)), (l0, l1)
    assert m0 == m1, "the trainer's own metrics dict (argmax counters) must be identical"
    live = [k for k in g0 if k in g1]
    assert len(live) > 100
    for k in live:
        assert (g1[k] - g0[k]).abs().max() <= 2e-3 * g0[k].abs().max() + 1e-7, k
    # Adam(lr=1e-3), first step: every element moves by lr * sign-ish(g); elements whose gradient is at the rounding-noise level
    # may move the other way, everything else must land on the same value
    total = differing = 0
    for k in live:
        d = (a1[k] - a0[k]).abs()
        assert d.max().item() <= 2.1e-3, (k, d.max().item())  # at most 2 * lr apart
        total += d.numel()
        differing += int((d > 2e-5).sum())
    assert differing <= 2e-3 * total, (differing, total)


def test_unmodified_trainer_drives_the_dropin_cpu_emulation():
    from oracle import build_emu
    from videocad_b200 import lib as L

    emu = L.load(build_emu.build(), require_cuda_build=False)
    batch = to.synthetic_batch(2, 5, 64, seed=21)
    _compare(_reference_step(batch), _dropin_step(batch, "cpu", emu=emu))


@pytest.mark.gpu
def test_unmodified_trainer_drives_the_dropin_on_the_gpu():
    """Same comparison with the CUDA library, then three more optimiser steps of both trainers on the same batch: the drop-in's
    eager, CUDA-graph capture and replay paths all sit behind the unmodified `_process_batch`, and the two loss trajectories
    (each step sees the weights its own Adam produced) must stay together."""
    from oracle import ref_trainer as rt
    from videocad_b200 import ModelFactory

    batch = to.synthetic_batch(3, 7, 224, seed=22)
    _compare(_reference_step(batch), _dropin_step(batch, "cuda"))
    ref_model = rt.build_reference_model(CFG, "cpu", seed=0)
    ref_model.eval()
    ref_tr = rt.make_trainer(ref_model, "cpu", lr=1e-4)
    model, _ = ModelFactory().create_model("autoregressive", dict(CFG, state_dim=1644, act_dim=7, encoder="vit"), "cuda",
                                           state_dict=to.seeded_state_dict(CFG, 0))
    model.eval()
    got_tr = rt.make_trainer(torch.compile(model, dynamic=False), "cuda", lr=1e-4)
    for step in range(4):
        l0, m0 = ref_tr._process_batch(batch)
        l1, m1 = got_tr._process_batch(batch)
        assert abs(l0.item() - l1.item()) < 2e-3 * abs(l0.item()), (step, l0.item(), l1.item())
        assert m0["total_predictions"] == m1["total_predictions"]
    assert l0.item() < _reference_step(batch)[0]  # and the steps did train
