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


def _dropin_step(batch, device, emu=None, cfg=CFG, compile_it=True, accelerate=False):
    from oracle import ref_trainer as rt
    from videocad_b200 import ModelFactory

    model, _ = ModelFactory().create_model("autoregressive", dict(cfg, state_dim=1644, act_dim=7, encoder="vit"), device,
                                           state_dict=to.seeded_state_dict(cfg, 0))
    if emu is not None:
        model._use_library_for_tests(emu)
    model.eval()
    wrapped = torch.compile(model, dynamic=False) if compile_it else model  # experiment.py:92-93
    trainer = rt.make_trainer(wrapped, device, lr=1e-3)
    if accelerate:
        from videocad_b200.trainer_accel import accelerate_trainer

        assert accelerate_trainer(trainer, _lib=emu) is trainer
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


def test_accelerated_trainer_cpu_emulation():
    """`accelerate_trainer` on the unmodified trainer object: fused loss + lazily read metrics + ClipAdam behind the same
    `_process_batch` call give the reference trainer's loss, metrics dict, clipped gradients and Adam-updated weights."""
    from oracle import build_emu
    from oracle import ref_trainer as rt
    from videocad_b200 import lib as L
    from videocad_b200.optim import ClipAdam
    from videocad_b200.trainer_accel import LazyMetrics, accelerate_trainer

    emu = L.load(build_emu.build(), require_cuda_build=False)
    batch = to.synthetic_batch(2, 5, 64, seed=21)
    ref = _reference_step(batch)
    got = _dropin_step(batch, "cpu", emu=emu, accelerate=True)
    assert isinstance(got[1], LazyMetrics) and dict(got[1]) == ref[1]
    _compare(ref, got)
    # what the patch refuses: a trainer on the plain-CE branch of compute_loss, a non-Adam optimizer
    model = rt.build_reference_model(CFG, "cpu", seed=0)
    tr = rt.make_trainer(model, "cpu")
    tr.use_mse = False
    with pytest.raises(ValueError):
        accelerate_trainer(tr, _lib=emu)
    tr.use_mse = True
    tr.optimizer = torch.optim.SGD(model.parameters(), lr=0.1)
    with pytest.raises(ValueError):
        accelerate_trainer(tr, _lib=emu)
    # loss only: the trainer keeps its own _process_batch (torch clip + Adam) and gets the fused loss
    tr = rt.make_trainer(model, "cpu")
    accelerate_trainer(tr, fuse_optimizer=False, _lib=emu)
    assert not isinstance(tr.optimizer, ClipAdam) and "_process_batch" not in vars(tr) and "compute_loss" in vars(tr)


@pytest.mark.gpu
def test_accelerated_trainer_on_the_gpu():
    """The same on the CUDA library, then three more steps next to the reference trainer (eager, capture, replay paths)."""
    from oracle import ref_trainer as rt
    from videocad_b200 import ModelFactory
    from videocad_b200.trainer_accel import accelerate_trainer

    batch = to.synthetic_batch(3, 7, 224, seed=22)
    _compare(_reference_step(batch), _dropin_step(batch, "cuda", accelerate=True))
    ref_model = rt.build_reference_model(CFG, "cpu", seed=0)
    ref_model.eval()
    ref_tr = rt.make_trainer(ref_model, "cpu", lr=1e-4)
    model, _ = ModelFactory().create_model("autoregressive", dict(CFG, state_dim=1644, act_dim=7, encoder="vit"), "cuda",
                                           state_dict=to.seeded_state_dict(CFG, 0))
    model.eval()
    got_tr = accelerate_trainer(rt.make_trainer(torch.compile(model, dynamic=False), "cuda", lr=1e-4))
    for step in range(4):
        l0, m0 = ref_tr._process_batch(batch)
        l1, m1 = got_tr._process_batch(batch)
        assert abs(l0.item() - l1.item()) < 2e-3 * abs(l0.item()), (step, l0.item(), l1.item())
        assert m0["total_predictions"] == m1["total_predictions"] and len(m1) == len(m0)


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
