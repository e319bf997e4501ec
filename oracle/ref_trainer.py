"""Drive the UNMODIFIED reference trainer (TEST INFRASTRUCTURE ONLY; also the timed reference arm of bench.py).

`create_trainer(...)._process_batch(batch)` of /root/reference/trainer.py:480-496 -- zero_grad -> prepare_batch -> model
forward (training mode) -> MultiClassesTrainer.compute_loss -> backward -> clip_grad_norm_(1.0) -> Adam.step -- on

  * the reference's own model (ModelFactory.create_model, with the vit_pytorch shim)            -> the CPU arm, "kind": "reference"
  * the drop-in module from videocad_b200 (handed in by the caller)                            -> the through-trainer GPU leg

The reference code comes from /root/reference where it exists, else from the copies staged by oracle/build_ref.py
(oracle/reference_model.py picks the root).  The trainer insists on `class_weights.json` in the working directory and
creates `logs/` / `checkpoints/` there (trainer.py:822, :87-99): everything happens inside a private temporary directory.
"""
from __future__ import annotations

import contextlib
import io
import os
import shutil
import tempfile
import time

import torch

from . import reference_model as rm
from . import torch_oracle as to


@contextlib.contextmanager
def _workdir():
    """A scratch working directory holding class_weights.json (the trainer reads it relative to the CWD)."""
    d = tempfile.mkdtemp(prefix="vc_ref_trainer_")
    shutil.copy(os.path.join(rm.REFERENCE_ROOT, "class_weights.json"), os.path.join(d, "class_weights.json"))
    old = os.getcwd()
    os.chdir(d)
    try:
        yield d
    finally:
        os.chdir(old)
        shutil.rmtree(d, ignore_errors=True)


def make_trainer(model, device, lr: float = 1e-5, quiet: bool = True):
    """create_trainer (trainer.py:1384) around `model` with empty loader packets; the caller feeds batches to _process_batch."""
    tr = rm.import_trainer()
    from model.model_factory import ModelType  # type: ignore  (the reference's enum; rm.import_trainer put it on sys.path)

    pk = {"loader": [], "sampler": None}
    cfg = {"lr": lr, "use_mse": True, "experiment_name": "videocad_b200_bench", "epochs": 1, "save_frequency": 10 ** 9,
           "val_frequency": 10 ** 9}
    with _workdir():
        out = io.StringIO()
        with contextlib.redirect_stdout(out if quiet else os.sys.stdout):
            return tr.create_trainer(pk, pk, pk, model, cfg, device, ModelType.MULTI_CLASSES)


def build_reference_model(cfg: dict, device="cpu", seed: int = 0):
    """The reference's own model with the oracle's seeded weights (same weights the drop-in is loaded with in the tests)."""
    model, _ = rm.build_reference_model(dict(cfg, encoder="vit"), device)
    model.load_state_dict(to.seeded_state_dict(cfg, seed), strict=False)
    return model


class ReferenceTrainStep:
    """One reference training step on the host cores: the real trainer around the real model."""

    def __init__(self, cfg: dict, seed: int = 0, lr: float = 1e-5):
        self.model = build_reference_model(cfg, "cpu", seed)
        self.model.train()
        self.trainer = make_trainer(self.model, "cpu", lr)

    def step(self, batch: dict) -> float:
        loss, _ = self.trainer._process_batch(batch)
        return float(loss.item())


def pick_threads(step, batch, cands=None) -> int:
    """Thread count that runs a step fastest on this host (torchrun pins OMP_NUM_THREADS=1; a 128-CPU host is not fastest with
    128 threads).  One probe step per candidate."""
    ncpu = os.cpu_count() or 1
    cands = cands or sorted({n for n in (8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= n <= ncpu})
    best, best_dt = torch.get_num_threads(), None
    for n in cands:
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        step(batch)
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = n, dt
    torch.set_num_threads(best)
    return best


def time_reference_train(cfg: dict, B: int, T: int, S: int, steps: int = 2, warmup: int = 1, budget_s: float = 25.0):
    """frames/s of the real reference trainer on the host cores.  `budget_s` bounds the timed sample: steps are cut so that the
    timed region stays near the budget (at least one step)."""
    runner = ReferenceTrainStep(cfg)
    batches = [to.synthetic_batch(B, T + 1, S, seed=100 + i) for i in range(2)]
    runner.step(batches[0])  # first step pays allocator / thread-pool start-up
    pick_threads(runner.step, batches[1])
    t0 = time.perf_counter()
    for i in range(max(warmup, 1)):
        runner.step(batches[i % 2])
    per_step = (time.perf_counter() - t0) / max(warmup, 1)
    steps = max(1, min(steps, int(budget_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    for i in range(steps):
        runner.step(batches[i % 2])
    dt = time.perf_counter() - t0
    return dict(value=B * T * steps / dt, unit="frames/s", cores=torch.get_num_threads(), host_cpus=os.cpu_count(), kind="reference",
                sample=f"{steps} steps of the unmodified reference trainer._process_batch (reference model, vit_pytorch shim) at "
                       f"batch {B} x T={T} x {S}x{S} after {max(warmup, 1)} warm-up", seconds=dt, steps=steps)
