"""Fused loss / metrics / optimizer step behind an UNMODIFIED reference trainer object.

The reference trainer (/root/reference/trainer.py) runs unmodified around the drop-in model, but its own `compute_loss`
(trainer.py:935-1063: Python loops, ~30 `.item()` host synchronisations, ~300 small torch kernels with their autograd) and
`clip_grad_norm_` + `Adam.step()` (trainer.py:493-494) then take 85 % of a C1 step.  `accelerate_trainer(trainer)` swaps, on
the INSTANCE (no source edit, no class patch):

  * `trainer.compute_loss`   -> `compute_loss_and_metrics_fused` (csrc/loss.cu): same loss value and gradients, same metrics
                                dict -- delivered as a dict that copies its 54 counters to the host when first read, i.e. after
                                the backward and the optimizer step have been queued (the training loop reads it in
                                `update_metrics`, trainer.py:451, 1287-1310);
  * `trainer.optimizer`      -> `ClipAdam` (csrc/optim.cu) over the same parameter groups / learning rates / betas / eps;
  * `trainer._process_batch` -> the same eight statements as trainer.py:480-496 without the separate `clip_grad_norm_` call
                                (ClipAdam clips to the same max_norm = 1.0 inside its step).

Everything else of the trainer (loops, logging, validation, checkpoints -- `ClipAdam.state_dict()` is torch.optim.Optimizer's)
is untouched.  Configurations the fused kernels do not cover are refused with a ValueError instead of silently diverging.
"""
from __future__ import annotations

import torch

from . import loss as vloss
from .optim import ClipAdam

CLIP_MAX_NORM = 1.0  # trainer.py:493


class LazyMetrics(dict):
    """The metrics dict of `compute_loss` (trainer.py:1036-1061), filled from the device counters on first read.

    The counters start their device-to-host copy (pinned, asynchronous) when the object is made; the host waits for it only
    when a value is needed."""

    def __init__(self, counts: torch.Tensor):
        super().__init__()
        self._event = None
        if counts.is_cuda:
            self._host = torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True)
            self._host.copy_(counts, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(counts.device))
        else:
            self._host = counts
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            if self._event is not None:
                self._event.synchronize()
            super().update(vloss.metrics_from_counts(self._host))
            self._host = self._event = None

    def __getitem__(self, k):
        self._fill()
        return super().__getitem__(k)

    def __contains__(self, k):
        self._fill()
        return super().__contains__(k)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()

    def __eq__(self, other):
        self._fill()
        return super().__eq__(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._fill()
        return super().__repr__()

    def get(self, k, default=None):
        self._fill()
        return super().get(k, default)

    def keys(self):
        self._fill()
        return super().keys()

    def values(self):
        self._fill()
        return super().values()

    def items(self):
        self._fill()
        return super().items()

    def copy(self):
        self._fill()
        return dict(self)


def _check_loss_config(trainer):
    if not getattr(trainer, "use_mse", False):
        raise ValueError("accelerate_trainer: the fused loss is the `use_mse=True` branch of compute_loss (main.py:96); this "
                         "trainer was built with use_mse=False")
    if tuple(getattr(trainer, "tolerances", ())) != vloss.TOLERANCES or tuple(getattr(trainer, "above", ())) != vloss.ABOVE \
            or tuple(getattr(trainer, "param_to_label", ())) != vloss.PARAM_TO_LABEL:
        raise ValueError("accelerate_trainer: tolerances / above / param_to_label differ from trainer.py:827-829")
    w = getattr(trainer, "cmd_weights", None)
    if w is None or len(w) != 5:
        raise ValueError("accelerate_trainer: trainer.cmd_weights must hold the 5 command-class weights (class_weights.json['Label'])")
    return tuple(float(x) for x in w)


def _clip_adam_like(old: torch.optim.Optimizer, _lib=None) -> ClipAdam:
    if not isinstance(old, torch.optim.Adam):
        raise ValueError("accelerate_trainer: trainer.optimizer is not torch.optim.Adam (trainer.py:251-253)")
    groups = []
    for g in old.param_groups:
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("accelerate_trainer: ClipAdam is plain Adam (no weight decay / amsgrad / maximize)")
        groups.append(dict(params=list(g["params"]), lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"])))
    new = ClipAdam(groups, max_norm=CLIP_MAX_NORM, _lib=_lib)
    for p, st in old.state.items():  # the reference never restores optimizer state; carried over for completeness
        if st:
            new.state[p] = dict(step=torch.as_tensor(float(st["step"])), exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
    return new


def accelerate_trainer(trainer, fuse_loss: bool = True, fuse_optimizer: bool = True, _lib=None):
    """Patch `trainer` (an instance built by the reference's `create_trainer`, trainer.py:1384) in place and return it.

    `_lib` is the tests' hook for the CPU emulation library; the product path loads the CUDA library and raises without it."""
    if fuse_loss:
        weights = _check_loss_config(trainer)

        def compute_loss(action_preds, actions, mse=True):
            loss, counts = vloss.compute_loss_and_metrics_fused(action_preds, actions, cmd_weights=weights, _lib=_lib)
            return loss, LazyMetrics(counts)

        trainer.compute_loss = compute_loss
    if fuse_optimizer:
        trainer.optimizer = _clip_adam_like(trainer.optimizer, _lib=_lib)

        def _process_batch(batch, noise=False):  # trainer.py:480-496
            trainer.optimizer.zero_grad()
            batch_dict = trainer.prepare_batch(batch)
            if noise:
                batch_dict["actions"] = trainer._add_noise_to_actions(batch_dict["actions"])
            model_inputs = trainer._prepare_model_inputs(batch_dict, noise)
            action_preds = trainer.model(model_inputs)
            loss, batch_metrics = trainer.compute_loss(action_preds, batch_dict["actions"][:, 1:])
            loss.backward()
            trainer.optimizer.step()  # clip_grad_norm_(1.0) + Adam.step() in one pass (csrc/optim.cu)
            return loss, batch_metrics

        trainer._process_batch = _process_batch
    return trainer
