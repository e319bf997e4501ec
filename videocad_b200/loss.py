"""Training loss of the reference trainer, restated without host synchronisation.

Follows MultiClassesTrainer.compute_loss (/root/reference/trainer.py:935-966) with `use_mse=True` (main.py:96):

    loss = 2 * CE_w(cmds, cmd_tgt; ignore -1, weight = class_weights["Label"])
         + sum_i  cmd_w[param_to_label[i]] * flexible_cross_entropy(params[..., i, :], tgt_i, tolerance_i)

`flexible_cross_entropy` (trainer.py:853-916) is called with `above=self.above` -- the whole list, always truthy
 -- so every parameter uses the one-sided window [t, t + tol): rows whose argmax already falls
in the window are dropped, the soft target is uniform over the (clamped) window, the mean runs over the remaining
rows; an empty selection contributes 0 and NaN terms are skipped.

The reference builds the soft targets with Python loops and ~30 `.item()` calls per step.  Two implementations here:

* `compute_loss`          -- the torch restatement (class-index masks instead of boolean indexing, nothing leaves the
                             device); ~300 small kernels per step with its autograd backward.  Kept as the readable
                             definition and as the checker of the fused kernel (tests/test_loss.py).
* `compute_loss_fused`    -- SURVEY.md 8(f) rank 1: the same loss as three native kernels (row pass, finalize, gradient)
                             behind `vc_loss_forward` / `vc_loss_backward` (videocad_b200/csrc/loss.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

from . import lib as L

TOLERANCE = 3
TOLERANCES = (TOLERANCE - 1, TOLERANCE - 1, 50, 200, 500, TOLERANCE - 1)
PARAM_TO_LABEL = (0, 0, 1, 1, 2, 3)
# class_weights.json["Label"] of the reference (5 numbers; data, not code)
DEFAULT_CMD_WEIGHTS = (0.04332685213392362, 0.02915898563179938, 0.267566828114559, 0.6005346809501417, 0.05941265316957628)


def flexible_cross_entropy(logits: torch.Tensor, targets: torch.Tensor, tolerance: int, num_classes: int = 1000,
                           ignore_index: int = -1) -> torch.Tensor:
    logits = logits.reshape(-1, num_classes)
    targets = targets.reshape(-1)
    valid_row = targets != ignore_index
    # allowed classes = {clamp(target + o, 0, C - 1) : 0 <= o < tolerance} (trainer.py:880-905): both ends clamped separately,
    # so the window is never empty (an out-of-range target keeps the single class C - 1, a negative one the class 0)
    t = targets.clamp(min=0, max=num_classes - 1)
    hi = (targets + (tolerance - 1)).clamp(min=0, max=num_classes - 1)
    preds = logits.argmax(dim=1)
    in_window = (preds >= t) & (preds <= hi)
    select = valid_row & ~in_window
    logp = F.log_softmax(logits.float(), dim=1)
    cls = torch.arange(num_classes, device=logits.device).unsqueeze(0)
    window = ((cls >= t.unsqueeze(1)) & (cls <= hi.unsqueeze(1))).to(logp.dtype)
    count = (hi - t + 1).to(logp.dtype)
    per_row = -(logp * window).sum(dim=1) / count
    sel = select.to(logp.dtype)
    n = sel.sum()
    val = (per_row * sel).sum() / n.clamp(min=1.0)
    return val


_W_CACHE = {}


def _weights_on(device, values):
    """class weights as a device tensor, built once per (device, values): torch.tensor(..., device=cuda) is a pageable
    host-to-device copy, i.e. a stream synchronisation in the middle of every training step."""
    key = (str(device), values)
    w = _W_CACHE.get(key)
    if w is None:
        w = torch.tensor(values, dtype=torch.float32, device=device)
        _W_CACHE[key] = w
    return w


def compute_loss(action_preds, actions: torch.Tensor, cmd_weights: Optional[Sequence[float]] = None) -> torch.Tensor:
    """action_preds = (cmds [B,T,5], params [B,T,6,1000]); actions [B,T,7] raw targets (-1 = ignore)."""
    pred_cmd, pred_params = action_preds
    actions = actions.long()
    w = _weights_on(pred_cmd.device, tuple(cmd_weights) if cmd_weights is not None else DEFAULT_CMD_WEIGHTS)
    loss_cmd = F.cross_entropy(pred_cmd.reshape(-1, pred_cmd.shape[-1]), actions[..., 0].reshape(-1), weight=w, ignore_index=-1)
    loss = 2.0 * loss_cmd
    for i in range(pred_params.shape[-2]):
        lp = flexible_cross_entropy(pred_params[..., i, :], actions[..., 1 + i], TOLERANCES[i], pred_params.shape[-1])
        lp = torch.where(torch.isnan(lp), torch.zeros_like(lp), lp)
        loss = loss + lp * w[PARAM_TO_LABEL[i]]
    return loss


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cmds, params, targets, cfg, lib, metrics):
        cmds, params, targets = cmds.contiguous().float(), params.contiguous().float(), targets.contiguous().float()
        dev = cmds.device
        stream = torch.cuda.current_stream(dev).cuda_stream if cmds.is_cuda else None
        ws = torch.empty(lib.vc_loss_workspace_floats(cfg.R, cfg.NP), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        L.check(lib.vc_loss_forward(C.byref(cfg), cmds.data_ptr(), params.data_ptr(), targets.data_ptr(), ws.data_ptr(), loss.data_ptr(),
                                    stream), lib)
        counts = torch.empty(L.METRIC_COUNT if metrics is not None else 0, dtype=torch.int64, device=dev)
        if metrics is not None:
            mcfg, T = metrics
            L.check(lib.vc_loss_metrics(C.byref(cfg), C.byref(mcfg), targets.data_ptr(), ws.data_ptr(), T, counts.data_ptr(), stream), lib)
        ctx.save_for_backward(cmds, params, targets, ws)
        ctx.cfg, ctx.lib = cfg, lib
        ctx.mark_non_differentiable(counts)
        return loss.view(()), counts

    @staticmethod
    def backward(ctx, g, _gcounts=None):
        cmds, params, targets, ws = ctx.saved_tensors
        cfg, lib = ctx.cfg, ctx.lib
        g = g.contiguous().float().view(1)
        dcmds, dparams = torch.empty_like(cmds), torch.empty_like(params)
        stream = torch.cuda.current_stream(cmds.device).cuda_stream if cmds.is_cuda else None
        L.check(lib.vc_loss_backward(C.byref(cfg), cmds.data_ptr(), params.data_ptr(), targets.data_ptr(), ws.data_ptr(), g.data_ptr(),
                                     dcmds.data_ptr(), dparams.data_ptr(), stream), lib)
        return dcmds, dparams, None, None, None, None


ABOVE = (False, False, True, True, True, False)   # trainer.py:829
TOPK = 30                                          # trainer.py:1003


def metrics_from_counts(counts, NC: int = 5, NP: int = 6) -> dict:
    """The `metrics` dict of MultiClassesTrainer.compute_loss (trainer.py:1036-1058) from the counters of vc_loss_metrics:
    ONE device-to-host copy (`counts` may already be a host tensor / list) instead of ~30 `.item()` calls."""
    c = counts.tolist() if torch.is_tensor(counts) else list(counts)
    m = {
        "correct_predictions": c[L.METRIC_CORRECT], "total_predictions": c[L.METRIC_TOTAL],
        "cmd_corrects": c[L.METRIC_CMD_CORRECTS:L.METRIC_CMD_CORRECTS + NC], "cmd_counts": c[L.METRIC_CMD_COUNTS:L.METRIC_CMD_COUNTS + NC],
        "param_corrects": c[L.METRIC_PARAM_CORRECTS:L.METRIC_PARAM_CORRECTS + NP],
        "param_counts": c[L.METRIC_PARAM_COUNTS:L.METRIC_PARAM_COUNTS + NP],
        "cmd_correct_topk": c[L.METRIC_CMD_CORRECT_TOPK], "cmd_counts_topk": c[L.METRIC_CMD_COUNTS_TOPK],
        "param_correct_topk": c[L.METRIC_PARAM_CORRECT_TOPK], "param_counts_topk": c[L.METRIC_PARAM_COUNTS_TOPK],
        "perfect_sequences": 0, "perfect_commands": 0, "total_sequences": 0, "perfect_sequence_accuracy": 0,
    }
    for i in range(NP):
        m[f"param_corrects_{i}"], m[f"param_counts_{i}"] = m["param_corrects"][i], m["param_counts"][i]
    for i in range(NC):
        m[f"cmd_corrects_{i}"], m[f"cmd_counts_{i}"] = m["cmd_corrects"][i], m["cmd_counts"][i]
    return m


def compute_loss_and_metrics_fused(action_preds, actions: torch.Tensor, cmd_weights: Optional[Sequence[float]] = None, _lib=None):
    """(loss, counts): `compute_loss_fused` plus the metrics half of compute_loss (trainer.py:968-1061) as a device tensor of
    integer counters (one extra single-CTA kernel, no host synchronisation); `metrics_from_counts(counts)` builds the
    reference's dict with a single copy whenever the caller wants the numbers on the host."""
    return compute_loss_fused(action_preds, actions, cmd_weights, _lib=_lib, _with_metrics=True)


def compute_loss_fused(action_preds, actions: torch.Tensor, cmd_weights: Optional[Sequence[float]] = None, _lib=None,
                       _with_metrics: bool = False):
    """Same value and gradients as `compute_loss`, computed by the native fused kernels (no CPU fallback: `_lib` is the
    tests' hook for the CPU emulation library)."""
    pred_cmd, pred_params = action_preds
    if not pred_cmd.is_cuda and _lib is None:
        raise RuntimeError("compute_loss_fused runs on CUDA tensors only")
    lib = _lib if _lib is not None else L.load()
    NC, NP, NV = pred_cmd.shape[-1], pred_params.shape[-2], pred_params.shape[-1]
    w = tuple(cmd_weights) if cmd_weights is not None else DEFAULT_CMD_WEIGHTS
    if NP > len(TOLERANCES) or NP > 8 or NC > 16 or len(w) < NC:
        raise ValueError("compute_loss_fused: unsupported head sizes")
    cfg = L.LossCfg()
    cfg.R, cfg.NC, cfg.NP, cfg.NV = pred_cmd.numel() // NC, NC, NP, NV
    for i in range(NC):
        cfg.cmd_w[i] = float(w[i])
    for i in range(NP):
        cfg.tolerance[i] = TOLERANCES[i]
        cfg.param_to_label[i] = PARAM_TO_LABEL[i]
    metrics = None
    if _with_metrics:
        if actions.dim() != 3:
            raise ValueError("metrics need actions as [B, T, 1 + NP] (the top-k counters are per time step)")
        mcfg = L.MetricsCfg()
        for i in range(NP):
            mcfg.above[i] = int(ABOVE[i])
            mcfg.tolerance[i] = TOLERANCES[i]
        mcfg.abs_tolerance, mcfg.topk = TOLERANCE, TOPK
        metrics = (mcfg, int(actions.shape[1]))
    loss, counts = _FusedLoss.apply(pred_cmd.reshape(-1, NC), pred_params.reshape(-1, NP, NV), actions.reshape(-1, 1 + NP), cfg, lib, metrics)
    return (loss, counts) if _with_metrics else loss
