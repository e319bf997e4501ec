"""Training loss of the reference trainer, restated without host synchronisation.

Follows MultiClassesTrainer.compute_loss (/root/reference/trainer.py:935-966) with `use_mse=True` (main.py:96):

    loss = 2 * CE_w(cmds, cmd_tgt; ignore -1, weight = class_weights["Label"])
         + sum_i  cmd_w[param_to_label[i]] * flexible_cross_entropy(params[..., i, :], tgt_i, tolerance_i)

`flexible_cross_entropy` (trainer.py:853-916) is called with `above=self.above` -- the whole list, always truthy
 -- so every parameter uses the one-sided window [t, t + tol): rows whose argmax already falls
in the window are dropped, the soft target is uniform over the (clamped) window, the mean runs over the remaining
rows; an empty selection contributes 0 and NaN terms are skipped.

The reference builds the soft targets with Python loops and ~30 `.item()` calls per step; here the window is a
class-index mask, masks replace boolean indexing and nothing leaves the device.  This module is
host-side glue around the native model (SURVEY.md 8(f) rank 1 lists a fused kernel for it as the next row).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F

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
    t = targets.clamp(min=0)
    hi = (t + (tolerance - 1)).clamp(max=num_classes - 1)        # window [t, hi], clamped like the reference
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
    return (per_row * sel).sum() / n.clamp(min=1.0)


def compute_loss(action_preds, actions: torch.Tensor, cmd_weights: Optional[Sequence[float]] = None) -> torch.Tensor:
    """action_preds = (cmds [B,T,5], params [B,T,6,1000]); actions [B,T,7] raw targets (-1 = ignore)."""
    pred_cmd, pred_params = action_preds
    actions = actions.long()
    w = torch.tensor(cmd_weights if cmd_weights is not None else DEFAULT_CMD_WEIGHTS, dtype=torch.float32, device=pred_cmd.device)
    loss_cmd = F.cross_entropy(pred_cmd.reshape(-1, pred_cmd.shape[-1]), actions[..., 0].reshape(-1), weight=w, ignore_index=-1)
    loss = 2.0 * loss_cmd
    for i in range(pred_params.shape[-2]):
        lp = flexible_cross_entropy(pred_params[..., i, :], actions[..., 1 + i], TOLERANCES[i], pred_params.shape[-1])
        lp = torch.where(torch.isnan(lp), torch.zeros_like(lp), lp)
        loss = loss + lp * w[PARAM_TO_LABEL[i]]
    return loss
