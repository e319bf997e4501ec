"""Fused optimizer step for the flat-parameter model: `clip_grad_norm_(params, max_norm)` + `torch.optim.Adam.step()` of
the reference trainer (/root/reference/trainer.py:253, 493-494) as three native kernels.

    opt = ClipAdam(model.parameters(), lr=1e-5, max_norm=1.0)
    loss.backward(); opt.step(); opt.zero_grad()          # step() clips (in place, like clip_grad_norm_) and updates

Same arithmetic as torch's Adam with default betas/eps, no weight decay, no amsgrad (the reference's configuration);
parameter groups with their own `lr` are supported (trainer.py:243-251).  `state_dict()` / `load_state_dict()` are
torch.optim.Optimizer's (state: step, exp_avg, exp_avg_sq), so checkpoints round-trip with torch.optim.Adam.
"""
from __future__ import annotations


import torch

from . import lib as L


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, _lib=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, max_norm=max_norm))
        self._lib = _lib  # tests: CPU emulation library
        self._scratch = None
        self.last_grad_norm = None  # device scalar: total gradient norm before clipping (what clip_grad_norm_ returns)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = self._lib if self._lib is not None else L.load()
        entries, keep = [], []
        g0 = self.param_groups[0]
        for group in self.param_groups:
            if group["betas"] != g0["betas"] or group["eps"] != g0["eps"] or group["max_norm"] != g0["max_norm"]:
                raise ValueError("ClipAdam: betas/eps/max_norm must be the same in every parameter group (lr may differ)")
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda and self._lib is None:
                    raise RuntimeError("ClipAdam runs on CUDA tensors only")
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise ValueError("ClipAdam: contiguous fp32 parameters and gradients required")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                entries.append((p, st, group["lr"]))
        if not entries:
            return loss
        if len(entries) > 16:
            raise ValueError("ClipAdam handles at most 16 tensors per step: use the flat-parameter model (3 tensors)")
        steps = {int(st["step"].item()) for _, st, _ in entries}
        if len(steps) != 1:
            raise ValueError("ClipAdam: all parameters must have the same step count")
        dev = entries[0][0].device
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.empty(lib.vc_clip_adam_scratch_floats(), dtype=torch.float32, device=dev)
        self.last_grad_norm = torch.empty(1, dtype=torch.float32, device=dev)
        arr = (L.AdamTensor * len(entries))()
        for i, (p, st, lr) in enumerate(entries):
            arr[i].p, arr[i].g = p.data_ptr(), p.grad.data_ptr()
            arr[i].m, arr[i].v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            arr[i].n, arr[i].lr = p.numel(), float(lr)
            keep.append((p, st))
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
        max_norm = g0["max_norm"] if g0["max_norm"] is not None else 0.0
        L.check(lib.vc_clip_adam_step(arr, len(entries), float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]),
                                      float(max_norm), steps.pop(), self._scratch.data_ptr(), self.last_grad_norm.data_ptr(), stream), lib)
        # the kernels wrote the parameters (and clipped the gradients) through raw pointers: bump the tensors' version counters
        # as an in-place torch op would, so that everything keyed on `_version` (the model's split-bf16 weight mirror,
        # autograd's saved-tensor checks) sees the update
        torch._C._increment_version([p for p, _ in keep] + [p.grad for p, _ in keep])
        return loss
