"""CPU port of one reference training step (TEST INFRASTRUCTURE; the timed CPU baseline, "kind": "port").

Follows BaseTrainer._process_batch (/root/reference/trainer.py:480-496): zero_grad -> forward (training mode,
dropout 0.1) -> MultiClassesTrainer.compute_loss -> backward -> clip_grad_norm_(1.0) -> Adam.step(lr=1e-5), on the
functional oracle (oracle/torch_oracle.py) in plain PyTorch fp32 on the host cores.  The unmodified reference itself
cannot travel to the GPU box (/root/reference does not exist there and vit_pytorch is not installed), hence a port.
"""
import os
import time

import torch

from . import torch_oracle as to
from videocad_b200.loss import compute_loss


class CpuTrainStep:
    def __init__(self, cfg: dict, seed: int = 0, lr: float = 1e-5, dropout_p: float = 0.1):
        self.cfg, self.dropout_p = cfg, dropout_p
        self.sd = {k: v.clone().requires_grad_(True) for k, v in to.seeded_state_dict(cfg, seed).items()}
        self.opt = torch.optim.Adam(list(self.sd.values()), lr=lr)

    def step(self, batch: dict) -> float:
        self.opt.zero_grad()
        inp = to.model_inputs_from_batch(batch)
        preds = to.forward(self.sd, self.cfg, inp, dropout_p=self.dropout_p)
        loss = compute_loss(preds, batch["actions"][:, 1:])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.sd.values()), 1.0)
        self.opt.step()
        return float(loss.item())


def pick_threads(runner: "CpuTrainStep", batch: dict) -> int:
    """Thread count that runs a step fastest on this host: a small batch does not scale to every hardware thread (64 threads
    were 2.4x slower than 8 on the GPU box's 128-CPU host), and torchrun pins OMP_NUM_THREADS=1.  One probe step per candidate."""
    ncpu = os.cpu_count() or 1
    cands = sorted({n for n in (4, 8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= n <= ncpu})
    best, best_dt = torch.get_num_threads(), None
    for n in cands:
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        runner.step(batch)
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = n, dt
    torch.set_num_threads(best)
    return best


def time_cpu_train(cfg: dict, B: int, T: int, S: int, steps: int = 2, warmup: int = 1):
    """frames/s of the CPU port on the host cores (thread count chosen by pick_threads)."""
    runner = CpuTrainStep(cfg)
    batches = [to.synthetic_batch(B, T + 1, S, seed=100 + i) for i in range(2)]
    runner.step(batches[0])  # first step pays allocator / thread-pool start-up
    pick_threads(runner, batches[1])
    for i in range(warmup):
        runner.step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        runner.step(batches[i % 2])
    dt = time.perf_counter() - t0
    return dict(value=B * T * steps / dt, unit="frames/s", cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                kind="port", sample=f"{steps} training steps (fwd+loss+bwd+clip+Adam, dropout {runner.dropout_p}) of the "
                f"same model at batch {B} x T={T} x {S}x{S} after {warmup} warm-up", seconds=dt)
