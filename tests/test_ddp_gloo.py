"""World-size-2 data-parallel path on CPU (gloo): the drop-in module inside the reference's DDP wrapper
(experiment.py:104-109, find_unused_parameters=True) must produce the average of the per-shard gradients, i.e. the
gradient of the full batch.  Kernels are the CPU emulation (host logic only; the NCCL/NVLink run is bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import build_emu
from oracle import torch_oracle as to

CFG = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2,
           enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(emu_path):
    from videocad_b200 import AutoRegressiveTransformer
    from videocad_b200 import lib as L

    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", dropout=0.0, vit_dropout=0.0, **CFG)
    m.load_state_dict(to.seeded_state_dict(CFG, 0), strict=False)
    m._use_library_for_tests(L.load(emu_path, require_cuda_build=False))
    return m


def _loss(model, inp):
    cmds, params = model(inp)
    g = torch.Generator().manual_seed(11)
    w = torch.randn(params.shape[1:], generator=g)
    return (cmds.sum(dim=(1, 2)) + (params * w).sum(dim=(1, 2, 3))).mean()


def _worker(rank, world, port, emu_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _build(emu_path)
        ddp = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)
        inp = to.model_inputs_from_batch(to.synthetic_batch(4, 4, 64, seed=77))
        shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in inp.items()}
        _loss(ddp, shard).backward()
        if rank == 0:
            torch.save({k: p.grad.clone() for k, p in m.named_weights() if p.grad is not None}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_world2_gloo_gradients_equal_full_batch(tmp_path):
    emu_path = build_emu.build()
    out_path = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), emu_path, out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    m = _build(emu_path)
    inp = to.model_inputs_from_batch(to.synthetic_batch(4, 4, 64, seed=77))
    _loss(m, inp).backward()
    assert len(got) > 100
    for k, p in m.named_weights():
        assert k in got, k
        ref = p.grad
        assert (got[k] - ref).abs().max() <= 2e-5 * ref.abs().max() + 1e-7, k


def _trainer_worker(rank, world, port, emu_path, out_path):
    """Two trainers per rank around two DDP-wrapped copies of the drop-in: the unmodified reference trainer, and the same trainer
    after accelerate_trainer (fused loss + ClipAdam).  Both see this rank's shard for two optimiser steps."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ref_trainer as rt
        from videocad_b200 import lib as L
        from videocad_b200.trainer_accel import accelerate_trainer

        emu = L.load(emu_path, require_cuda_build=False)
        batch = to.synthetic_batch(4, 5, 64, seed=31)
        shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in batch.items()}
        out = {}
        for name in ("plain", "accelerated"):
            m = _build(emu_path)
            m.eval()  # no dropout: the two trainers must follow the same trajectory
            ddp = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)
            trainer = rt.make_trainer(ddp, "cpu", lr=1e-3)
            if name == "accelerated":
                accelerate_trainer(trainer, _lib=emu)
            losses = []
            for _ in range(2):
                loss, metrics = trainer._process_batch(shard)
                losses.append(float(loss.item()))
                assert metrics["total_predictions"] > 0
            out[name] = dict(losses=losses, weights={k: v.detach().clone() for k, v in m.state_dict().items()})
        torch.save(out, f"{out_path}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_accelerated_trainer_under_ddp_gloo(tmp_path):
    from oracle import reference_model as rm

    if not rm.available():
        pytest.skip("reference sources neither under /root/reference nor staged in oracle/_ref")
    emu_path = build_emu.build()
    out_path = str(tmp_path / "trainers.pt")
    mp.spawn(_trainer_worker, args=(2, _free_port(), emu_path, out_path), nprocs=2, join=True)
    r0, r1 = torch.load(out_path + ".0"), torch.load(out_path + ".1")
    for name in ("plain", "accelerated"):  # DDP keeps the replicas identical: same averaged gradients, same update
        for k, v in r0[name]["weights"].items():
            assert torch.equal(v, r1[name]["weights"][k]), (name, k)
    p, a = r0["plain"], r0["accelerated"]
    for l0, l1 in zip(p["losses"], a["losses"]):
        assert abs(l0 - l1) < 2e-4 * max(1.0, abs(l0)), (p["losses"], a["losses"])
    assert a["losses"][1] < a["losses"][0]  # the steps trained
    total = differing = 0
    for k, v in p["weights"].items():  # Adam(lr = 1e-3), two steps: elements with rounding-level gradients may move the other way
        d = (a["weights"][k] - v).abs()
        assert d.max().item() <= 4.2e-3, (k, d.max().item())
        total += d.numel()
        differing += int((d > 4e-5).sum())
    assert differing <= 4e-3 * total, (differing, total)
