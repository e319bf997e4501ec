"""World-size-2 data-parallel path on REAL GPUs (NCCL over NVLink): the CUDA kernels inside the reference's DDP wrapper
(experiment.py:104-109, find_unused_parameters=True) must produce the gradient of the full batch (the average of the
per-shard gradients), on the eager, CUDA-graph capture and replay paths.  Skips on a box with fewer than two GPUs
(run it with `gpurun --gpus 2`); tests/test_ddp_gloo.py covers the same host logic on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_oracle as to

pytestmark = pytest.mark.gpu

CFG = dict(hidden_size=256, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=3,
           enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
B, T, S = 8, 6, 224


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(device):
    from videocad_b200 import AutoRegressiveTransformer

    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", dropout=0.0, vit_dropout=0.0, **CFG)
    m.load_state_dict(to.seeded_state_dict(CFG, 0), strict=False)
    return m.to(device).train()


def _loss(model, inp):
    cmds, params = model(inp)
    g = torch.Generator().manual_seed(11)
    w = torch.randn(params.shape[1:], generator=g).to(params.device)
    return (cmds.sum(dim=(1, 2)) + (params * w).sum(dim=(1, 2, 3))).mean()


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        m = _build(dev)
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank], output_device=rank, find_unused_parameters=True)
        inp = to.model_inputs_from_batch(to.synthetic_batch(B, T, S, seed=77))
        per = B // world
        shard = {k: v[rank * per:(rank + 1) * per].to(dev) for k, v in inp.items()}
        grads = []
        for it in range(3):  # eager, capture, replay
            m.zero_grad(set_to_none=True)
            _loss(ddp, shard).backward()
            torch.cuda.synchronize()
            grads.append({k: p.grad.detach().cpu().clone() for k, p in m.named_weights() if p.grad is not None})
        if rank == 0:
            torch.save(grads, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_ddp_world2_nccl_gradients_equal_full_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out_path = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    m = _build(torch.device("cuda", 0))
    inp = {k: v.cuda() for k, v in to.model_inputs_from_batch(to.synthetic_batch(B, T, S, seed=77)).items()}
    _loss(m, inp).backward()
    ref = {k: p.grad.detach().cpu() for k, p in m.named_weights() if p.grad is not None}
    assert len(ref) > 100
    for it, g in enumerate(got):
        for k, r in ref.items():
            assert k in g, (it, k)
            assert (g[k] - r).abs().max() <= 1e-4 * r.abs().max() + 1e-7, (it, k)  # split-K atomics and the all-reduce reorder sums
