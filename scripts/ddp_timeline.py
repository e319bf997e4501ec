"""Kernel timeline of one DDP training step (rank 0): which kernels and NCCL all-reduces run in the tail of the backward.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/ddp_timeline.py --config c1 --out FILE"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from videocad_b200 import AutoRegressiveTransformer
from videocad_b200 import loss as vloss
from videocad_b200.optim import ClipAdam

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c1")
ap.add_argument("--out", default="gpurun_out/ddp_timeline.txt")
ap.add_argument("--no-sync", action="store_true")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl")
torch.manual_seed(0)
model = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", **cfg["model"]).to(dev).train()
net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[lr], output_device=lr, find_unused_parameters=True)
opt = ClipAdam(model.parameters(), lr=1e-5, max_norm=1.0)
B, T, S = cfg["B"], cfg["T"], cfg["S"]
batches = bench.make_batches(B, T, S, rank, device=dev)


def step(b):
    opt.zero_grad()
    acts = b["actions"]
    inputs = {"frames": b["frames"][:, :-1], "actions": model.normalize_actions(acts[:, :-1].clone()), "cad_image": b["cad_image"]}
    loss = vloss.compute_loss_fused(net(inputs), acts[:, 1:])
    loss.backward()
    opt.step()


import contextlib
ctx = net.no_sync() if args.no_sync else contextlib.nullcontext()
with ctx:
    for i in range(8):
        step(batches[i % len(batches)])
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            step(batches[i % len(batches)])
        torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    with open(args.out, "w") as f:
        f.write("start_us dur_us stream name\n")
        for e in evs:
            f.write(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f} {getattr(e, 'stream', -1)!s:>4} {e.name[:110]}\n")
    print("kernels", len(evs), "span_us", evs[-1].time_range.end - t0)
dist.barrier()
dist.destroy_process_group()
