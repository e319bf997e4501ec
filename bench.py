#!/usr/bin/env python
"""bench.py -- train frames/sec of the VideoCAD behaviour-cloning hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config c1|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one optimiser step of the reference trainer (trainer.py:480-496: zero_grad -> model forward in training
mode (dropout 0.1) -> compute_loss -> backward -> clip_grad_norm_(1.0) -> Adam(lr=1e-5).step) over one synthetic batch of
the named shape; frames = B*T model frames per step, summed over ranks (weak scaling: B per GPU is fixed).

  value      whole-job frames/s with inputs already resident in HBM when the timed region starts (CUDA events, barrier +
             synchronize on both sides, max over ranks)
  e2e        the same metric through the public API with HOST (pinned) inputs: H2D copy of frames/actions/cad and a
             D2H read of the loss inside the timed region, every step
  roofline   the tcgen05 GEMM kernel (dominant kernel): algorithmic FLOPs (2*M*N*K per launch, counted by the library)
             / summed CUDA-event launch durations recorded around every launch on the launching stream during the timed
             region, against the measured bf16 tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own trainer._process_batch around the reference's own model on the host cores (bounded sample;
             "kind": "reference" when the staged reference sources are present -- oracle/build_ref.py -- else the oracle's port)
  through_trainer  the UNMODIFIED reference trainer object (its compute_loss with the .item() syncs, torch clip + Adam) driving the
             drop-in on the GPU: what a user of trainer.py gets without touching it
  through_trainer_fused  the same trainer object after videocad_b200.trainer_accel.accelerate_trainer(trainer): fused loss + metrics
             kernels and ClipAdam behind the unmodified training loop; loss and metrics read on the host every step
  rollout    BASELINE config C4 on this GPU: 186-step action-feedback rollout of the H=1024 model, 8 sequences, with the decode
             step's HBM roofline (weight + cache bytes per step / step time / measured copy bandwidth)
  c3         (N > 1 only) BASELINE config C3 -- T=32, H=1024, 32 samples per GPU under DDP -- timed in the same launch

--impl reference times the reference's CPU implementation only (rank 0), printing the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1]: 1xB200, 8-frame context, 224x224, d_model=512, batch=32
    "c1": dict(model=dict(hidden_size=512, nhead=4, num_decoder_layers=8, dim_feedforward=512, window_size=10,
                          enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
               B=32, T=8, S=224, cpu_B=32),
    # BASELINE.json configs[3]: DDP, 32-frame context, H=1024, 32 samples per GPU
    "c3": dict(model=dict(hidden_size=1024, nhead=4, num_decoder_layers=8, dim_feedforward=1024, window_size=10,
                          enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
               B=32, T=32, S=224, cpu_B=4),
}
# BASELINE.json configs[4]: 186-step autoregressive rollout of the H=1024 model, 64 sequences over 8 GPUs = 8 per GPU
ROLLOUT = dict(model=CONFIGS["c3"]["model"], B=8, T=186, S=224)
NUM_BATCHES = 4  # distinct synthetic batches rotated through the timed region (4 x 58 MB of frames at c1 > 126 MB L2)


def fwd_flops_per_sample(cfg, T, S):
    """SURVEY.md 8(d) / BASELINE.md section 3: algorithmic forward FLOPs per sample."""
    H, Ff, L = cfg["hidden_size"], cfg["dim_feedforward"], cfg["num_decoder_layers"]
    N = (S // 32) ** 2
    n = N + 1
    f_vit = 2 * N * 1024 * 512 + 6 * (2 * n * 512 * 3072 + 4 * n * n * 1024 + 2 * n * 1024 * 512 + 4 * n * 512 * 512)
    glue = 2 * T * 512 * H + 2 * 512 * H + 2 * T * 2 * H * H + 2 * T * 7 * H
    dec = L * (16 * T * H * H + 4 * T * H * Ff + 8 * T * T * H)
    head = 2 * T * H * 6005
    return T * f_vit + f_vit + glue + dec + head


def pair_kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the `ncu --set full` capture
    summarised under profiles/ (profiles/ncu_traffic.json names the capture it was read from); None until a capture exists."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)["gemm_tc_pair_kernel"]
        return int(d["dram_bytes_per_launch"]), d.get("note", "")
    except Exception:
        return None, "no capture recorded in profiles/ncu_traffic.json"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), hbm_gbs=float(d.get("hbm_gbs", 6500.0)),
                    source="measured (MEASURED_PEAKS.json, sustained bf16 / copy bandwidth)")
    return dict(tflops=1400.0, hbm_gbs=6500.0, source="fallback (B200_PROFILING.md sustained figures)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region (one looping nvidia-smi process)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.strip()]
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


def make_batches(B, T, S, rank, device=None, pinned=False):
    from videocad_b200.synthetic import synthetic_batch

    out = []
    for i in range(NUM_BATCHES):
        b = synthetic_batch(B, T + 1, S, seed=1234 + 1000 * rank + i)
        if device is not None:
            b = {k: v.to(device) for k, v in b.items()}
        elif pinned:
            b = {k: v.pin_memory() for k, v in b.items()}
        out.append(b)
    return out


def workload_config(args, cfg, world):
    """The `config` object of the JSON line: the same for both arms (they measure the same workload)."""
    B, T, S = cfg["B"], cfg["T"], cfg["S"]
    return dict(workload=f"{args.config}: {world} x (batch {B}, T={T}, {S}x{S} frames), H={cfg['model']['hidden_size']}, "
                         f"8 decoder layers, window 10, dropout 0.1, Adam lr 1e-5, clip 1.0",
                global_batch=world * B, parallelism=f"dp{world}" if world > 1 else "single",
                l2="inputs rotate over 4 distinct batches (232 MB of frames at c1 > 126 MB L2); activations (>5 GB/step) stream through HBM",
                fwd_gflop_per_sample=fwd_flops_per_sample(cfg["model"], T, S) / 1e9)


def reference_available():
    from oracle import reference_model as rm

    return rm.available()


def run_reference(args, cfg):
    """The reference arm: the reference's OWN CPU implementation of the path -- the unmodified trainer._process_batch
    (trainer.py:480-496) around the unmodified model (vit_pytorch shim) -- on the host cores, rank 0 only.  Falls back to the
    oracle's port only if the reference sources were not staged (oracle/build_ref.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_oracle as to

    B, T, S = cfg["cpu_B"], cfg["T"], cfg["S"]
    batches = [to.synthetic_batch(B, T + 1, S, seed=100 + i) for i in range(2)]
    if reference_available():
        from oracle.ref_trainer import ReferenceTrainStep, pick_threads

        runner, kind = ReferenceTrainStep(cfg["model"]), "reference"
        what = "unmodified reference trainer._process_batch around the reference model (vit_pytorch shim)"
    else:
        from oracle.train_port import CpuTrainStep, pick_threads as _pt

        runner, kind = CpuTrainStep(cfg["model"]), "port"
        what = "oracle port of trainer._process_batch (reference sources not staged)"
        pick_threads = lambda step, batch: _pt(runner, batch)  # noqa: E731
    runner.step(batches[0])
    pick_threads(runner.step, batches[1])  # all the host threads that help (torchrun would pin OMP_NUM_THREADS=1)
    for i in range(args.warmup):
        runner.step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        runner.step(batches[i % 2])
    dt = time.perf_counter() - t0
    fps = B * T * args.steps / dt
    sample = (f"{args.steps} steps of the {what} at batch {B} x T={T} x {S}x{S}"
              + ("" if B == cfg["B"] else f" (bounded sample of the {cfg['B']}-sample per-GPU batch)") + f", host_cpus={os.cpu_count()}")
    line = dict(metric="train frames/sec", value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=workload_config(args, cfg, args.gpus),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=torch.get_num_threads(), kind=kind, sample=sample),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


_JSON_OUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under NCCL_DEBUG): keep a
    private duplicate of fd 1 for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measure_rollout(dev, peak, world=1, iters=3):
    """BASELINE config C4 on one GPU: 8 sequences x 186 steps of action feedback through AutoRegressiveTransformer.sequential_inference
    (every frame encoded once, one key/value-cached decode step per position).  HBM roofline of the decode step: the decoder and head
    weights (read once per step) plus the self-attention cache rows read at step t, averaged over t."""
    import videocad_b200.model as vmodel
    from videocad_b200 import AutoRegressiveTransformer

    cfg = ROLLOUT
    B, T, S = cfg["B"], cfg["T"], cfg["S"]
    mc = cfg["model"]
    H, Ff, L = mc["hidden_size"], mc["dim_feedforward"], mc["num_decoder_layers"]
    torch.manual_seed(0)
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", **mc).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    frames = torch.randn(B, T, 1, S, S, device=dev, generator=g).clamp_(-1, 1)
    cad = torch.randn(B, 1, S, S, device=dev, generator=g).clamp_(-1, 1)
    total, decode = [], []
    for it in range(iters + 2):  # the first two calls run eagerly / capture the CUDA graphs
        vmodel._ROLLOUT_EVENTS = []
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cmds, params = m.sequential_inference(frames, cad, action=True)
        e1.record()
        torch.cuda.synchronize()
        if it > 1:
            total.append(e0.elapsed_time(e1))
            d0, d1 = vmodel._ROLLOUT_EVENTS[0]
            decode.append(d0.elapsed_time(d1))
    vmodel._ROLLOUT_EVENTS = None
    assert cmds.shape == (B, T, 5) and params.shape == (B, T, 6, 1000) and bool(torch.isfinite(params).all())
    ms = sorted(total)[len(total) // 2]
    ms_dec = sorted(decode)[len(decode) // 2]
    if world > 1:  # every rank rolls out its own 8 sequences (no communication): the job finishes with the slowest rank
        import torch.distributed as dist

        both = torch.tensor([ms, ms_dec], device=dev)
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        ms, ms_dec = both[0].item(), both[1].item()
    step_ms = ms_dec / T
    w_bytes = 4 * (L * (6 * H * H + 2 * H * Ff) + 6005 * H)            # in_proj (3), out_proj, cross q, cross out, linear1/2; heads
    kv_bytes = 4 * L * B * ((T + 1) / 2.0) * 2 * H                      # keys + values of the t+1 cached positions, mean over t
    step_bytes = w_bytes + kv_bytes
    gbs = step_bytes / (step_ms * 1e-3) / 1e9
    del m, frames, cad
    torch.cuda.empty_cache()
    return dict(workload=f"c4: {world} x ({B} sequences x {T} steps), H={H}, {S}x{S}, eval, argmax feedback (exact incremental decoding)",
                n_gpus=world, sequences=world * B, ms_per_rollout=ms, frames_per_s=world * B * T / (ms / 1e3),
                ms_encode_frames=ms - ms_dec, ms_per_decode_step=step_ms,
                roofline=dict(bound="hbm", kernel="decode step (all kernels of one position)", achieved=gbs, peak=peak["hbm_gbs"], unit="GB/s",
                              frac=gbs / peak["hbm_gbs"], algorithmic_bytes_per_step=step_bytes,
                              note="fp32 decoder + head weights read once per step plus the cached keys/values of the positions so far "
                                   "(mean over the 186 steps); the step's time includes the full-length pass that builds the cache"))


def measure_through_trainer(net, dev_batches, B, T, world, steps, accelerate=False):
    """frames/s of the UNMODIFIED reference trainer (oracle/_ref staged copy or /root/reference) driving the drop-in on the GPU:
    trainer._process_batch = zero_grad -> prepare_batch -> model -> its own compute_loss (argmax metrics with ~30 .item() syncs) ->
    backward -> torch clip_grad_norm_ -> torch Adam.  `accelerate`: the same trainer object after
    `videocad_b200.trainer_accel.accelerate_trainer` (fused loss + metrics, ClipAdam); the loss and the metrics dict are read on the
    host every step, as the reference's training loop does (trainer.py:450-451)."""
    if not reference_available():
        return dict(unavailable="reference sources not staged (python -m oracle.build_ref in the build container)")
    import torch.distributed as dist
    from oracle import ref_trainer as rt

    dev = dev_batches[0]["frames"].device
    trainer = rt.make_trainer(net, dev, lr=1e-5)
    if accelerate:
        from videocad_b200.trainer_accel import accelerate_trainer

        accelerate_trainer(trainer)

    def one(i):
        loss, metrics = trainer._process_batch(dev_batches[i % NUM_BATCHES])
        if accelerate:  # what _train_epoch does with the two results (the unpatched compute_loss has synchronised ~30 times already)
            return loss.item() + metrics["total_predictions"]

    for i in range(3):
        one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    note = ("unmodified trainer object after accelerate_trainer(): its _process_batch with the fused loss + metrics kernels and ClipAdam; "
            "loss.item() and the metrics dict read on the host every step; inputs resident in HBM") if accelerate else \
           ("unmodified trainer._process_batch (trainer.py:480-496) around the drop-in: the reference's compute_loss with its "
            ".item() syncs, torch clip_grad_norm_ and torch Adam; inputs resident in HBM")
    return dict(value=world * B * T * steps / (ms / 1e3), unit="frames/s", ms_per_step=ms / steps, steps=steps, note=note)


def measure_c3(args, rank, world, local_rank, steps=10):
    """BASELINE config C3 inside the same N-rank launch: T=32, H=Ff=1024, 32 samples per GPU, DDP gradient all-reduce (~508 MB).
    Also times the same steps under no_sync() (no all-reduce): the difference is the all-reduce time the backward could not hide."""
    import torch.distributed as dist
    from videocad_b200 import AutoRegressiveTransformer
    from videocad_b200 import loss as vloss
    from videocad_b200.optim import ClipAdam

    cfg = CONFIGS["c3"]
    B, T, S = cfg["B"], cfg["T"], cfg["S"]
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    model = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", precision=args.precision, **cfg["model"]).to(dev)
    model.train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank,
                                                    find_unused_parameters=True) if world > 1 else model
    opt = ClipAdam(model.parameters(), lr=1e-5, max_norm=1.0)
    from videocad_b200.synthetic import synthetic_batch

    batches = [{k: v.to(dev) for k, v in synthetic_batch(B, T + 1, S, seed=4321 + 1000 * rank + i).items()} for i in range(2)]

    def step(i):
        opt.zero_grad()
        b = batches[i % 2]
        inputs = {"frames": b["frames"][:, :-1], "actions": model.normalize_actions(b["actions"][:, :-1].clone()), "cad_image": b["cad_image"]}
        loss = vloss.compute_loss_fused(net(inputs), b["actions"][:, 1:])
        loss.backward()
        opt.step()

    def timed(n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n

    for i in range(3):
        step(i)
    ms = timed(steps)
    ms_nosync = None
    if world > 1:
        with net.no_sync():
            step(0)
            ms_nosync = timed(steps)
    grad_bytes = 4 * sum(p.numel() for p in model.parameters())
    f_fwd = fwd_flops_per_sample(cfg["model"], T, S)
    out = dict(workload=f"c3: {world} x (batch {B}, T={T}, {S}x{S} frames), H=1024, DDP gradient all-reduce of {grad_bytes / 1e6:.0f} MB",
               global_batch=world * B, steps=steps, ms_per_step=ms, value=world * B * T / (ms / 1e3), unit="frames/s",
               ms_per_step_without_allreduce=ms_nosync, exposed_allreduce_ms=(ms - ms_nosync) if ms_nosync is not None else None,
               grad_bytes_per_step=grad_bytes, whole_step_algorithmic_tflops_per_gpu=B * T / (ms / 1e3) * 3.0 * f_fwd / T / 1e12)
    del model, net, opt, batches
    torch.cuda.empty_cache()
    return out


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp32x3", choices=["fp32x3", "bf16"])
    ap.add_argument("--loss", default="fused", choices=["fused", "torch"],
                    help="trainer loss: the native fused kernels (default) or the torch restatement (videocad_b200/loss.py)")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="clip_grad_norm_(1.0) + Adam: the native fused step (videocad_b200.optim.ClipAdam, default) or the two "
                         "torch calls of the reference trainer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the C4 rollout leg")
    ap.add_argument("--quick", action="store_true",
                    help="A/B runs: only the `value` pass and the per-segment pass, then a short JSON line (no roofline / e2e / "
                         "trainer / rollout / CPU legs)")
    ap.add_argument("--profile-step", action="store_true",
                    help="after the warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch.distributed as dist
    from videocad_b200 import AutoRegressiveTransformer
    from videocad_b200 import lib as L
    from videocad_b200 import loss as vloss

    compute_loss = vloss.compute_loss_fused if args.loss == "fused" else vloss.compute_loss

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} needs a {args.gpus}-rank launch (torch.distributed.run --nproc-per-node {args.gpus}); "
                         f"WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    lib = L.load()

    torch.manual_seed(0)
    model = AutoRegressiveTransformer(state_dim=1644, act_dim=7, encoder="vit", precision=args.precision, **cfg["model"]).to(dev)
    model.train()
    net = model
    if world > 1:
        # the wrapper the reference builds (experiment.py:104-109)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank,
                                                        find_unused_parameters=True)
    if args.optimizer == "fused":
        from videocad_b200.optim import ClipAdam

        opt = ClipAdam(model.parameters(), lr=1e-5, max_norm=1.0)  # clip_grad_norm_(1.0) + Adam.step() in one native step
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-5)
    B, T, S = cfg["B"], cfg["T"], cfg["S"]

    def train_step(batch):
        opt.zero_grad()
        acts = batch["actions"]
        inputs = {"frames": batch["frames"][:, :-1], "actions": model.normalize_actions(acts[:, :-1].clone()),
                  "cad_image": batch["cad_image"]}
        loss = compute_loss(net(inputs), acts[:, 1:])
        loss.backward()
        if args.optimizer != "fused":
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, steps, profile_gemm=False):
        barrier()
        if profile_gemm:
            lib.vc_gemm_profile(1)
        lib.vc_launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            run_step(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), lib.vc_launch_count()

    # ---------------- value: inputs resident in HBM
    dev_batches = make_batches(B, T, S, rank, device=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up
    for i in range(args.warmup):
        train_step(dev_batches[i % NUM_BATCHES])
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        train_step(dev_batches[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            sampler.stop()
        return
    ms, launches = timed(lambda i: train_step(dev_batches[i % NUM_BATCHES]), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    frames_per_step = world * B * T
    value = frames_per_step * args.steps / (ms / 1000.0)

    # ---------------- segment pass: CUDA events around each of the six graph replays of a step (frame ViT, CAD ViT and
    # sequence transformer, forward and backward); what remains of the step is loss, clip, Adam and input staging
    import videocad_b200.model as vmodel

    segments = None
    if vmodel._GRAPHS:
        vmodel.timing_begin()
        ms_seg, _ = timed(lambda i: train_step(dev_batches[i % NUM_BATCHES]), args.steps)
        rep = vmodel.timing_report()
        segments = {k: round(v[1] / args.steps, 4) for k, v in sorted(rep.items())}
        segments["step_ms"] = round(ms_seg / args.steps, 4)
        segments["outside_graphs_ms"] = round(ms_seg / args.steps - sum(v[1] for v in rep.values()) / args.steps, 4)

    if args.quick:
        if rank == 0:
            emit(dict(metric="train frames/sec", value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                      ms_per_step=ms / args.steps, quick=True, clocks=clocks, segments_ms_per_step=segments,
                      env={k: v for k, v in os.environ.items() if k.startswith(("VC_", "VIDEOCAD_B200_"))}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline pass: the same K steps with the native segments launched kernel by kernel (CUDA-graph
    # replay off) so that a CUDA-event pair can be recorded around every GEMM launch on the launching stream
    graphs_were_on = vmodel._GRAPHS
    vmodel._GRAPHS = False
    overlap_was_on = vmodel._OVERLAP
    vmodel._OVERLAP = False          # per-launch durations need the kernels one after the other: no second encoder stream,
    lib.vc_side_streams_enable(0)    # no auxiliary streams inside the library
    train_step(dev_batches[0])
    ms_inst, launches_inst = timed(lambda i: train_step(dev_batches[i % NUM_BATCHES]), args.steps, profile_gemm=True)
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_longlong()
    lib.vc_gemm_profile_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n))
    b_ms, b_fl, b_n = C.c_double(), C.c_double(), C.c_longlong()
    lib.vc_gemm_profile_read_min(5e9, C.byref(b_ms), C.byref(b_fl), C.byref(b_n))  # the image-encoder-sized GEMMs
    if os.environ.get("VC_GEMM_DUMP"):
        lib.vc_gemm_profile_dump(os.environ["VC_GEMM_DUMP"].encode())
    lib.vc_gemm_profile(0)
    lib.vc_side_streams_enable(1)
    vmodel._OVERLAP = overlap_was_on
    vmodel._GRAPHS = graphs_were_on
    if not graphs_were_on:
        launches = launches_inst

    # ---------------- the same steps without the gradient all-reduce (DDP no_sync): what the backward could not hide
    ms_nosync = None
    if world > 1:
        with net.no_sync():
            train_step(dev_batches[0])
            ms_nosync, _ = timed(lambda i: train_step(dev_batches[i % NUM_BATCHES]), args.steps)

    # ---------------- through the unmodified reference trainer object (its own loss / metrics / clip / Adam)
    through_trainer = measure_through_trainer(net, dev_batches, B, T, world, max(5, args.steps // 2))
    try:  # an auxiliary leg: a failure here is reported in the line, it does not take the headline measurement down
        through_trainer_fused = measure_through_trainer(net, dev_batches, B, T, world, args.steps, accelerate=True)
    except Exception as e:  # noqa: BLE001
        through_trainer_fused = dict(error=f"{type(e).__name__}: {e}"[:300])

    # ---------------- e2e: host (pinned) inputs, H2D + loss read-back inside the timed region
    del dev_batches
    host_batches = make_batches(B, T, S, rank, pinned=True)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())

    # Input pipeline of the e2e arm: pinned host batches, H2D copy of batch i+1 issued on a copy stream while step i
    # computes (what a DataLoader(pin_memory=True) + non_blocking .to() gives a trainer); every step still moves its own
    # h2d bytes inside the timed region and reads the loss back.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}

    def prefetch(i):
        hb = host_batches[i % NUM_BATCHES]
        with torch.cuda.stream(copy_stream):
            batch = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending[i] = (batch, ev)

    def e2e_step(i):
        if i not in pending:
            prefetch(i)
        batch, ev = pending.pop(i)
        torch.cuda.current_stream().wait_event(ev)
        for v in batch.values():
            v.record_stream(torch.cuda.current_stream())
        prefetch(i + 1)
        return float(train_step(batch).item())

    for i in range(2):
        e2e_step(i)
    pending.clear()
    ms_e2e, _ = timed(e2e_step, args.steps)
    pending.clear()
    e2e_value = frames_per_step * args.steps / (ms_e2e / 1000.0)

    peak = measured_peaks()
    # ---------------- BASELINE's other GPU configs, measured in the same launch
    del host_batches
    pending.clear()
    torch.cuda.empty_cache()
    c3 = measure_c3(args, rank, world, local_rank) if (world > 1 and args.config == "c1") else None
    rollout = measure_rollout(dev, peak, world) if (args.config == "c1" and not args.no_rollout) else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    f_fwd = fwd_flops_per_sample(cfg["model"], T, S)
    gemm_tflops = (g_fl.value / 1e12) / (g_ms.value / 1e3) if g_ms.value > 0 else 0.0
    step_tflops = value / world * 3.0 * f_fwd / T / 1e12  # per GPU, whole step (fwd + 2x bwd) algorithmic
    passes = 3 if args.precision == "fp32x3" else 1
    big_tflops = (b_fl.value / 1e12) / (b_ms.value / 1e3) if b_ms.value > 0 else 0.0
    pair_launches = lib.vc_gemm_pair_launch_count()
    # dominant kernel: the 2-SM 256x256 GEMM (gemm_tc_pair_kernel) that runs every image-encoder GEMM of >= 5 GFLOP
    traffic, traffic_note = pair_kernel_traffic()
    roofline = dict(bound="tensor", kernel="gemm_tc_pair_kernel (tcgen05 cta_group::2, 256x256 tiles, split-bf16 x3)",
                    achieved=big_tflops, peak=peak["tflops"], unit="TFLOP/s", frac=big_tflops / peak["tflops"],
                    traffic=traffic, traffic_note=traffic_note,
                    peak_source=peak["source"], mma_passes_per_flop=passes, mma_issue_frac=passes * big_tflops / peak["tflops"],
                    launches_per_step=b_n.value / args.steps, ms_per_step=b_ms.value / args.steps,
                    share_of_step=(b_ms.value / args.steps) / (ms_inst / args.steps),
                    flops_share_of_all_gemms=b_fl.value / max(g_fl.value, 1.0),
                    algorithmic_flops="2*M*N*K per launch (SURVEY.md 8(d)); each is issued as 3 bf16 MMA passes",
                    all_gemms=dict(note="every tensor-core GEMM launch of the step, decoder-sized ones included",
                                   launches_per_step=g_n.value / args.steps, ms_per_step=g_ms.value / args.steps,
                                   achieved=gemm_tflops, frac=gemm_tflops / peak["tflops"],
                                   share_of_step=(g_ms.value / args.steps) / (ms_inst / args.steps)),
                    pair_kernel_launches_total=int(pair_launches),
                    measured_in=f"instrumented pass of the same {args.steps} steps without CUDA-graph replay and without stream overlap "
                                f"({ms_inst / args.steps:.2f} ms/step, one CUDA-event pair per GEMM launch on the launching stream); "
                                "the headline value uses graph replay",
                    whole_step_algorithmic_tflops_per_gpu=step_tflops, whole_step_frac=step_tflops / peak["tflops"])
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        if reference_available():
            from oracle.ref_trainer import time_reference_train

            cpu_baseline = time_reference_train(cfg["model"], cfg["cpu_B"], T, S, steps=3, warmup=1)
        else:
            from oracle.train_port import time_cpu_train

            cpu_baseline = time_cpu_train(cfg["model"], min(cfg["cpu_B"], 2), T, S, steps=6, warmup=1)
    line = dict(metric="train frames/sec", value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (3-pass split-bf16 tensor-core GEMMs, fp32 accumulate)" if passes == 3 else "bf16",
                data="synthetic", impl="native",
                config=workload_config(args, cfg, world), step_impl=f"{args.loss} loss, {args.optimizer} clip+Adam",
                clocks=clocks, e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                                        ms_per_step=ms_e2e / args.steps),
                gpu_launches=launches_inst, gpu_launches_note="native kernels per %d steps (%d per step); with CUDA-graph replay "
                "the same kernels run from 6 graph launches per step" % (args.steps, launches_inst // max(args.steps, 1)),
                cuda_graphs=bool(graphs_were_on), segments_ms_per_step=segments, roofline=roofline,
                cpu_baseline=cpu_baseline, through_trainer=through_trainer, through_trainer_fused=through_trainer_fused,
                rollout=rollout, c3=c3,
                ms_per_step_without_allreduce=(ms_nosync / args.steps) if ms_nosync is not None else None,
                exposed_allreduce_ms=((ms - ms_nosync) / args.steps) if ms_nosync is not None else None)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
