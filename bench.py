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
  cpu_baseline  the oracle's CPU port of the same training step on the host cores (bounded sample)

--impl reference times the CPU port only (rank 0), printing the same JSON line with "impl": "reference".
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
               B=32, T=8, S=224, cpu_B=2),
    # BASELINE.json configs[3]: DDP, 32-frame context, H=1024, 32 samples per GPU
    "c3": dict(model=dict(hidden_size=1024, nhead=4, num_decoder_layers=8, dim_feedforward=1024, window_size=10,
                          enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
               B=32, T=32, S=224, cpu_B=1),
}
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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE gemm_tc_pair_kernel launch (to_qkv forward, M=12800 N=3072 K=512), from the
# `ncu --set full` capture summarised in profiles/ (see profiles/README.md); None until a capture exists
PAIR_KERNEL_DRAM_BYTES_PER_LAUNCH = 137445888
PAIR_KERNEL_TRAFFIC_NOTE = ("profiles/r01m_ncu_gemm_pair_qkv_fwd_summary.txt: 32.9 MB read (= the split-bf16 operands, read once) + 104.5 MB written "
                            "of the 157.3 MB fp32 output (the rest was still in L2 when the kernel ended); algorithmic bytes 189.8 MB")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, source="fallback (B200_PROFILING.md sustained figure)")


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


def run_reference(args, cfg):
    """The reference arm: the CPU implementation of the path (oracle port) on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.train_port import CpuTrainStep, pick_threads
    from oracle import torch_oracle as to

    B, T, S = cfg["cpu_B"], cfg["T"], cfg["S"]
    runner = CpuTrainStep(cfg["model"])
    batches = [to.synthetic_batch(B, T + 1, S, seed=100 + i) for i in range(2)]
    runner.step(batches[0])
    pick_threads(runner, batches[1])  # all the host threads that help (torchrun would pin OMP_NUM_THREADS=1)
    for i in range(args.warmup):
        runner.step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        runner.step(batches[i % 2])
    dt = time.perf_counter() - t0
    fps = B * T * args.steps / dt
    line = dict(metric="train frames/sec", value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"{args.config}: CPU port of trainer._process_batch, batch {B} x T={T} x {S}x{S}, "
                            f"H={cfg['model']['hidden_size']} (bounded sample of the {cfg['B']}-sample GPU batch)"),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                                  sample=f"{args.steps} steps at batch {B}, host_cpus={os.cpu_count()}"),
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak = measured_peaks()
    f_fwd = fwd_flops_per_sample(cfg["model"], T, S)
    gemm_tflops = (g_fl.value / 1e12) / (g_ms.value / 1e3) if g_ms.value > 0 else 0.0
    step_tflops = value / world * 3.0 * f_fwd / T / 1e12  # per GPU, whole step (fwd + 2x bwd) algorithmic
    passes = 3 if args.precision == "fp32x3" else 1
    big_tflops = (b_fl.value / 1e12) / (b_ms.value / 1e3) if b_ms.value > 0 else 0.0
    pair_launches = lib.vc_gemm_pair_launch_count()
    # dominant kernel: the 2-SM 256x256 GEMM (gemm_tc_pair_kernel) that runs every image-encoder GEMM of >= 5 GFLOP
    roofline = dict(bound="tensor", kernel="gemm_tc_pair_kernel (tcgen05 cta_group::2, 256x256 tiles, split-bf16 x3)",
                    achieved=big_tflops, peak=peak["tflops"], unit="TFLOP/s", frac=big_tflops / peak["tflops"],
                    traffic=PAIR_KERNEL_DRAM_BYTES_PER_LAUNCH, traffic_note=PAIR_KERNEL_TRAFFIC_NOTE,
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
        from oracle.train_port import time_cpu_train

        cpu_baseline = time_cpu_train(cfg["model"], cfg["cpu_B"], T, S, steps=6, warmup=1)
    line = dict(metric="train frames/sec", value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (3-pass split-bf16 tensor-core GEMMs, fp32 accumulate)" if passes == 3 else "bf16",
                data="synthetic", impl="native",
                config=dict(workload=f"{args.config}: {world} x (batch {B}, T={T}, {S}x{S} frames), H={cfg['model']['hidden_size']}, "
                            f"8 decoder layers, window 10, dropout 0.1, Adam lr 1e-5, clip 1.0, {args.loss} loss, {args.optimizer} clip+Adam",
                            global_batch=world * B, parallelism=f"dp{world}" if world > 1 else "single",
                            l2="inputs rotate over 4 distinct batches (232 MB of frames at c1 > 126 MB L2); activations (>5 GB/step) stream through HBM",
                            fwd_gflop_per_sample=f_fwd / 1e9),
                clocks=clocks, e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                                        ms_per_step=ms_e2e / args.steps),
                gpu_launches=launches_inst, gpu_launches_note="native kernels per %d steps (%d per step); with CUDA-graph replay "
                "the same kernels run from 6 graph launches per step" % (args.steps, launches_inst // max(args.steps, 1)),
                cuda_graphs=bool(graphs_were_on), segments_ms_per_step=segments, roofline=roofline,
                cpu_baseline=cpu_baseline)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
