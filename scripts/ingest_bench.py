#!/usr/bin/env python
"""Frame ingestion measured: the device transform `videocad_b200.ingest.FrameTransform`
(vc_frames_rgb_u8_ingest: Pillow-exact Resize -> Grayscale -> ToTensor -> Normalize) against its HBM roofline, next to the
reference's per-frame PIL transform (main.py:103-108, data_loader.py:434-446) on the host cores.

    python scripts/ingest_bench.py [--frames 256] [--iters 20] [--cpu-frames 64]

Cases: stored size 224 x 224 (the `data_resized` dataset: no resampling, one kernel) and 448 x 448 / 720 x 1280 (two resampling passes).
Algorithmic bytes per output frame: 3 B per INPUT pixel read + 4 B per output pixel written (the horizontal pass' uint8 intermediate is
extra traffic the roofline does not credit).  Inputs rotate over enough batches to exceed the 126 MB L2.  One JSON line per case."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measured_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6555.0, "fallback (copy bandwidth of this pool's B200s)"


def cpu_reference_fps(h, w, n):
    """frames/s of the reference's transform on PIL images, one thread (a DataLoader worker)."""
    from PIL import Image
    from torchvision import transforms

    t = transforms.Compose([transforms.Resize((224, 224)), transforms.Grayscale(1), transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
    rng = np.random.default_rng(0)
    imgs = [Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)) for _ in range(min(n, 16))]
    torch.set_num_threads(1)
    t(imgs[0])
    t0 = time.perf_counter()
    for i in range(n):
        t(imgs[i % len(imgs)])
    return n / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256, help="frames per batch (C1: 32 x 8)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-frames", type=int, default=64)
    ap.add_argument("--cpu-only", action="store_true")
    args = ap.parse_args()
    peak, peak_src = measured_hbm_gbs()
    for (h, w) in [(224, 224), (448, 448), (720, 1280)]:
        line = dict(metric="ingested frames/sec", unit="frames/s", config=dict(workload=f"{args.frames} RGB uint8 frames {h}x{w} -> fp32 [1,224,224]"))
        line["cpu_baseline"] = dict(value=cpu_reference_fps(h, w, args.cpu_frames), unit="frames/s", cores=1, kind="reference",
                                    sample=f"{args.cpu_frames} frames through torchvision Resize/Grayscale/ToTensor/Normalize on PIL images")
        if not args.cpu_only:
            from videocad_b200 import ingest

            ft = ingest.FrameTransform((224, 224))
            nbuf = max(2, int(160e6 // (args.frames * h * w * 3)) + 1)  # > 126 MB of distinct inputs
            g = torch.Generator(device="cuda").manual_seed(0)
            bufs = [torch.randint(0, 256, (args.frames, h, w, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(nbuf)]
            out = torch.empty(args.frames, 1, 224, 224, device="cuda")
            for i in range(3):
                ft(bufs[i % nbuf], out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                ft(bufs[i % nbuf], out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            algo = args.frames * (h * w * 3 + 224 * 224 * 4)
            gbs = algo / (ms * 1e-3) / 1e9
            line.update(value=args.frames / (ms * 1e-3), ms_per_batch=ms, n_buffers=nbuf,
                        roofline=dict(bound="hbm", achieved=gbs, peak=peak, unit="GB/s", frac=gbs / peak, peak_source=peak_src,
                                      algorithmic_bytes_per_batch=algo, kernel="vc_frames_rgb_u8_ingest (" +
                                      ("rgb_u8_gray_norm_kernel" if (h, w) == (224, 224) else "resample_h_kernel + resample_v_gray_norm_kernel") + ")"))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
