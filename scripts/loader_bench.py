#!/usr/bin/env python
"""Host side of the data path measured: batches per second a single loader thread produces

  * reference: the unmodified `DatasetBase.__getitem__` (unpickle the sample, PIL Resize / Grayscale / ToTensor / Normalize per frame,
    cv2 for the CAD image; data_loader.py:434-508 with main.py:103-110's transforms) + `collate_with_padding` -> fp32 batch;
  * here: `RawSequenceDataset` over the memory-mapped store + `collate_u8` -> uint8 batch (the transform then runs on the GPU:
    `profiles/r02ag_ingest_bench.jsonl`, 21.5 us per 256 frames).

    python scripts/loader_bench.py [--samples 32] [--frames 9] [--size 224]

Same synthetic dataset on disk for both, page cache warm, one thread (one DataLoader worker).  One JSON line."""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=32, help="samples per batch (C1: 32)")
    ap.add_argument("--frames", type=int, default=9, help="stored frames per sample (C1: T + 1 = 9)")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import cv2
    from torchvision import transforms

    from oracle import reference_model as rm
    from videocad_b200 import ingest
    from videocad_b200.sequence_store import MmapSequenceRetriever, convert_dataset_dir

    rm._prepare_path()
    for mod in [m for m in sys.modules if m == "data_loader" or m.startswith("data_loader.")]:
        sys.modules.pop(mod)
    from data_loader.data_loader import DatasetBase  # type: ignore  (the reference's module)

    torch.set_num_threads(1)
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory(prefix="vc_loader_bench_") as root:
        for i in range(args.samples):
            sid = f"{i + 3:08d}"
            d = os.path.join(root, sid[:4])
            os.makedirs(d, exist_ok=True)
            frames = rng.integers(0, 256, size=(args.frames, args.size, args.size, 3), dtype=np.uint8)
            actions = np.full((args.frames, 7), -1.0)
            actions[:, 0] = rng.integers(0, 5, size=args.frames)
            with open(os.path.join(d, f"{sid}_data.pkl"), "wb") as f:
                pickle.dump({"frames": frames, "actions": actions, "timesteps": np.arange(args.frames)}, f)
            cv2.imwrite(os.path.join(d, f"{sid}_frame.png"), rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8))
        frame_t = transforms.Compose([transforms.Resize((224, 224)), transforms.Grayscale(1), transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
        ds = DatasetBase(root, frame_transform=frame_t, image_transform=transforms.Normalize(mean=[0.5], std=[0.5]), image_size=(224, 224),
                         image_dir=root)
        store = os.path.join(root, "ds.vcseq")
        convert_dataset_dir(root, store)
        raw = ingest.RawSequenceDataset(MmapSequenceRetriever(ds.data_files, ds.image_files, store), ds.image_loader, (224, 224))

        def ref_batch():
            return ds.collate_with_padding([ds[i] for i in range(len(ds))])

        def raw_batch():
            return ingest.collate_u8([raw[i] for i in range(len(raw))], pin=False)

        out = {}
        for name, fn in (("reference DatasetBase + collate_with_padding (fp32 batch)", ref_batch), ("RawSequenceDataset + collate_u8 (uint8 batch)", raw_batch)):
            b = fn()  # warm
            t0 = time.perf_counter()
            for _ in range(args.rounds):
                b = fn()
            dt = (time.perf_counter() - t0) / args.rounds
            nbytes = sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))
            out[name] = dict(ms_per_batch=dt * 1e3, frames_per_s=args.samples * args.frames / dt, batch_bytes=nbytes)
        print(json.dumps(dict(metric="loader frames/sec (one host thread)", config=dict(workload=f"batch of {args.samples} samples x {args.frames} frames "
                              f"{args.size}x{args.size}x3 + one CAD image each, page cache warm"), results=out)))


if __name__ == "__main__":
    main()
