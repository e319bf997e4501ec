#!/usr/bin/env python
"""On-disk sequence format measured on the host: the reference's per-sample pickles read by the unmodified
`BaseSequenceRetriever.get_sequence` (data_loader/sequence_retriver.py:29-33) against `MmapSequenceRetriever` over one
memory-mapped store, on the same synthetic dataset (page cache warm: what is measured is the deserialisation and copying each
format forces, not the disk).

    python scripts/sequence_store_bench.py [--samples 12] [--frames 186] [--size 224]

Per access: `retrieve` = get_sequence(idx) alone; `retrieve + collate` = get_sequence followed by the copy of the frames into a
batch buffer (what a collate function does with them).  One JSON line."""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=12)
    ap.add_argument("--frames", type=int, default=186)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    from oracle import reference_model as rm
    from videocad_b200.sequence_store import MmapSequenceRetriever, convert_pickles

    rm._prepare_path()
    sys.modules.pop("data_loader", None)
    from data_loader.sequence_retriver import BaseSequenceRetriever  # type: ignore  (the reference's own class)

    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory(prefix="vc_store_bench_") as d:
        data_files, image_files = [], []
        for i in range(args.samples):
            sid = f"{i:08d}"
            frames = rng.integers(0, 256, size=(args.frames, args.size, args.size, 3), dtype=np.uint8)
            actions = rng.integers(0, 1000, size=(args.frames, 7)).astype(np.float64)
            path = os.path.join(d, f"{sid}_data.pkl")
            with open(path, "wb") as f:  # generate_dataset.py:194-199
                pickle.dump({"frames": frames, "actions": actions, "timesteps": np.arange(args.frames)}, f)
            data_files.append(path)
            image_files.append(os.path.join(d, f"{sid}_frame.png"))
        store = os.path.join(d, "ds.vcseq")
        t0 = time.perf_counter()
        convert_pickles(data_files, store)
        convert_s = time.perf_counter() - t0
        ref, got = BaseSequenceRetriever(data_files, image_files), MmapSequenceRetriever(data_files, image_files, store)
        batch = np.empty((args.frames, args.size, args.size, 3), dtype=np.uint8)
        out = {}
        for name, r in (("pickle (reference)", ref), ("memory-mapped store", got)):
            for i in range(args.samples):  # warm the page cache and check the values once
                f, a, _ = r.get_sequence(i)
                assert f.shape == batch.shape
            res = {}
            for mode in ("retrieve", "retrieve + collate"):
                t0 = time.perf_counter()
                for _ in range(args.rounds):
                    for i in range(args.samples):
                        f, a, _ = r.get_sequence(i)
                        if mode != "retrieve":
                            np.copyto(batch, f)
                res[mode] = (time.perf_counter() - t0) / (args.rounds * args.samples) * 1e3
            out[name] = res
        rf, gf = ref.get_sequence(3)[0], got.get_sequence(3)[0]
        assert np.array_equal(rf, gf)
        mb = batch.nbytes / 1e6
        print(json.dumps(dict(metric="ms per sample access", config=dict(workload=f"{args.samples} samples x {args.frames} frames x {args.size}x{args.size}x3 uint8 "
                              f"({mb:.1f} MB of frames per sample), page cache warm, one host thread"),
                              ms_per_access=out, convert_seconds=convert_s, store_bytes=os.path.getsize(store),
                              pickle_bytes=sum(os.path.getsize(p) for p in data_files))))


if __name__ == "__main__":
    main()
