#!/usr/bin/env python
"""Micro-benchmark of the tensor-core GEMM on the shapes the C1/C3 training step issues (CUDA events, L2 flushed
between iterations).  Usage: python scripts/gemm_bench.py [--passes 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videocad_b200 import lib as L  # noqa: E402

SHAPES = [
    # name, M, N, K, a_mn, b_mn, splitk, out
    ("vit qkv fwd", 12800, 3072, 512, 0, 0, 1, "f32"),
    ("vit out fwd", 12800, 512, 1024, 0, 0, 1, "f32"),
    ("vit fc1 fwd", 12800, 512, 512, 0, 0, 1, "split"),
    ("vit qkv dgrad", 12800, 512, 3072, 0, 1, 1, "f32"),
    ("vit out dgrad", 12800, 1024, 512, 0, 1, 1, "f32"),
    ("vit qkv wgrad", 3072, 512, 12800, 1, 1, 6, "f32"),
    ("vit fc wgrad", 512, 512, 12800, 1, 1, 9, "f32"),
    ("vit out wgrad", 512, 1024, 12800, 1, 1, 5, "f32"),
    ("patch fwd", 12544, 512, 1024, 0, 0, 1, "f32"),
    ("dec in_proj fwd", 256, 1536, 512, 0, 0, 1, "f32"),
    ("dec ffn fwd", 256, 512, 512, 0, 0, 1, "split"),
    ("head fwd", 256, 6000, 512, 0, 0, 1, "f32"),
    ("head wgrad", 6000, 512, 256, 1, 1, 1, "f32"),
    ("c3 qkv fwd", 51200, 3072, 512, 0, 0, 1, "f32"),
    # the same image-encoder shapes with the epilogues the training step actually uses
    ("epi qkv fwd split", 12800, 3072, 512, 0, 0, 1, "split"),
    ("epi out fwd bias+drop+res", 12800, 512, 1024, 0, 0, 1, "drop_res"),
    ("epi fc2 fwd bias+drop+res", 12800, 512, 512, 0, 0, 1, "drop_res"),
    ("epi fc1 fwd gelu+drop", 12800, 512, 512, 0, 0, 1, "act"),
    ("epi fc2 dgrad gelu'+drop+colsum", 12800, 512, 512, 0, 1, 1, "bwd"),
    ("epi out dgrad split", 12800, 1024, 512, 0, 1, 1, "split"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="comma-separated substrings selecting shapes")
    ap.add_argument("--no-flush", action="store_true", help="leave the operands in L2 between iterations (warm-cache timing)")
    args = ap.parse_args()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    only = [t for t in args.only.split(",") if t]
    for name, M, N, K, a_mn, b_mn, sk, out in SHAPES:
        if only and not any(t in name for t in only):
            continue
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        a, b = L.split(A), L.split(B)
        o32 = torch.zeros(M, N, device="cuda") if out in ("f32", "drop_res") else None
        osp = (torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), torch.empty(M, N, dtype=torch.bfloat16, device="cuda")) \
            if out in ("split", "act", "bwd") else None
        kw = {}
        if out == "drop_res":
            kw = dict(bias=torch.randn(N, device="cuda"), drop=L.make_drop(0.1, 5, 123), residual=torch.randn(M, N, device="cuda"))
        elif out == "act":
            kw = dict(bias=torch.randn(N, device="cuda"), drop=L.make_drop(0.1, 5, 123), act=L.ACT_GELU,
                      preact=torch.empty(M, N, device="cuda"))
        elif out == "bwd":
            kw = dict(drop=L.make_drop(0.1, 5, 123), act=L.ACT_GELU, act_backward=True, act_aux=torch.randn(M, N, device="cuda"),
                      colsum=torch.zeros(N, device="cuda"))
        ts = []
        for it in range(args.iters + 2):
            if not args.no_flush:
                flush.zero_()
            torch.cuda._sleep(200_000)  # the host enqueues the launch while the GPU is busy: events bracket device time only
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.gemm(a, b, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), passes=args.passes, splitk=sk, out_f32=o32, out_split=osp, **kw)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"{name:32s} M{M:6d} N{N:5d} K{K:6d} a{a_mn} b{b_mn} s{sk}: {ms * 1000:8.1f} us  {tf:7.1f} TF/s alg  ({tf * args.passes:7.1f} MMA)", flush=True)


if __name__ == "__main__":
    main()
