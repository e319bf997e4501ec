#!/usr/bin/env python
"""Micro-benchmark of the attention kernels on the shapes a C1 training step issues (CUDA events, L2 flushed between
iterations): the image-encoder attention (256 frames x 16 heads x 50 tokens x 64, split-bf16 operands) and the decoder
attention (32 samples x 4 heads x T=8 x 128, causal / banded), forward and backward, with the achieved HBM bandwidth
against the algorithmic bytes (inputs read once, outputs written once).

    python scripts/attn_bench.py [--iters 10] [--once]     (--once: one launch of each kernel, for an ncu capture)
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kwrap as K  # noqa: E402
from videocad_b200 import lib as L  # noqa: E402


def timeit(fn, iters, flush):
    ts = []
    for it in range(iters + 2):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400_000)  # keep the GPU busy while the host enqueues: the events then bracket device time only
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1000.0  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()
    lib = L.load()
    flush = None if args.once else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    iters = 0 if args.once else args.iters
    g = torch.Generator(device="cuda").manual_seed(1)

    # ---- image-encoder attention: qkv as written by the to_qkv GEMM epilogue (split-bf16 [M, 3072])
    F, n, nh, d = args.frames, 50, 16, 64
    H = nh * d
    qkv = torch.randn(F * n, 3 * H, device="cuda", generator=g) * 0.5
    hi, lo = L.split(qkv)
    drop = L.make_drop(0.1, 3, 1234)
    a = K.attn_desc_split((hi[:, :H], lo[:, :H]), (hi[:, H:2 * H], lo[:, H:2 * H]), (hi[:, 2 * H:], lo[:, 2 * H:]), F, n, nh, d, drop=drop)
    o, lse = K.attention_fwd(a, F, n, nh, d)
    dsplit = L.split(torch.randn(F * n, H, device="cuda", generator=g))
    us = timeit(lambda: K.attention_fwd(a, F, n, nh, d), iters, flush) if not args.once else 0.0
    by = F * n * (3 * H * 4 + H * 4)
    print(f"vit attn fwd   F={F}: {us:8.1f} us  {by / 1e6:7.1f} MB algorithmic  {by / max(us, 1e-9) / 1e6:8.2f} TB/s", flush=True)
    us = timeit(lambda: K.attention_bwd_split(a, o, lse, dsplit, F, n, nh, d), iters, flush) if not args.once else 0.0
    by = F * n * (3 * H * 4 + H * 4 + 3 * H * 4)
    print(f"vit attn bwd   F={F}: {us:8.1f} us  {by / 1e6:7.1f} MB algorithmic  {by / max(us, 1e-9) / 1e6:8.2f} TB/s", flush=True)
    if args.once:
        K.attention_bwd_split(a, o, lse, dsplit, F, n, nh, d)
        torch.cuda.synchronize()

    # ---- image-encoder LayerNorm forward / backward (with the fused dropout-mask + split second output), [12800, 512]
    rows, Cc = F * n, 512
    x = torch.randn(rows, Cc, device="cuda", generator=g)
    gamma, beta = torch.randn(Cc, device="cuda", generator=g), torch.randn(Cc, device="cuda", generator=g)
    dy, dres = torch.randn(rows, Cc, device="cuda", generator=g), torch.randn(rows, Cc, device="cuda", generator=g)
    y, ys, mean, rstd = K.layernorm_fwd(x, gamma, beta, want_f32=False)
    us = timeit(lambda: K.layernorm_fwd(x, gamma, beta, want_f32=False), iters, flush) if not args.once else 0.0
    by = rows * Cc * (4 + 4)
    print(f"ln fwd  [{rows},{Cc}] -> split: {us:8.1f} us  {by / 1e6:7.1f} MB algorithmic  {by / max(us, 1e-9) / 1e6:8.2f} TB/s", flush=True)
    K.layernorm_bwd_fused(dy, x, mean, rstd, gamma, dres, drop)
    us = timeit(lambda: K.layernorm_bwd_fused(dy, x, mean, rstd, gamma, dres, drop), iters, flush) if not args.once else 0.0
    by = rows * Cc * (4 * 3 + 4 + 4)
    print(f"ln bwd fused [{rows},{Cc}]:       {us:8.1f} us  {by / 1e6:7.1f} MB algorithmic  {by / max(us, 1e-9) / 1e6:8.2f} TB/s", flush=True)

    # ---- decoder attention (fp32 q/k/v from the in_proj GEMM), short-sequence kernels vs the generic ones
    for (B, T, nh2, d2, mask, window) in [(32, 8, 4, 128, L.MASK_CAUSAL, 1), (32, 8, 4, 128, L.MASK_WINDOW, 10),
                                         (32, 32, 4, 256, L.MASK_WINDOW, 10)]:
        H2 = nh2 * d2
        x = torch.randn(B * T, 3 * H2, device="cuda", generator=g)
        q, k, v = x[:, :H2], x[:, H2:2 * H2], x[:, 2 * H2:]
        ad = K.attn_desc(q, k, v, B, T, T, nh2, d2, mask=mask, window=window, drop=L.make_drop(0.1, 4, 99))
        dout = torch.randn(B * T, H2, device="cuda", generator=g)
        for small in (1, 0):
            lib.vc_attention_small_enable(small)
            o2, lse2 = K.attention_fwd(ad, B, T, nh2, d2)
            uf = timeit(lambda: K.attention_fwd(ad, B, T, nh2, d2), iters, flush) if not args.once else 0.0
            ub = timeit(lambda: K.attention_bwd_split_bias(ad, o2, lse2, dout, B, T, nh2, d2), iters, flush) if not args.once else 0.0
            print(f"dec attn B={B} T={T} d={d2} mask={mask} {'short-seq' if small else 'generic  '}: fwd {uf:7.1f} us   bwd(+split+bias) {ub:7.1f} us",
                  flush=True)
        lib.vc_attention_small_enable(1)


if __name__ == "__main__":
    main()
