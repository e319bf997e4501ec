#!/bin/bash
# ncu --set full of ONE launch of the dominant kernel (2-SM GEMM, to_qkv forward shape) -> gpurun_out/pair_qkv.ncu-rep
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_pair -c 1 --launch-skip 3 -f -o gpurun_out/pair_qkv python scripts/gemm_bench.py --only "epi qkv fwd" --iters 2 > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
