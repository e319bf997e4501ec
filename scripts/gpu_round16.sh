#!/bin/bash
mkdir -p gpurun_out
echo "=== ncu full: pair kernel, fc2 fwd drop+res"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 2 -c 1 -o gpurun_out/prof_pair_fc2 -f python scripts/gemm_bench.py --only "epi fc2 fwd" --iters 2 > gpurun_out/ncu_pair2.log 2>&1; tail -2 gpurun_out/ncu_pair2.log | cut -c1-200
