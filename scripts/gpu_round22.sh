#!/bin/bash
# where does the pair GEMM's epilogue time go?  VC_GEMM_DEBUG experiments on the N=512 image-encoder GEMMs + ncu source capture
mkdir -p gpurun_out
for dbg in 0 1 2 4; do echo "=== VC_GEMM_DEBUG=$dbg"; VC_GEMM_DEBUG=$dbg timeout 300 python scripts/gemm_bench.py --only "vit fc1 fwd,vit qkv fwd,epi fc2 fwd,epi fc1 fwd,epi fc2 dgrad" 2>&1 | tee -a gpurun_out/gemm_debug.txt; done
echo "=== ncu"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_pair -c 1 --launch-skip 3 -f -o gpurun_out/pair_fc2 python scripts/gemm_bench.py --only "epi fc2 fwd" --iters 2 > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
