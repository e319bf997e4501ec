#!/bin/bash
# ncu --set full (with source) of the heaviest kernels of one C1 training step, captured INSIDE the step (launch indices from
# profiles/r02aa_step_kernels_ncu.csv): 63-67 = to_qkv forward pair GEMM, image-encoder attention forward, out-proj forward pair GEMM
# (bias + dropout + residual), LayerNorm forward, fc1 forward pair GEMM (GELU + dropout); 390-396 = fc2 dgrad pair GEMM (GELU' +
# dropout + column sums), wgrad pair GEMMs, fused LayerNorm backward, image-encoder attention backward; 220-223 = decoder GEMMs
mkdir -p gpurun_out
for R in "63 5 fwd" "390 7 bwd" "220 4 dec"; do
  set -- $R
  VIDEOCAD_B200_GRAPHS=0 timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none --launch-skip $1 -c $2 -f -o gpurun_out/r02ab_step_$3 python bench.py --warmup 3 --profile-step > gpurun_out/r02ab_ncu_$3.log 2>&1
  tail -1 gpurun_out/r02ab_ncu_$3.log | cut -c1-160; ls -la gpurun_out/r02ab_step_$3.ncu-rep; date +%s
done
