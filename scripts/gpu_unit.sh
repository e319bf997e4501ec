#!/bin/bash
# Runs the GPU kernel unit tests group by group, each under its own timeout, so that one hanging kernel
# cannot take the whole call (or the box) with it.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "library or split" "gemm_majors" "gemm_single or gemm_epilogue or gemm_rowadd or gemm_splitk or gemm_strided" "layernorm or patch" "assemble or act_dropout" "attention"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider 2>&1 | tail -40 | tee "gpurun_out/unit_${name}.log"
  echo "exit: ${PIPESTATUS[0]}"
done
