#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or attention" 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | head -2 > gpurun_out/r02i_attn_bench_pipe.txt
VC_VIT_ATTN_PIPE=0 timeout 300 python scripts/attn_bench.py 2>&1 | head -2 > gpurun_out/r02i_attn_bench_old.txt
cat gpurun_out/r02i_attn_bench_pipe.txt gpurun_out/r02i_attn_bench_old.txt
