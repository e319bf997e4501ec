#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== gemm bench"; timeout 200 python scripts/gemm_bench.py --iters 6 2>&1 | tee gpurun_out/gemm_bench.txt | tail -16
echo "=== gemm bench nostores"; VC_GEMM_DEBUG=1 timeout 200 python scripts/gemm_bench.py --only "qkv fwd,fc1 fwd" --iters 6 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== ncu full gemm (qkv fwd)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/prof_gemm_r1c -f python scripts/gemm_bench.py --only "qkv fwd" --iters 4 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== ncu full attention bwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:vit_attn_bwd -s 2 -c 1 -o gpurun_out/prof_attnbwd_r1c -f python -m pytest tests/test_model_gpu.py -m gpu -q -k full_size -p no:cacheprovider > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
