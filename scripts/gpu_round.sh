#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== gemm bench"; timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt | tail -20
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-1500; tail -3 gpurun_out/bench.err
