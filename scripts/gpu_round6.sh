#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== gemm bench"; timeout 200 python scripts/gemm_bench.py --iters 6 --only "dec,head" 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== bench graphs off"; VIDEOCAD_B200_GRAPHS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_nog.err | tee gpurun_out/bench_c1_nographs.json | cut -c1-200
