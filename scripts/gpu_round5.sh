#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== gemm bench"; timeout 200 python scripts/gemm_bench.py --iters 6 --only "fc1 fwd,dec,head" 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== ncu launch list (graphs off)"; VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1300 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-120; wc -l gpurun_out/launches_r1d.csv
