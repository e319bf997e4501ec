#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== ncu launch list (graphs off)"; VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1300 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-120; wc -l gpurun_out/launches_r1e.csv
echo "=== ncu full: wgrad gemm + attn bwd"; VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vit_attn_bwd|layernorm_bwd" -s 20 -c 3 -o gpurun_out/prof_r1e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
