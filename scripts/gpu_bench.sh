#!/bin/bash
mkdir -p gpurun_out
echo "=== bench native c1"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_c1.err | tee gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
echo "=== bench reference c1"; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref_c1.json
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 2500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
echo "=== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 3 -o gpurun_out/prof_gemm_r1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
