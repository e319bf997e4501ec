#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for dbg in 0 1 2 4; do
  echo "=== gemm bench VC_GEMM_DEBUG=$dbg"; VC_GEMM_DEBUG=$dbg timeout 200 python scripts/gemm_bench.py --only "qkv fwd,fc1 fwd,qkv dgrad,out fwd,fc wgrad" --iters 6 2>&1 | tee gpurun_out/gemm_bench_dbg$dbg.txt | tail -8
done
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-300; tail -3 gpurun_out/bench.err
echo "=== ncu launch list (graphs off)"; VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1500 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_r1b.csv
echo "=== ncu full gemm (qkv fwd)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_r1b -f python scripts/gemm_bench.py --only "qkv fwd" --iters 4 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
