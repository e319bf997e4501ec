#!/bin/bash
mkdir -p gpurun_out
VC_GEMM_DUMP=gpurun_out/gemm_dump.csv timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dump.json 2>gpurun_out/bench_dump.err; tail -3 gpurun_out/bench_dump.err; wc -l gpurun_out/gemm_dump.csv
