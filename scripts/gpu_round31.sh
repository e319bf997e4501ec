#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "gemm and not pair" 2>&1 | tail -3
echo "=== gemm bench dec"; timeout 300 python scripts/gemm_bench.py --only "dec,head" 2>&1 | tee gpurun_out/gemm_bench_dec.txt
for i in 1 2; do
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -2 gpurun_out/bench.err | cut -c1-200
done
