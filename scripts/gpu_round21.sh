#!/bin/bash
# new dropout generator + packed split conversion: full GPU test suite, bench, attention micro-benchmark
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
echo "=== attn bench"; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt | head -3
echo "=== gemm bench"; timeout 300 python scripts/gemm_bench.py --only epi,dec 2>&1 | tee gpurun_out/gemm_bench.txt
