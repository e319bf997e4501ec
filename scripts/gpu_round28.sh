#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests (all)"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== rollout"; timeout 600 python scripts/rollout_bench.py 2>&1 | tee gpurun_out/rollout_c4.json | tail -2
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
