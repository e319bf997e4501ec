#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests (all)"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== rollout"; timeout 600 python scripts/rollout_bench.py 2>&1 | tee gpurun_out/rollout_c4.json | tail -2
echo "=== bench c3"; timeout 600 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c3.err > gpurun_out/bench_c3.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c3.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_c3.err
