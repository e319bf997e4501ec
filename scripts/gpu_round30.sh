#!/bin/bash
mkdir -p gpurun_out
for pr in 0 -1 0 -1; do
echo "=== bench side priority $pr"; VIDEOCAD_B200_SIDE_PRIORITY=$pr timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_pr$pr.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_pr$pr.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -2 gpurun_out/bench.err | cut -c1-200
done
