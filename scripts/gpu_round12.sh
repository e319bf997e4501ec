#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench.err
echo "=== bench torch optimizer"; timeout 600 python bench.py --steps 20 --warmup 4 --optimizer torch --no-cpu-baseline 2> gpurun_out/bench_to.err > gpurun_out/bench_c1_torchopt.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1_torchopt.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -3 gpurun_out/bench_to.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; cut -c1-400 gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
