#!/bin/bash
mkdir -p gpurun_out
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -3 gpurun_out/bench.err
echo "=== bench no overlap"; VIDEOCAD_B200_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench_no.err | tee gpurun_out/bench_c1_nooverlap.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -3 gpurun_out/bench_no.err
