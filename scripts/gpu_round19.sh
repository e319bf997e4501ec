#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "layernorm or patch" 2>&1 | tail -4
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
echo "=== launch list"; VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step3.csv python bench.py --warmup 3 --profile-step > gpurun_out/ncu_step3.log 2>&1; tail -1 gpurun_out/ncu_step3.log | cut -c1-200; python scripts/summarize_launches.py gpurun_out/launches_step3.csv | grep -i "ln_\|total"
