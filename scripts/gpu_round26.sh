#!/bin/bash
mkdir -p gpurun_out
echo "=== ln tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "layernorm or patch" 2>&1 | tail -3
echo "=== attn/ln bench"; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt | head -4
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
