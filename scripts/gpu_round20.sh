#!/bin/bash
# round 1, session 2, call 1: parity of the new decoder kernels + bench + attention micro-benchmark + ncu of the ViT attention
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
echo "=== attn bench"; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt | tail -12
echo "=== ncu attn"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:vit_attn -c 2 -f -o gpurun_out/vit_attn python scripts/attn_bench.py --once > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log; ls -la gpurun_out/
