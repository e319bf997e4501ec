#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python scripts/gemm_bench.py --only "epi,vit" > gpurun_out/r02l_gemm_bench.txt 2>&1
cat gpurun_out/r02l_gemm_bench.txt
VC_GEMM_DEBUG=8 timeout 300 python scripts/gemm_bench.py --only "epi fc2 fwd,epi fc1 fwd,epi qkv fwd,epi fc2 dgrad" --iters 1 > gpurun_out/r02l_pair_timeline.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r02l_bench_c1.json 2> gpurun_out/r02l_bench_c1.err
python -c "import json; d=json.load(open('gpurun_out/r02l_bench_c1.json')); print(d['value'], d['ms_per_step'], d['segments_ms_per_step'], d['roofline']['frac'])"
