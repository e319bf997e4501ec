#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -6
echo "--- isolated GEMMs: tail split on / off"
timeout 300 python scripts/gemm_bench.py --only epi 2>&1 | grep -v Warn | tail -14 | tee gpurun_out/r02y_gemm_bench_tail_split.txt
VC_GEMM_TAIL_SPLIT=0 timeout 300 python scripts/gemm_bench.py --only epi 2>&1 | grep -v Warn | tail -14 | tee gpurun_out/r02y_gemm_bench_no_tail_split.txt
echo "--- step: tail split on / off"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout 2> gpurun_out/r02y_bench.err > gpurun_out/r02y_bench_c1.json
python -c "import json; d=json.load(open('gpurun_out/r02y_bench_c1.json')); print('split', d['value'], d['ms_per_step'], d['segments_ms_per_step'], d['roofline']['frac'])"
VC_GEMM_TAIL_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout 2> gpurun_out/r02y_bench0.err > gpurun_out/r02y_bench_c1_nosplit.json
python -c "import json; d=json.load(open('gpurun_out/r02y_bench_c1_nosplit.json')); print('nosplit', d['value'], d['ms_per_step'], d['segments_ms_per_step'], d['roofline']['frac'])"
