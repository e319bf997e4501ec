#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests EW16"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "gemm" 2>&1 | tail -4
echo "=== gemm tests EW8"; VC_GEMM_PAIR_EW=8 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "pair" 2>&1 | tail -3
echo "=== gemm bench EW16"; timeout 300 python scripts/gemm_bench.py --only "vit,epi,patch" 2>&1 | tee gpurun_out/gemm_bench_ew16.txt | tail -16
echo "=== gemm bench EW8"; VC_GEMM_PAIR_EW=8 timeout 300 python scripts/gemm_bench.py --only "vit,epi,patch" 2>&1 | tee gpurun_out/gemm_bench_ew8.txt | tail -16
echo "=== bench EW16"; VC_GEMM_DUMP=gpurun_out/gemm_dump.csv timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'], d['roofline']['mma_issue_frac'], d['roofline']['ms_per_step'])"; tail -3 gpurun_out/bench.err
echo "=== bench EW8"; VC_GEMM_PAIR_EW=8 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench8.err > gpurun_out/bench_c1_ew8.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1_ew8.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench8.err
