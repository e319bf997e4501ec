#!/bin/bash
# 256 x 128 tiles for the 2-SM GEMM: parity of both widths, micro-benchmark auto vs forced 256, bench
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k gemm 2>&1 | tail -5
echo "=== gemm bench auto"; timeout 300 python scripts/gemm_bench.py --only "vit,epi" 2>&1 | tee gpurun_out/gemm_bench_auto.txt
echo "=== gemm bench bn256"; VC_GEMM_PAIR_BN=256 timeout 300 python scripts/gemm_bench.py --only "vit,epi" 2>&1 | tee gpurun_out/gemm_bench_bn256.txt
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
