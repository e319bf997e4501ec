#!/bin/bash
mkdir -p gpurun_out
echo "=== pair gemm tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "pair or majors" 2>&1 | tail -15
echo "rc=$?"
echo "=== attention tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "attention" 2>&1 | tail -6
echo "=== gemm bench pair"; timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_pair.txt | tail -16
echo "=== gemm bench single"; VC_GEMM_PAIR=0 timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_single.txt | tail -16
echo "=== all tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== bench nopair"; VC_GEMM_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench_np.err | tee gpurun_out/bench_c1_nopair.json | cut -c1-200
