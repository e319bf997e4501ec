#!/bin/bash
mkdir -p gpurun_out
for ew in 16 8; do
  VC_GEMM_PAIR_EW=$ew timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -2
  echo "== EW=$ew" >> gpurun_out/r02g_gemm_bench.txt
  VC_GEMM_PAIR_EW=$ew timeout 300 python scripts/gemm_bench.py --only "epi,vit" >> gpurun_out/r02g_gemm_bench.txt 2>&1
  VC_GEMM_PAIR_EW=$ew timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r02g_bench_c1_ew$ew.json 2> gpurun_out/r02g_bench_c1_ew$ew.err
  python -c "import json; d=json.load(open('gpurun_out/r02g_bench_c1_ew$ew.json')); print($ew, d['value'], d['ms_per_step'], d['segments_ms_per_step'], d['roofline']['frac'])"
done
cat gpurun_out/r02g_gemm_bench.txt
VC_GEMM_PAIR_EW=8 VC_GEMM_DEBUG=8 timeout 300 python scripts/gemm_bench.py --only "epi fc2 fwd,epi out fwd,epi fc2 dgrad,epi qkv fwd" --iters 1 > gpurun_out/r02g_pair_timeline_ew8.txt 2>&1
