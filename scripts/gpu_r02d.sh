#!/bin/bash
mkdir -p gpurun_out
for dbg in 8 9 12; do
  echo "== VC_GEMM_DEBUG=$dbg" >> gpurun_out/r02d_pair_timeline.txt
  VC_GEMM_DEBUG=$dbg timeout 300 python scripts/gemm_bench.py --only "epi fc2 fwd,epi fc1 fwd,epi out fwd,epi qkv fwd,epi fc2 dgrad" --iters 1 >> gpurun_out/r02d_pair_timeline.txt 2>&1
done
grep -c timeline gpurun_out/r02d_pair_timeline.txt
timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback > gpurun_out/r02d_rollout.json 2>&1
cat gpurun_out/r02d_rollout.json
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10000 --launch-count 80 --csv \
  --log-file gpurun_out/r02d_rollout_launches.csv python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback --calls 1 > gpurun_out/r02d_rollout_ncu.log 2>&1
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dec_gemv --launch-skip 600 -c 3 -f -o gpurun_out/r02d_dec_gemv \
  python scripts/rollout_bench.py --batch 8 --steps 40 --only-feedback --calls 1 > gpurun_out/r02d_ncu_dec.log 2>&1
tail -2 gpurun_out/r02d_ncu_dec.log
