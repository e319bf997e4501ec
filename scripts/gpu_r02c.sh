#!/bin/bash
# round 2, third GPU pass: GPU suite, rollout timing + launch list, ncu --set full of the fc2-forward pair GEMM and of one decode GEMV
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -rs > gpurun_out/r02c_pytest.txt 2>&1
tail -8 gpurun_out/r02c_pytest.txt
timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback > gpurun_out/r02c_rollout.json 2>&1
cat gpurun_out/r02c_rollout.json
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10000 --launch-count 140 --csv \
  --log-file gpurun_out/r02c_rollout_launches.csv python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback --calls 1 > gpurun_out/r02c_rollout_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_pair -c 1 --launch-skip 3 -f -o gpurun_out/r02c_pair_fc2 \
  python scripts/gemm_bench.py --only "epi fc2 fwd" --iters 2 > gpurun_out/r02c_ncu_pair.log 2>&1
tail -2 gpurun_out/r02c_ncu_pair.log
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dec_gemv --launch-skip 3000 -c 2 -f -o gpurun_out/r02c_dec_gemv \
  python scripts/rollout_bench.py --batch 8 --steps 60 --only-feedback --calls 1 > gpurun_out/r02c_ncu_dec.log 2>&1
tail -2 gpurun_out/r02c_ncu_dec.log
timeout 300 python scripts/gemm_bench.py --only "epi" > gpurun_out/r02c_gemm_bench.txt 2>&1
cat gpurun_out/r02c_gemm_bench.txt
