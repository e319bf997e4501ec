#!/bin/bash
# round 2, first GPU pass: full GPU test suite, C1 bench line (with the new legs), launch list of the rollout's decode steps
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_env.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.txt 2>&1
tail -5 gpurun_out/r02a_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_c1.json 2> gpurun_out/r02a_bench_c1.err
tail -c 600 gpurun_out/r02a_bench_c1.json
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 --launch-count 300 --csv \
  --log-file gpurun_out/r02a_rollout_launches.csv python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback --calls 1 > gpurun_out/r02a_rollout_ncu.log 2>&1
tail -3 gpurun_out/r02a_rollout_ncu.log
