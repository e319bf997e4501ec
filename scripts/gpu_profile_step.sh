#!/bin/bash
# launch list of exactly one steady-state training step (CUDA graphs off so that every kernel is listed)
mkdir -p gpurun_out
VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python bench.py --warmup 3 --profile-step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log | cut -c1-200; wc -l gpurun_out/launches_step.csv
