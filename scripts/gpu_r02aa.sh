#!/bin/bash
# round 2, late pass (one GPU): full GPU suite incl. the edge-shape and accelerate_trainer tests, then per-kernel ncu metrics
# (time, DRAM bytes, tensor-pipe activity, occupancy) of every kernel of one C1 training step and of one rollout decode step
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r02aa_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r02aa_pytest_gpu.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
echo "=== per-kernel metrics: one C1 training step"; date +%s
VIDEOCAD_B200_GRAPHS=0 timeout 480 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02aa_step_kernels.csv python bench.py --warmup 3 --profile-step > gpurun_out/r02aa_ncu_step.log 2>&1; tail -2 gpurun_out/r02aa_ncu_step.log | cut -c1-200; wc -l gpurun_out/r02aa_step_kernels.csv; date +%s
python scripts/summarize_kernel_metrics.py gpurun_out/r02aa_step_kernels.csv > gpurun_out/r02aa_step_kernels_summary.txt 2>&1; head -45 gpurun_out/r02aa_step_kernels_summary.txt
echo "=== per-kernel metrics: one rollout decode step (C4 model, 8 sequences)"
VIDEOCAD_B200_GRAPHS=0 timeout 300 ncu --metrics $M --clock-control none -k regex:dec_ --launch-skip 1980 -c 66 --csv --log-file gpurun_out/r02aa_decode_kernels.csv python scripts/rollout_bench.py --batch 8 --steps 40 --only-feedback --calls 1 > gpurun_out/r02aa_ncu_dec.log 2>&1; tail -1 gpurun_out/r02aa_ncu_dec.log | cut -c1-200; date +%s
python scripts/summarize_kernel_metrics.py gpurun_out/r02aa_decode_kernels.csv > gpurun_out/r02aa_decode_kernels_summary.txt 2>&1; cat gpurun_out/r02aa_decode_kernels_summary.txt
