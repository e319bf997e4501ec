#!/bin/bash
# persistent decode kernel: smoke, parity tests, A/B timing against the per-kernel step, sanitizer
mkdir -p gpurun_out
echo "--- smoke (eager)"; timeout 180 python scripts/sanitize_decode.py 2>&1 | tail -4
echo "--- parity tests"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "rollout" 2>&1 | tail -8
echo "--- rollout bench: persistent kernel"; timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | tail -2 | tee gpurun_out/r02s_rollout_mega.json
echo "--- rollout bench: per-kernel step"; VIDEOCAD_B200_DECODE_MEGA=0 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | tail -2 | tee gpurun_out/r02s_rollout_kernels.json
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_decode.py > gpurun_out/r02s_sanitize_$tool.log 2>&1
  tail -3 gpurun_out/r02s_sanitize_$tool.log
done
