#!/bin/bash
mkdir -p gpurun_out
echo "--- persistent kernel, no trailing fence"; timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | tail -1 | tee gpurun_out/r02v_rollout_mega.json
echo "--- persistent kernel, acquire fence"; VC_MEGA_FENCED=1 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | tail -1 | tee gpurun_out/r02v_rollout_mega_fenced.json
echo "--- per-kernel step (weights evict-first)"; VIDEOCAD_B200_DECODE_MEGA=0 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | tail -1 | tee gpurun_out/r02v_rollout_kernels.json
echo "--- parity"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "rollout" 2>&1 | tail -3
