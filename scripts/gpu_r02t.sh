#!/bin/bash
mkdir -p gpurun_out
echo "--- lean barrier"; VC_MEGA_DEBUG=100 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | grep -v Warn | tail -12 | tee gpurun_out/r02t_mega_timeline.txt
echo "--- fenced barrier"; VC_MEGA_FENCED=1 VC_MEGA_DEBUG=100 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback 2>&1 | grep -v Warn | tail -12 | tee -a gpurun_out/r02t_mega_timeline.txt
echo "--- parity (lean)"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "rollout" 2>&1 | tail -3
