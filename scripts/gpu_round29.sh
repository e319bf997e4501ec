#!/bin/bash
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_loss.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== rollout"; timeout 600 python scripts/rollout_bench.py 2>&1 | tee gpurun_out/rollout_c4.json | tail -2
