#!/bin/bash
# follow-up of r02ac: the derivative-store kernel test on the 2-SM shape with its full report, the model suite with the split-K heads
# dgrad, and the step with / without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "derivative_store" 2>&1 | grep -v Warning | tail -40
echo "=== model tests (split-K heads dgrad, derivative store)"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_edge_shapes.py tests/test_reference_trainer.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for v in 8 1; do
  echo "=== step, VC_HEAD_DGRAD_SPLITK=$v"; VC_HEAD_DGRAD_SPLITK=$v timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | tee -a gpurun_out/r02ad_ab.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['segments_ms_per_step'])"
done
