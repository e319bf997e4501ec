#!/bin/bash
# A/B of the GELU derivative store (VC_ACT_GELU_DSTORE / VC_ACT_MUL_AUX, default on; VC_GELU_DSTORE=0 = previous path): kernel and
# model parity tests, the affected GEMMs in isolation, the C1 step back to back on the same box
mkdir -p gpurun_out
echo "=== kernel + model tests (new path)"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_edge_shapes.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
echo "=== model tests, previous path"; VC_GELU_DSTORE=0 timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -x -k "golden or c2_full or dropout" 2>&1 | tail -2
for i in 1 2; do
  for v in 1 0; do
    echo "=== step, VC_GELU_DSTORE=$v (run $i)"; VC_GELU_DSTORE=$v timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | tee -a gpurun_out/r02ac_ab.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['segments_ms_per_step'], d['clocks'])"
  done
done
