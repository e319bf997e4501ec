#!/bin/bash
# in-situ A/B of the 2-SM GEMM tile-width choice (automatic vs forced 256), two runs each
mkdir -p gpurun_out
for i in 1 2; do
for bn in 0 256; do
echo "=== bench VC_GEMM_PAIR_BN=$bn run $i"; VC_GEMM_PAIR_BN=$bn timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_bn${bn}_$i.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_bn${bn}_$i.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'], d['roofline']['all_gemms']['ms_per_step'])"; tail -3 gpurun_out/bench.err
done; done
