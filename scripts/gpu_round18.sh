#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests (PDL on)"; timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_loss.py tests/test_optim.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== model tests (PDL on)"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== bench PDL on"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
echo "=== bench PDL off"; VC_PDL=0 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench_nopdl.err > gpurun_out/bench_c1_nopdl.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1_nopdl.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_nopdl.err
