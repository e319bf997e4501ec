#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_model_gpu.py tests/test_loss.py tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== bench fused"; timeout 600 python bench.py --steps 20 --warmup 4 2> gpurun_out/bench.err | tee gpurun_out/bench_c1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench.err
echo "=== bench torch loss"; timeout 600 python bench.py --steps 20 --warmup 4 --loss torch --no-cpu-baseline 2> gpurun_out/bench_tl.err | tee gpurun_out/bench_c1_torchloss.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -3 gpurun_out/bench_tl.err
echo "=== bench c3"; timeout 600 python bench.py --steps 10 --warmup 3 --config c3 --no-cpu-baseline 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -3 gpurun_out/bench_c3.err
