#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -5
echo "=== bench c1 x2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 4 2> gpurun_out/bench_2gpu.err | tee gpurun_out/bench_c1_2gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['segments_ms_per_step'])"; tail -4 gpurun_out/bench_2gpu.err | cut -c1-300
echo "=== bench c1 x1 GPU (same box)"; timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
