#!/bin/bash
# N-GPU weak-scaling line (run with: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'), launched exactly as the driver does
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 4 2> gpurun_out/bench_${N}gpu.err | tee gpurun_out/bench_c1_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['segments_ms_per_step'])"
tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
