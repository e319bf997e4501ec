#!/bin/bash
# 8-GPU pass, launched exactly as the driver does
N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 \
  2> gpurun_out/r02p_bench_${N}gpu.err > gpurun_out/r02p_bench_c1_${N}gpu.json
tail -3 gpurun_out/r02p_bench_${N}gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_c1_${N}gpu.json'))
print(d['value'], d['ms_per_step'], d['ms_per_step_without_allreduce'], d['exposed_allreduce_ms'], d['e2e']['value'])
print('c3', d['c3'])
print('rollout', {k:v for k,v in d['rollout'].items() if k!='roofline'})
print('through_trainer', d['through_trainer'])"
