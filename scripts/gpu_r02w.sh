#!/bin/bash
# three gradient parts per image encoder: NCCL gradient equality on 2 GPUs, then the bench under the driver's launch line
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_ddp_nccl.py -m gpu -x -q 2>&1 | tail -3; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline \
  2> gpurun_out/r02w_bench_${N}gpu.err > gpurun_out/r02w_bench_c1_${N}gpu.json
tail -2 gpurun_out/r02w_bench_${N}gpu.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_c1_${N}gpu.json'))
print('c1', d['value'], d['ms_per_step'], d['ms_per_step_without_allreduce'], d['exposed_allreduce_ms'], d['e2e']['value'])
print('c3', d['c3']['ms_per_step'], d['c3']['value'], d['c3']['exposed_allreduce_ms'])
print('rollout', d['rollout']['ms_per_rollout'], d['rollout']['ms_per_decode_step'])"
