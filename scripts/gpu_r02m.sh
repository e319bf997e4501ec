#!/bin/bash
# 2-GPU pass: NCCL gradient-equality test, torchrun bench with the C3 and rollout legs
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_ddp_nccl.py -m gpu -q -x -rs 2>&1 | tail -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02m_bench_c1_2gpu.json 2> gpurun_out/r02m_bench_c1_2gpu.err
tail -3 gpurun_out/r02m_bench_c1_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_c1_2gpu.json'))
print(d['value'], d['ms_per_step'], d['exposed_allreduce_ms'], d['through_trainer'])
print('c3', d['c3'])
print('rollout', {k:v for k,v in d['rollout'].items() if k!='roofline'})"
VIDEOCAD_B200_SIDE_PRIORITY=-1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-rollout \
  > gpurun_out/r02m_bench_c1_2gpu_prio.json 2> gpurun_out/r02m_bench_c1_2gpu_prio.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_c1_2gpu_prio.json'))
print('prio', d['value'], d['ms_per_step'], d['exposed_allreduce_ms']); print('c3', d['c3'])"
