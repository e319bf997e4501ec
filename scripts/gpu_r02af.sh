#!/bin/bash
# 2 x B200, final code, launched as the driver does: the N > 1 line with the through_trainer_fused leg, C3 under DDP and the rollout
mkdir -p gpurun_out
timeout 230 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r02af_bench_2gpu.err > gpurun_out/r02af_bench_c1_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02af_bench_c1_2gpu.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'exposed', d['exposed_allreduce_ms'])
print('through_trainer', d['through_trainer'].get('value'), 'fused', {k: v for k, v in d['through_trainer_fused'].items() if k != 'note'})
print('c3', d['c3']); print('rollout', d['rollout']['frames_per_s'])"
tail -3 gpurun_out/r02af_bench_2gpu.err | cut -c1-300
