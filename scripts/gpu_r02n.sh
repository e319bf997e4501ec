#!/bin/bash
# 2-GPU pass after splitting the encoder backward into three autograd nodes: full GPU suite on GPU 0, torchrun bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-rollout \
  > gpurun_out/r02n_bench_c1_2gpu.json 2> gpurun_out/r02n_bench_c1_2gpu.err
tail -2 gpurun_out/r02n_bench_c1_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_c1_2gpu.json'))
print(d['value'], d['ms_per_step'], d['ms_per_step_without_allreduce'], d['exposed_allreduce_ms'], d['segments_ms_per_step'])
print('c3', d['c3'])"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout > gpurun_out/r02n_bench_c1.json 2> gpurun_out/r02n_bench_c1.err
python -c "import json; d=json.load(open('gpurun_out/r02n_bench_c1.json')); print(d['value'], d['ms_per_step'], d['segments_ms_per_step'])"
