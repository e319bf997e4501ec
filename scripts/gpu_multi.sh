#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -5
echo "=== tests (attention alias)"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "attention or golden or graph" 2>&1 | tail -5
echo "=== bench c1 x2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 4 2> gpurun_out/bench_2gpu.err | tee gpurun_out/bench_c1_2gpu.json | cut -c1-220; tail -4 gpurun_out/bench_2gpu.err | cut -c1-300
echo "=== bench c1 x1 GPU"; timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2> gpurun_out/bench_1gpu.err | tee gpurun_out/bench_c1_1gpu.json | cut -c1-220
echo "=== bench c3 x1 GPU"; timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_1gpu.json | cut -c1-220; tail -3 gpurun_out/bench_c3.err | cut -c1-300
