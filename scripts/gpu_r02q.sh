#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/ddp_timeline.py --config c1 --out gpurun_out/r02q_ddp_timeline_c1.txt 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/ddp_timeline.py --config c1 --no-sync --out gpurun_out/r02q_ddp_timeline_c1_nosync.txt 2>&1 | tail -3
