#!/bin/bash
mkdir -p gpurun_out
VIDEOCAD_B200_GRAPHS=0 VC_MEGA_DEBUG=100 timeout 300 python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback --calls 1 2>&1 | grep -v Warn | tail -75 | tee gpurun_out/r02u_mega_trace.txt
