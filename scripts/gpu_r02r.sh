#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/adam_bench.py 2>&1 | grep -v Warning | tee gpurun_out/r02r_adam_bench.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "adam or optim or Adam" 2>&1 | tail -3
