#!/bin/bash
mkdir -p gpurun_out
echo "=== attn/ln bench"; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt | head -6
echo "=== ncu"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:"vit_attn|ln_bwd" -c 3 -f -o gpurun_out/vit_attn2 python scripts/attn_bench.py --once > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
