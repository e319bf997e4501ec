#!/bin/bash
# One GPU call that re-establishes the state of the repository on a B200 box (run with: gpurun -- 'bash scripts/gpu_check.sh'):
# the whole -m gpu suite, the default bench line (with cpu_baseline), the reference arm, and the micro-benchmarks.
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags)"; timeout 900 python bench.py 2> gpurun_out/bench.err > gpurun_out/bench_c1.json; python -c "import json; d=json.loads(open('gpurun_out/bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'], d['cpu_baseline'])"; tail -3 gpurun_out/bench.err | cut -c1-200
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; cut -c1-400 gpurun_out/bench_ref.json
echo "=== bench c3 (H=1024, T=32) on one GPU"; timeout 600 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c3.err > gpurun_out/bench_c3.json; python -c "import json; d=json.loads(open('gpurun_out/bench_c3.json').read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
echo "=== micro-benchmarks"; timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt | tail -3; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt | tail -3
