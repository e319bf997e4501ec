#!/bin/bash
# round 2 evidence pass (one GPU): full suite + smoke, default bench line, reference arm, one-step launch list, ncu --set full of the
# dominant kernel (traffic), of the decode kernels and of the ingest kernel, micro-benchmarks
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags)"; timeout 1200 python bench.py 2> gpurun_out/r02o_bench_c1.err > gpurun_out/r02o_bench_c1.json
python -c "import json; d=json.loads(open('gpurun_out/r02o_bench_c1.json').read()); print(d['value'], d['e2e']['value'], d['segments_ms_per_step'], d['roofline']['frac'], d['cpu_baseline'], d['through_trainer'], d['rollout']['ms_per_rollout'])"
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 2> gpurun_out/r02o_bench_ref.err > gpurun_out/r02o_bench_ref.json; cut -c1-300 gpurun_out/r02o_bench_ref.json
echo "=== bench c3 one GPU"; timeout 600 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --no-rollout 2> gpurun_out/r02o_bench_c3.err > gpurun_out/r02o_bench_c3.json; python -c "import json; d=json.loads(open('gpurun_out/r02o_bench_c3.json').read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
echo "=== one-step launch list"
VIDEOCAD_B200_GRAPHS=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_launches_one_step.csv python bench.py --warmup 3 --profile-step > gpurun_out/r02o_ncu_step.log 2>&1; wc -l gpurun_out/r02o_launches_one_step.csv
echo "=== ncu full: pair GEMM to_qkv forward"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_pair -c 1 --launch-skip 3 -f -o gpurun_out/r02o_pair_qkv python scripts/gemm_bench.py --only "epi qkv fwd" --iters 2 > gpurun_out/r02o_ncu_pair.log 2>&1; tail -1 gpurun_out/r02o_ncu_pair.log
echo "=== ncu full: decode kernels"
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --set full --clock-control none -k regex:dec_ --launch-skip 2000 -c 12 -f -o gpurun_out/r02o_decode python scripts/rollout_bench.py --batch 8 --steps 60 --only-feedback --calls 1 > gpurun_out/r02o_ncu_dec.log 2>&1; tail -1 gpurun_out/r02o_ncu_dec.log
echo "=== micro-benchmarks"; timeout 300 python scripts/gemm_bench.py > gpurun_out/r02o_gemm_bench.txt 2>&1; tail -3 gpurun_out/r02o_gemm_bench.txt; timeout 300 python scripts/attn_bench.py > gpurun_out/r02o_attn_bench.txt 2>&1; head -4 gpurun_out/r02o_attn_bench.txt
timeout 300 python scripts/rollout_bench.py > gpurun_out/r02o_rollout.json 2>&1; cat gpurun_out/r02o_rollout.json
