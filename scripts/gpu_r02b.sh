#!/bin/bash
# round 2, second GPU pass: C3 gradient diagnostic, GPU suite, sanitizer on the decode kernels, bench, decode launch list
mkdir -p gpurun_out
timeout 600 python scripts/diag_c3.py 8 32 > gpurun_out/r02b_diag_c3.txt 2>&1
timeout 600 python scripts/diag_c3.py 3 32 >> gpurun_out/r02b_diag_c3.txt 2>&1
timeout 600 python scripts/diag_c3.py 8 12 >> gpurun_out/r02b_diag_c3.txt 2>&1
head -12 gpurun_out/r02b_diag_c3.txt
timeout 1700 python -m pytest tests -m gpu -q -rs > gpurun_out/r02b_pytest.txt 2>&1
tail -15 gpurun_out/r02b_pytest.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_decode.py > gpurun_out/r02b_sanitize_$tool.log 2>&1
  tail -2 gpurun_out/r02b_sanitize_$tool.log
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_c1.json 2> gpurun_out/r02b_bench_c1.err
python -c "import json; d=json.load(open('gpurun_out/r02b_bench_c1.json')); print(d['value'], d['rollout'])"
VIDEOCAD_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10000 --launch-count 200 --csv \
  --log-file gpurun_out/r02b_rollout_launches.csv python scripts/rollout_bench.py --batch 8 --steps 186 --only-feedback --calls 1 > gpurun_out/r02b_rollout_ncu.log 2>&1
tail -2 gpurun_out/r02b_rollout_ncu.log
