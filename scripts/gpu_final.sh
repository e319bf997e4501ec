#!/bin/bash
# end-of-round verification on one B200: GPU suite, smoke, both bench arms with default flags, launch list of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/final_pytest.txt 2>&1; tail -6 gpurun_out/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-400 gpurun_out/final_bench_reference.json
timeout 900 python bench.py > gpurun_out/final_bench_c1.json 2> gpurun_out/final_bench_c1.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench_c1.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('segments', d['segments_ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['achieved'], 'cpu', d['cpu_baseline'])
print('rollout', {k:v for k,v in d['rollout'].items() if k!='roofline'}, d['rollout']['roofline']['frac'])
print('through_trainer', d['through_trainer']['value'], 'fused', d['through_trainer_fused'])"
VIDEOCAD_B200_GRAPHS=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_one_step.csv python bench.py --warmup 3 --profile-step > gpurun_out/final_ncu_step.log 2>&1; wc -l gpurun_out/final_launches_one_step.csv
python scripts/summarize_launches.py gpurun_out/final_launches_one_step.csv | head -24
