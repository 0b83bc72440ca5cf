#!/bin/bash
# round 2, GPU call AC: the driver's sequence on the current build (all GPU tests, smoke, both bench arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_ac.log 2>&1
tail -3 gpurun_out/r2_tests_ac.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_ac.json 2>/dev/null
head -c 300 gpurun_out/r2_bench_ref_ac.json; echo
time python bench.py > gpurun_out/r2_bench_ac.json 2> gpurun_out/r2_bench_ac.err || tail -20 gpurun_out/r2_bench_ac.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_ac.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['sustained']['frac'], 'whole', d['roofline']['whole_step_frac'], d['stage_ms_per_step'])
print({k:(v.get('value')) for k,v in d['configs'].items() if 'value' in v}, d['clocks'])
print(d['configs']['train'])
PY
