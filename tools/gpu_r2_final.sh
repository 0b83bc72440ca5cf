#!/bin/bash
# round 2, final validation on one GPU: the driver's sequence (GPU tests, smoke, both bench arms) on the final build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests_final.log 2>&1
tail -4 gpurun_out/r2_tests_final.log
python __graft_entry__.py smoke
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>/dev/null
head -c 300 gpurun_out/r2_bench_ref_final.json; echo
time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_bench_final_20.json 2> gpurun_out/r2_bench_final_20.err
time python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_final_20.json','gpurun_out/r2_bench_final.json'):
    d=json.load(open(f))
    print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['sustained']['frac'], 'whole', d['roofline']['whole_step_frac'], d['stage_ms_per_step'])
    print({k:(v.get('value')) for k,v in d['configs'].items() if 'value' in v}, d['clocks'])
PY
