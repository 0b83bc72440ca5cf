#!/bin/bash
# round 2, GPU call E: full GPU test-suite on the list-based exact path, bench line, launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests_e.log 2>&1
tail -6 gpurun_out/r2_tests_e.log; grep -E "^mAP|VGGish\(preprocess" gpurun_out/r2_tests_e.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err || tail -30 gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e.json'))
for k in ('value','ms_per_step','stage_ms_per_step','modes'):
    print(k, json.dumps(d.get(k))[:1500])
print('sustained', d['sustained']['value'], d['sustained']['ms_per_step'], d['sustained']['stage_ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_e_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_e_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-30:]: print(r[0], r[4][:70], r[-1])
PY
