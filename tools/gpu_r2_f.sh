#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "logmel or pcm16 or library_is or sanitize or full" > gpurun_out/r2_tests_f.log 2>&1
tail -5 gpurun_out/r2_tests_f.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_f_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_f_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-30:-26]: print(r[0], r[4][:70], r[-1])
PY
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_f.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['modes'])"
