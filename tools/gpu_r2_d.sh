#!/bin/bash
# round 2, GPU call D: exact-kernel latency after the lag split, logmel tests, bench line, train-step launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "logmel or pcm16 or library_is" > gpurun_out/r2_tests_d.log 2>&1
tail -5 gpurun_out/r2_tests_d.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err || tail -30 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
for k in ('value','ms_per_step','stage_ms_per_step','modes','configs','sustained','roofline','clocks'):
    print(k, json.dumps(d.get(k))[:1500])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_d_ncu.log 2>&1
grep -E "logmel" gpurun_out/r2_d_launches.csv | tail -4 | cut -c1-40,200-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_d_train_launches.csv python bench_train.py --steps 2 --warmup 3 > gpurun_out/r2_d_train_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_d_train_launches.csv')) if len(r)>10 and r[0].isdigit()]
# the last step = the last ~60 launches
last=rows[-70:]
for r in last: print(r[0], r[4][:90], r[-1])
PY
