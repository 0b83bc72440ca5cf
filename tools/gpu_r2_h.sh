#!/bin/bash
# round 2, GPU call H: stream chunking experiment, full bench line, ncu launch list + --set full capture of one step
set -x
mkdir -p gpurun_out
for c in 4096 2048 1024 512; do python bench_stream.py --precision split --chunk $c --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $c, d['value'], d['seconds_per_stream'])"; done
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err || tail -30 gpurun_out/r2_bench_h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_h.json'))
for k in ('value','ms_per_step','stage_ms_per_step','modes'):
    print(k, json.dumps(d.get(k))[:1500])
print('sustained', d['sustained']['value'], d['sustained']['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_h_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_h_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-30:-26]: print(r[0], r[4][:70], r[-1])
PY
# one whole step under ncu --set full: launches 90..119 of the list above = the last step (30 kernels)
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip 90 --launch-count 30 -o gpurun_out/r2_step_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_h_ncu_full.log 2>&1
ls -la gpurun_out/r2_step_full.ncu-rep
