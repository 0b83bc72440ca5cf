#!/bin/bash
# round 2, GPU call AO: the leg watchdog of bench.py — fired on purpose (1 s), and a normal run
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0 --legs-timeout 1 > gpurun_out/r2_ao_timeout.json 2> gpurun_out/r2_ao_timeout.err; echo "exit code $?"
python -c "
import json; d=json.load(open('gpurun_out/r2_ao_timeout.json')); print('lines ok; value', round(d['value']), {k: (v.get('error') or 'finished') for k, v in d['configs'].items()})"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r2_ao_normal.json 2> gpurun_out/r2_ao_normal.err; echo "exit code $?"
python -c "
import json; d=json.load(open('gpurun_out/r2_ao_normal.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: (v.get('error') or v.get('value') or 'ok') for k, v in d['configs'].items()})"
wc -l gpurun_out/r2_ao_timeout.json gpurun_out/r2_ao_normal.json
