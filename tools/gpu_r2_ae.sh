#!/bin/bash
# round 2, GPU call AE: head of micro-batch k on a tail stream next to the front end of micro-batch k + 1 (host-buffer pipeline) — tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "host or pipeline or async or pcm16 or fullsize or stream or batch" > gpurun_out/r2_tests_ae.log 2>&1
tail -3 gpurun_out/r2_tests_ae.log
for f in 1 0 1 0; do VMB_PIPE_TAIL=$f timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2_ae_bench_t$f.json; python -c "
import json; d=json.load(open('gpurun_out/r2_ae_bench_t$f.json')); e=d['e2e']; print('tail', $f, 'value', round(d['value']), d['ms_per_step'], 'e2e', round(e['value']), e['ms_per_step'], 'serial', round(e['serial_value']), 'pcm16', round(e['pcm16_value']), 'b8192', round(d['configs']['batch8192']['value']), d['configs']['batch8192']['ms_each_pass'], 'match', e['matches_device_path'])"; done
