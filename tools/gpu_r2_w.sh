#!/bin/bash
# round 2, GPU call W: training step A/B on one box — previous build (tools/ab/prev.so) against this one; tests; launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_w.log 2>&1
tail -3 gpurun_out/r2_tests_w.log
for lib in tools/ab/prev.so "" tools/ab/prev.so ""; do VMB_LIB=$lib timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu lib=[$lib]', d['value'], d['ms_per_step'], d['phase_ms'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_w_train_launches.csv python bench_train.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_w_train_launches.csv | head -12
