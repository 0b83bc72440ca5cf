#!/bin/bash
# round 2, GPU call AB: attention backward with per-class values in registers, one-launch weight split — tests, A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_ab.log 2>&1
tail -3 gpurun_out/r2_tests_ab.log
for lib in "" tools/ab/prev.so "" tools/ab/prev.so; do VMB_LIB=$lib timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu lib=[$lib]', round(d['value']), d['ms_per_step'], d['phase_ms'], d['gpu_launches'])"; done
VMB_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_ab_train_launches.csv python bench_train.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_ab_train_launches.csv > gpurun_out/r2_ab_train_launches.txt; head -22 gpurun_out/r2_ab_train_launches.txt
