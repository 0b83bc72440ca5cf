#!/bin/bash
# round 2, GPU call V: training element-wise kernels (64x64 tile kernel with paired stores, attention backward reduction) — tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_v.log 2>&1
tail -4 gpurun_out/r2_tests_v.log
for i in 1 2; do timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu', d['value'], d['ms_per_step'], d['phase_ms'], d['gpu_launches'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_v_train_launches.csv python bench_train.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_v_train_launches.csv | head -30
