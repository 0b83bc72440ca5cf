#!/bin/bash
# round 2, GPU call AL: launch list of the training step after the MN-major change; tile-width mask re-check
mkdir -p gpurun_out
for m in 5 0 7 4; do VMB_PLANES_NARROW=$m timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu narrow mask=$m', round(d['value']), d['ms_per_step'])"; done
VMB_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_al_train_launches.csv python bench_train.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_al_train_launches.csv > gpurun_out/r2_al_train_launches.txt; cat gpurun_out/r2_al_train_launches.txt
