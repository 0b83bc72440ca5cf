#!/bin/bash
# round 2, GPU call AG: 128 x 64 tiles for the training GEMMs (2.7 waves of half tiles instead of 1.35 waves of full ones) — tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_ag.log 2>&1
tail -3 gpurun_out/r2_tests_ag.log
for f in 1 0 1 0; do VMB_PLANES_NARROW=$f timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu narrow=$f', round(d['value']), d['ms_per_step'], d['phase_ms'], d['final_loss'])"; done
VMB_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ag_train_launches.csv python bench_train.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_ag_train_launches.csv | head -9
