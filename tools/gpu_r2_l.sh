#!/bin/bash
# round 2, GPU call L: dW GEMMs on a side stream in head training (tests + A/B timing)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s > gpurun_out/r2_tests_l.log 2>&1
tail -4 gpurun_out/r2_tests_l.log
for f in 1 0 1 0; do VMB_TRAIN_FORK=$f python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fork', $f, d['value'], d['ms_per_step'], d['phase_ms'])"; done
