#!/bin/bash
# round 2, GPU call AR: norm0 statistics of levels >= 1 from the pass that writes the embedding; tile pass without the shared-memory copy when nothing reads it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_ar.log 2>&1
tail -4 gpurun_out/r2_tests_ar.log | cut -c1-250
for f in 1 0 1 0; do VMB_TRAIN_ESTATS_FUSE=$f timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu estats=$f', round(d['value']), d['ms_per_step'], d['final_loss'])"; done
