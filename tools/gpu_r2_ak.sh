#!/bin/bash
# round 2, GPU call AK: weight-gradient GEMMs on MN-major operands (no transposed planes) — tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_ak.log 2>&1
tail -15 gpurun_out/r2_tests_ak.log | cut -c1-300
VMB_TRAIN_MN_DW=0 timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -2
for f in 1 0 1 0; do VMB_TRAIN_MN_DW=$f timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu mn_dw=$f', round(d['value']), d['ms_per_step'], d['final_loss'])"; done
