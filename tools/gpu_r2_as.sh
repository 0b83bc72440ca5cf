#!/bin/bash
# round 2, GPU call AS: pooling kernel writes the output Linear's operand planes; vectorised statistics kernel — tests, A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_as.log 2>&1
tail -4 gpurun_out/r2_tests_as.log | cut -c1-250
for lib in "" tools/ab/prev.so "" tools/ab/prev.so; do VMB_LIB=$lib timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu lib=[$lib]', round(d['value']), d['ms_per_step'], d['final_loss'], d['gpu_launches'])"; done
