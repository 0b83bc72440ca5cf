#!/bin/bash
# round 2, GPU call X: the training step as a CUDA graph — tests, A/B against eager launches and the previous build, host enqueue time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s > gpurun_out/r2_tests_x.log 2>&1
tail -5 gpurun_out/r2_tests_x.log
for g in 1 0 1 0; do VMB_TRAIN_GRAPH=$g timeout 300 python bench_train.py --steps 200 --warmup 10 2>gpurun_out/r2_x_err_$g.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu graph=$g', round(d['value']), d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], d['phase_ms'], d['gpu_launches'], d['final_loss'])" || tail -5 gpurun_out/r2_x_err_$g.log; done
VMB_LIB=tools/ab/prev.so timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu prev.so', round(d['value']), d['ms_per_step'], d['phase_ms'])"
