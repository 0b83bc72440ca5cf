#!/bin/bash
# round 2, GPU call AA: BatchNorm-backward reductions in the dX GEMM epilogue — tests, A/B with the switch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_aa.log 2>&1
tail -3 gpurun_out/r2_tests_aa.log
for f in 1 0 1 0; do VMB_TRAIN_GRADSTATS_FUSE=$f timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu gradstats=$f', round(d['value']), d['ms_per_step'], d['phase_ms'], d['gpu_launches'], d['final_loss'])"; done
