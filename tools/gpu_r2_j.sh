#!/bin/bash
# round 2, GPU call J: full GPU test-suite, smoke, head-training timing after the dW changes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_j.log 2>&1
tail -4 gpurun_out/r2_tests_j.log
python __graft_entry__.py smoke
python bench_train.py --steps 50 --warmup 5 > gpurun_out/r2_bench_train_1gpu.json 2> gpurun_out/r2_bench_train.err
cat gpurun_out/r2_bench_train_1gpu.json | head -c 1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_j_train_launches.csv python bench_train.py --steps 2 --warmup 3 > gpurun_out/r2_j_train_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_j_train_launches.csv | head -14
