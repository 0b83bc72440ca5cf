#!/bin/bash
# round 2, GPU call A: fp16 body + float64 log-mel path: precision table, parity tests, short bench
set -x
mkdir -p gpurun_out
python tools/precision_table.py --out gpurun_out/r2_precision_table.json > gpurun_out/r2_precision_table.log 2>&1
tail -40 gpurun_out/r2_precision_table.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests_a.log 2>&1
tail -15 gpurun_out/r2_tests_a.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
cat gpurun_out/r2_bench_a.json | head -c 3000
