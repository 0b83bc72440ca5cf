#!/bin/bash
# round 2, GPU call C: full GPU test-suite, precision table (256 clips, label-set spread), new bench line, launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests_c.log 2>&1
tail -12 gpurun_out/r2_tests_c.log
python tools/precision_table.py --out gpurun_out/r2_precision_table.json > gpurun_out/r2_precision_table.log 2>&1
grep -v "^        " gpurun_out/r2_precision_table.log | tail -8
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err || tail -30 gpurun_out/r2_bench_c.err
head -c 6000 gpurun_out/r2_bench_c.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_c_ncu.log 2>&1
grep -E "logmel" gpurun_out/r2_c_launches.csv | tail -4 | cut -c1-40,200-330
