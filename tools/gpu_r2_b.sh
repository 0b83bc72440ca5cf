#!/bin/bash
# round 2, GPU call B: exact-kernel latency (launch list), mAP mix probe, logmel tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "logmel or pcm16" > gpurun_out/r2_tests_b.log 2>&1
tail -25 gpurun_out/r2_tests_b.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_ncu.log 2>&1
grep -E "logmel|conv1" gpurun_out/r2_b_launches.csv | tail -8
timeout 1200 python tools/map_mix_probe.py > gpurun_out/r2_map_mix_probe.log 2>&1
cat gpurun_out/r2_map_mix_probe.log
