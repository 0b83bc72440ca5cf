#!/bin/bash
# round 2, GPU call K: head fork on a side stream (tests + A/B timing), fp16 stream checks, bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -s -k "head or ensemble or pipeline or shard or batch_size or one_hour or standalone or just_bottlenecks" > gpurun_out/r2_tests_k.log 2>&1
tail -5 gpurun_out/r2_tests_k.log; grep -E "1-hour stream" gpurun_out/r2_tests_k.log
for f in 1 0; do VMB_MLA_FORK=$f python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-config-legs --sustained-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fork', $f, d['value'], d['ms_per_step'], d['stage_ms_per_step']['mla'], d['single_clip_latency_ms'])"; done
for f in 1 0; do VMB_MLA_FORK=$f python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-config-legs --sustained-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fork', $f, d['value'], d['ms_per_step'], d['stage_ms_per_step']['mla'], d['single_clip_latency_ms'])"; done
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err || tail -30 gpurun_out/r2_bench_k.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_k.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step']); print(d['configs']['stream_1h']); print(d['configs']['train']['value'], d['configs']['e2e_dropin'])"
