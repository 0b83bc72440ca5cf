#!/bin/bash
# round 2, GPU call P: fused head epilogues — bit identity vs the separate kernels, head tests, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "head or mla or ensemble or map or module or standalone or bottleneck or pipeline or invariance" > gpurun_out/r2_tests_p.log 2>&1
tail -5 gpurun_out/r2_tests_p.log; grep -n "launches fused" gpurun_out/r2_tests_p.log
for f in 1 0 1 0; do VMB_MLA_FUSE=$f timeout 600 python bench.py --steps 50 --warmup 5 --no-config-legs --no-cpu-baseline 2>/dev/null > gpurun_out/r2_p_bench_f$f.json; python -c "
import json; d=json.load(open('gpurun_out/r2_p_bench_f$f.json')); print('fuse', $f, d['value'], d['ms_per_step'], d['stage_ms_per_step']['mla'], d['single_clip_latency_ms'])"; done
