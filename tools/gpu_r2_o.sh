#!/bin/bash
# round 2, GPU call O: re-validation of HEAD after the container was re-created (tests, smoke, bench, training bench A/B)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_o.log 2>&1
tail -3 gpurun_out/r2_tests_o.log
python __graft_entry__.py smoke 2>&1 | tail -2
for f in 1 0; do VMB_PLANES_GEMM=$f timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train planes', $f, d['value'], d['ms_per_step'], d['phase_ms'])"; done
for f in 1 0; do VMB_PLANES_GEMM=$f timeout 600 python bench.py --steps 50 --warmup 5 --no-config-legs 2>/dev/null > gpurun_out/r2_o_bench_p$f.json; python -c "
import json; d=json.load(open('gpurun_out/r2_o_bench_p$f.json')); print('planes', $f, d['value'], d['ms_per_step'], d['stage_ms_per_step'])"; done
