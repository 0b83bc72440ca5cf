#!/bin/bash
# round 2, GPU call M: planes GEMM kernel (shared-plane stages) — head + training tests, A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s > gpurun_out/r2_tests_m.log 2>&1
tail -4 gpurun_out/r2_tests_m.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "head or mla or ensemble or map or module or standalone" >> gpurun_out/r2_tests_m.log 2>&1
tail -4 gpurun_out/r2_tests_m.log
for f in 1 0 1 0; do VMB_PLANES_GEMM=$f timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('planes', $f, d['value'], d['ms_per_step'], d['phase_ms'])"; done
for f in 1 0; do VMB_PLANES_GEMM=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-config-legs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('planes', $f, d['value'], d['ms_per_step'], d.get('stages_ms') or d.get('stage_ms'))"; done
