#!/bin/bash
# round 2, GPU call S: training — BatchNorm statistics in the GEMM epilogue, weight splits on the side stream; fused head
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_s.log 2>&1
tail -3 gpurun_out/r2_tests_s.log
for f in 1 0 1 0; do VMB_TRAIN_STATS_FUSE=$f timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train stats-fuse', $f, d['value'], d['ms_per_step'], d['phase_ms'])"; done
for c in "" "1,2,1" "3"; do python tools/time_head.py 256 300 $c; VMB_MLA_FUSE=0 python tools/time_head.py 256 300 $c; done
