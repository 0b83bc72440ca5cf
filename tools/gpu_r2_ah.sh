#!/bin/bash
# round 2, GPU call AH: which of the training GEMM variants gain from 128 x 64 tiles (mask: 1 plain, 2 statistics, 4 gradient statistics)
for m in 0 4 5 1 7 0 4 5; do VMB_PLANES_NARROW=$m timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu narrow mask=$m', round(d['value']), d['ms_per_step'])"; done
