#!/bin/bash
# round 2, GPU call T (2 GPUs): overlapped all-reduce — tests, A/B of bench_train at N = 2, 1-GPU step with the new output-layer kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s > gpurun_out/r2_tests_t.log 2>&1
tail -4 gpurun_out/r2_tests_t.log
timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu', d['value'], d['ms_per_step'], d['phase_ms'])"
for flag in "" "--no-overlap" "" "--no-overlap"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_train.py --gpus 2 --steps 100 --warmup 10 $flag 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 2gpu $flag', d['value'], d['ms_per_step'], d['phase_ms'])"
done
