#!/bin/bash
# round 2, GPU call U (2 GPUs): peer-memory optimiser step — tests, A/B of the three gradient exchanges at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s > gpurun_out/r2_tests_u.log 2>&1
tail -6 gpurun_out/r2_tests_u.log
grep -n "peer-memory step\|overlapped vs" gpurun_out/r2_tests_u.log
for ex in peer nccl nccl-overlap peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_train.py --gpus 2 --steps 100 --warmup 10 --exchange $ex 2>gpurun_out/r2_u_$ex.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 2gpu $ex', d['value'], d['ms_per_step'], d['phase_ms'], d['final_loss'])" || tail -5 gpurun_out/r2_u_$ex.err
done
timeout 300 python bench_train.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train 1gpu', d['value'], d['ms_per_step'], d['phase_ms'])"
