#!/bin/bash
# round 2, GPU call AM (8 GPUs): head training on the end-of-round build, peer-memory step against the NCCL all-reduce
mkdir -p gpurun_out
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench_train.py --gpus 8 --steps 200 --warmup 10 --exchange $ex 2>gpurun_out/r2_am_$ex.err > gpurun_out/r2_bench_train_8gpu_${ex}_v2.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_8gpu_${ex}_v2.json')); print('train 8gpu $ex', round(d['value']), d['ms_per_step'], d['phase_ms'], d['final_loss'])" || tail -5 gpurun_out/r2_am_$ex.err
done
timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null > gpurun_out/r2_bench_train_1gpu_v2.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_1gpu_v2.json')); print('train 1gpu (same box)', round(d['value']), d['ms_per_step'])"
