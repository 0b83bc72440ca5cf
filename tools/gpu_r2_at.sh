#!/bin/bash
# round 2, GPU call AT (8 GPUs): head training at HEAD, peer-memory step, and the single-GPU step on the same box
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench_train.py --gpus 8 --steps 200 --warmup 10 --exchange peer 2>gpurun_out/r2_at_peer.err > gpurun_out/r2_bench_train_8gpu_peer_v3.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_8gpu_peer_v3.json')); print('train 8gpu peer', round(d['value']), d['ms_per_step'], d['final_loss'])" || tail -5 gpurun_out/r2_at_peer.err
timeout 300 python bench_train.py --steps 200 --warmup 10 2>/dev/null > gpurun_out/r2_bench_train_1gpu_v3.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_1gpu_v3.json')); print('train 1gpu (same box)', round(d['value']), d['ms_per_step'])"
