#!/bin/bash
# round 2, GPU call AP (2 GPUs): bench.py under torchrun after the legs moved behind the line assembly (watchdog)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --sustained-seconds 0 > gpurun_out/r2_ap_2gpu.json 2> gpurun_out/r2_ap_2gpu.err; echo "exit code $?"
python -c "
import json; d=json.load(open('gpurun_out/r2_ap_2gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: (v.get('error') or v.get('value') or 'ok') for k, v in d['configs'].items()}, d['configs']['train']['gradient_exchange'][:40])" || tail -20 gpurun_out/r2_ap_2gpu.err
wc -l gpurun_out/r2_ap_2gpu.json
