#!/bin/bash
# round 2, GPU call AI (4 GPUs): 2-GPU tests and the full bench line on the end-of-round build
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_4gpu_v2.json 2> gpurun_out/r2_bench_4gpu_v2.err || tail -40 gpurun_out/r2_bench_4gpu_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_4gpu_v2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pcm16', d['e2e']['pcm16_value'], 'h2d/gpu', d['e2e']['h2d_gbs_per_gpu'], 'ceiling', d['e2e']['h2d_ceiling_gbs'])
print({k:(v.get('value'), v.get('ms_per_step') or v.get('ms') or v.get('ms_per_stream'), v.get('error')) for k,v in d['configs'].items()})
print(d['configs']['train'])
PY
