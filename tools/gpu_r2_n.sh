#!/bin/bash
# round 2, final 8-GPU record on the final build (every leg of bench.py under torchrun)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err || tail -40 gpurun_out/r2_bench_8gpu_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_8gpu_final.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pcm16', d['e2e']['pcm16_value'], 'h2d/gpu', d['e2e']['h2d_gbs_per_gpu'], 'ceiling', d['e2e']['h2d_ceiling_gbs'])
print({k:(v.get('value'), v.get('ms_per_step') or v.get('ms') or v.get('ms_per_stream')) for k,v in d['configs'].items() if 'value' in v})
print(d['configs']['train']['phase_ms'], d['configs']['stream_1h'].get('fp16_mode'))
print('sustained', d['sustained']['value'], d['stage_ms_per_step'])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -x -q -k "two_devices or data_parallel or head" | tail -3
