#!/bin/bash
# round 2, GPU call I (8 GPUs): bench.py under torchrun at N = 8 and N = 4 (all legs, H2D probe), stream chunk check
set -x
mkdir -p gpurun_out
for c in 4096 2048 1024; do python bench_stream.py --precision split --chunk $c --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $c, d['value'], d['seconds_per_stream'])"; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err || tail -40 gpurun_out/r2_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err || tail -40 gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_8gpu.json','gpurun_out/r2_bench_4gpu.json'):
    d=json.load(open(f))
    print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pcm16', d['e2e']['pcm16_value'], 'h2d/gpu', d['e2e']['h2d_gbs_per_gpu'], 'ceiling', d['e2e']['h2d_ceiling_gbs'], d['e2e']['h2d_probe'])
    print({k:(v.get('value'), v.get('ms_per_step') or v.get('ms') or v.get('ms_per_stream')) for k,v in d['configs'].items() if 'value' in v})
    print(d['configs']['train']['phase_ms'], d['configs']['e2e_dropin'])
    print('sustained', d['sustained']['value'])
PY
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1; head -20 gpurun_out/r2_topo_8gpu.txt
