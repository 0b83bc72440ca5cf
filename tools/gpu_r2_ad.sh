#!/bin/bash
# round 2, GPU call AD (8 GPUs): every leg of bench.py under torchrun on the current build; the three gradient exchanges of head training
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu_ad.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu_ad.json 2> gpurun_out/r2_bench_8gpu_ad.err || tail -40 gpurun_out/r2_bench_8gpu_ad.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_8gpu_ad.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pcm16', d['e2e']['pcm16_value'], 'h2d/gpu', d['e2e']['h2d_gbs_per_gpu'], 'ceiling', d['e2e']['h2d_ceiling_gbs'])
print({k:(v.get('value'), v.get('ms_per_step') or v.get('ms') or v.get('ms_per_stream')) for k,v in d['configs'].items() if 'value' in v})
print(d['configs']['train'])
print('sustained', d['sustained']['value'], d['stage_ms_per_step'])
PY
for ex in peer nccl nccl-overlap; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench_train.py --gpus 8 --steps 200 --warmup 10 --exchange $ex 2>gpurun_out/r2_ad_$ex.err > gpurun_out/r2_bench_train_8gpu_$ex.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_train_8gpu_$ex.json')); print('train 8gpu $ex', round(d['value']), d['ms_per_step'], d['phase_ms'], d['final_loss'])" || tail -5 gpurun_out/r2_ad_$ex.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -x -q -k "two_devices or peer or overlapped" 2>&1 | tail -3
