#!/bin/bash
# round 2, GPU call G (2 GPUs): two devices in one process, bench.py under torchrun with every leg, reference arm under torchrun
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "two_devices or logmel_ill or library_is" > gpurun_out/r2_tests_g.log 2>&1
tail -4 gpurun_out/r2_tests_g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err || tail -40 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
for k in ('value','ms_per_step','n_gpus','modes','configs'):
    print(k, json.dumps(d.get(k))[:1800])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='api'})
print('sustained', d['sustained']['value'], d['sustained']['ms_per_step'])
PY
VMB_BENCH_CPU_BUDGET_S=10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_2gpu.json 2> gpurun_out/r2_bench_ref_2gpu.err
head -c 600 gpurun_out/r2_bench_ref_2gpu.json
