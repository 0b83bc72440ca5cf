#!/bin/bash
# round 2, GPU call AF: ncu launch list of the inference step on the end-of-round build + --set full capture of the head's kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2_af_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config-legs --sustained-seconds 0 > gpurun_out/r2_af_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_af_launches.csv > gpurun_out/r2_af_launches.txt; cat gpurun_out/r2_af_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"planes_gemm|attention_pool|rows_affine|sigmoid_affine" -s 20 -c 12 -o gpurun_out/r2_af_head_full -f python tools/time_head.py 256 6 > gpurun_out/r2_af_ncu_full.log 2>&1
ls -la gpurun_out/r2_af_head_full.ncu-rep
