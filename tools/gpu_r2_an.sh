#!/bin/bash
# round 2, GPU call AN: ncu --set full of the training GEMM variants (planes_gemm_kernel<3, ...>) on the final build
mkdir -p gpurun_out
VMB_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"planes_gemm_kernel" -s 36 -c 18 -o gpurun_out/r2_an_train_gemm -f python bench_train.py --steps 3 --warmup 3 > gpurun_out/r2_an_ncu.log 2>&1
tail -2 gpurun_out/r2_an_ncu.log | cut -c1-200
ls -la gpurun_out/r2_an_train_gemm.ncu-rep
