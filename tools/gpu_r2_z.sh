#!/bin/bash
# round 2, GPU call Z: ncu --set full on the element-wise kernels of the training step (what stalls them)
mkdir -p gpurun_out
VMB_TRAIN_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"att_backward|bn_time_backward|tile_split_kernel" -s 40 -c 14 -o gpurun_out/r2_z_train_elem -f python bench_train.py --steps 3 --warmup 3 > gpurun_out/r2_z_ncu.log 2>&1
tail -3 gpurun_out/r2_z_ncu.log
ls -la gpurun_out/*.ncu-rep
