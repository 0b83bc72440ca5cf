#!/bin/bash
# round 2, GPU call AU: the documented A/B switches still select working paths (old split GEMM kernel, transposed planes, eager launches)
mkdir -p gpurun_out
VMB_PLANES_GEMM=0 timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -2
VMB_TRAIN_MN_DW=0 VMB_TRAIN_GRAPH=0 VMB_TRAIN_ESTATS_FUSE=0 VMB_TRAIN_STATS_FUSE=0 VMB_TRAIN_GRADSTATS_FUSE=0 timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -2
VMB_PLANES_GEMM=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "head" 2>&1 | tail -2
