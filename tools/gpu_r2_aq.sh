#!/bin/bash
# round 2, GPU call AQ: training tests incl. the single-rank optimiser-step test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/r2_tests_aq.log 2>&1
tail -4 gpurun_out/r2_tests_aq.log | cut -c1-250
