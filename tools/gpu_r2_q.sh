#!/bin/bash
# round 2, GPU call Q: where the head's time goes — A/B over the switches + ncu launch list of the fused and unfused head
mkdir -p gpurun_out
for env in "" "VMB_MLA_FUSE=0" "VMB_MLA_FORK=0" "VMB_MLA_FUSE=0 VMB_MLA_FORK=0" "VMB_PDL=0" "VMB_PLANES_GEMM=0"; do
  env $env python tools/time_head.py 256 300
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_q_head_fused.csv python tools/time_head.py 256 4 > /dev/null 2>&1
VMB_MLA_FUSE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_q_head_plain.csv python tools/time_head.py 256 4 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2_q_head_fused.csv | head -20
python tools/ncu_summary.py launches gpurun_out/r2_q_head_plain.csv | head -20
