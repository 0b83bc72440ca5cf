#!/bin/bash
# round 2, GPU call R: the fused head without the side stream is 2x slower than the unfused one — find out why
mkdir -p gpurun_out
for env in "VMB_MLA_FORK=0" "VMB_MLA_FORK=0 VMB_MLA_FUSE=0" "VMB_MLA_FORK=0 VMB_PDL=0" "VMB_MLA_FORK=0 VMB_MLA_FUSE_OUT=0" "VMB_MLA_FUSE_OUT=0" ; do
  env $env python tools/time_head.py 256 300
done
for conf in 1 3 1,2,1; do
  for env in "" "VMB_MLA_FUSE=0" "VMB_MLA_FUSE_OUT=0"; do env $env python tools/time_head.py 256 300 $conf; done
done
