#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]  -> tools/ab/NAME.so (igemm_sm100.cu rebuilt with the flags, other objects reused)
set -e
cd "$(dirname "$0")/.."
P=$(echo audio-*_b200)
name=$1; shift
python $P/b200/build.py > /dev/null
mkdir -p tools/ab /tmp/vmb_variant
for f in "${VARIANT_SRCS:-igemm_sm100}"; do :; done
objs=""
for o in $P/csrc/build/*.o; do
  b=$(basename $o .o)
  if [[ " ${VARIANT_SRCS:-igemm_sm100} " == *" $b "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static "$@" -I include -c $P/csrc/$b.cu -o /tmp/vmb_variant/$name.$b.o
    objs="$objs /tmp/vmb_variant/$name.$b.o"
  else
    objs="$objs $o"
  fi
done
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o tools/ab/$name.so $objs
echo tools/ab/$name.so
