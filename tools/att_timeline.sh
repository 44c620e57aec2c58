#!/bin/bash
# builds a -DADA_BRINGUP copy of the library on the GPU box and prints the attention2 timeline for a few variants
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DADA_BRINGUP -o /tmp/libada_bringup.so amodal-depth-anything_b200/csrc/ada_api.cu || exit 1
for cfg in ${CFGS:-"2 0 6" "2 0 0" "0 1 6"}; do
  set -- $cfg
  echo "===== wait=$1 stagger=$2 emu=$3"
  ADA_B200_LIB=/tmp/libada_bringup.so ADA_ATT_IMPL=1 ADA_ATT_WAIT=$1 ADA_ATT_STAGGER=$2 ADA_ATT_EMU=$3 python tools/att2_timeline.py
done
