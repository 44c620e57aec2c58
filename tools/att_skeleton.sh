#!/bin/bash
# attention2 timing skeletons (bring-up build): where is the floor of the pipeline without the softmax arithmetic?
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DADA_BRINGUP -o /tmp/libada_bringup.so amodal-depth-anything_b200/csrc/ada_api.cu 2>&1 | grep -i "error" 
export ADA_B200_LIB=/tmp/libada_bringup.so ADA_ATT_IMPL=1
for emu in 0 6 -1 -2; do for w in 2 0; do
  echo -n "emu=$emu wait=$w: "; ADA_ATT_EMU=$emu ADA_ATT_WAIT=$w ADA_ATT_STAGGER=0 python tools/bench_attention.py
done; done
echo "old kernel skeleton (variant 2 = exps replaced by a copy):"
ADA_ATT_IMPL=0 ADA_ATT_VARIANT=2 python tools/bench_attention.py
ADA_ATT_IMPL=0 ADA_ATT_VARIANT=1 python tools/bench_attention.py
ADA_ATT_IMPL=0 ADA_ATT_VARIANT=0 python tools/bench_attention.py
