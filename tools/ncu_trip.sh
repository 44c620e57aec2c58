#!/bin/bash
# ncu --set full captures of the attention kernel and the fc1/fc2 GEMMs (batch 8, first forward). Outputs in gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --profile-steps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 6 -c 1 -o gpurun_out/prof_att -f $B > gpurun_out/ncu_att.log 2>&1; tail -2 gpurun_out/ncu_att.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 23 -c 2 -o gpurun_out/prof_fc -f $B > gpurun_out/ncu_fc.log 2>&1; tail -2 gpurun_out/ncu_fc.log
ls -la gpurun_out/*.ncu-rep
