#!/bin/bash
# Round-end evidence: ncu launch list of the bench command + ncu --set full of the dominant kernels (B=32 workload).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-steps 0"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_launches.log 2>&1; tail -1 gpurun_out/final_launches.log | cut -c1-200
echo "=== full: gemm (one encoder block: qkv, proj, fc1, fc2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 142 -c 4 -o gpurun_out/final_gemm -f $B > gpurun_out/final_gemm.log 2>&1; tail -1 gpurun_out/final_gemm.log | cut -c1-200
echo "=== full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 30 -c 1 -o gpurun_out/final_att -f $B > gpurun_out/final_att.log 2>&1; tail -1 gpurun_out/final_att.log | cut -c1-200
echo "=== full: layernorm"
timeout 900 ncu --set full --clock-control none -k regex:layernorm_rows -s 60 -c 1 -o gpurun_out/final_ln -f $B > gpurun_out/final_ln.log 2>&1; tail -1 gpurun_out/final_ln.log | cut -c1-200
ls -la gpurun_out/final_*
