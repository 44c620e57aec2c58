#!/bin/bash
# Round-end evidence: ncu launch list of the bench command + ncu --set full of the dominant kernels (B=32 workload).
R=${R:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --profile-steps 0"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/${R}_launches.log 2>&1; tail -1 gpurun_out/${R}_launches.log | cut -c1-200
echo "=== full: gemm (one encoder block: qkv, proj, fc1, fc2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 142 -c 4 -o gpurun_out/${R}_gemm -f $B > gpurun_out/${R}_gemm.log 2>&1; tail -1 gpurun_out/${R}_gemm.log | cut -c1-200
echo "=== full: attention (product kernel, in the model)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 30 -c 1 -o gpurun_out/${R}_att -f $B > gpurun_out/${R}_att.log 2>&1; tail -1 gpurun_out/${R}_att.log | cut -c1-200
echo "=== full: attention2 (alternative kernel, all exponentials on MUFU, stand-alone)"
ADA_ATT_IMPL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fa -s 3 -c 1 -o gpurun_out/${R}_att2 -f python tools/bench_attention.py > gpurun_out/${R}_att2.log 2>&1; tail -1 gpurun_out/${R}_att2.log | cut -c1-200
echo "=== full: layernorm"
timeout 900 ncu --set full --clock-control none -k regex:layernorm_rows -s 60 -c 1 -o gpurun_out/${R}_ln -f $B > gpurun_out/${R}_ln.log 2>&1; tail -1 gpurun_out/${R}_ln.log | cut -c1-200
echo "=== full: head kernels (tensor-core tail, upsample, channel LN)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tail_mma|tail_gather|upsample_bilinear|channel_ln" -s 32 -c 8 -o gpurun_out/${R}_head -f $B > gpurun_out/${R}_head.log 2>&1; tail -1 gpurun_out/${R}_head.log | cut -c1-200
echo "=== full: tensor-core tail stand-alone"
MODE=mma ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_mma -s 2 -c 1 -o gpurun_out/${R}_tail_mma -f python tools/bench_tail.py > gpurun_out/${R}_tail_mma.log 2>&1; tail -1 gpurun_out/${R}_tail_mma.log | cut -c1-200
ls -la gpurun_out/${R}_*
