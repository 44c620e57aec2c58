#!/bin/bash
# Round-end evidence: ncu launch list of the bench command + ncu --set full of the dominant kernels (B=32 workload).
# The .ncu-rep files are summarised ON THE BOX (tools/ncu_summary.py -> gpurun_out/${R}_ncu_*.txt) and deleted unless KEEP=1:
# together they exceed the 64 MiB that travel back. Sections: launches gemm att att2 ln head tail (default: all).
R=${R:-r02}
SECTIONS=${@:-launches gemm att att2 ln head tail}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --profile-steps 0"
summ() {  # summ <name> [kernel ids for the per-instruction stall table]
  local n=$1; shift
  python tools/ncu_summary.py gpurun_out/${R}_$n.ncu-rep "$@" > gpurun_out/${R}_ncu_$n.txt 2>&1
  [ "$KEEP" = "1" ] || rm -f gpurun_out/${R}_$n.ncu-rep
  head -3 gpurun_out/${R}_ncu_$n.txt | cut -c1-160
}
for s in $SECTIONS; do
case $s in
launches)
  echo "=== launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/${R}_launches.log 2>&1; tail -1 gpurun_out/${R}_launches.log | cut -c1-200
  python tools/ncu_launch_summary.py gpurun_out/${R}_launches.csv > gpurun_out/${R}_ncu_launch_summary.txt 2>&1; head -12 gpurun_out/${R}_ncu_launch_summary.txt ;;
gemm)
  echo "=== full: gemm (one encoder block: qkv, proj, fc1, fc2)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 142 -c 4 -o gpurun_out/${R}_gemm -f $B > gpurun_out/${R}_gemm.log 2>&1; tail -1 gpurun_out/${R}_gemm.log | cut -c1-200
  summ gemm ;;
att)
  echo "=== full: attention (product kernel, in the model)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 30 -c 1 -o gpurun_out/${R}_att -f $B > gpurun_out/${R}_att.log 2>&1; tail -1 gpurun_out/${R}_att.log | cut -c1-200
  summ att 1 ;;
att2)
  echo "=== full: attention2 (alternative kernel, all exponentials on MUFU, stand-alone)"
  ADA_ATT_IMPL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fa -s 3 -c 1 -o gpurun_out/${R}_att2 -f python tools/bench_attention.py > gpurun_out/${R}_att2.log 2>&1; tail -1 gpurun_out/${R}_att2.log | cut -c1-200
  summ att2 1 ;;
ln)
  echo "=== full: layernorm (norm2 and norm1 of one block)"
  timeout 900 ncu --set full --clock-control none -k regex:layernorm_rows -s 60 -c 2 -o gpurun_out/${R}_ln -f $B > gpurun_out/${R}_ln.log 2>&1; tail -1 gpurun_out/${R}_ln.log | cut -c1-200
  summ ln ;;
head)
  echo "=== full: head kernels (tensor-core tail, upsample, channel LN)"
  timeout 900 ncu --set full --clock-control none -k regex:"tail_mma|tail_gather|upsample_bilinear|channel_ln" -s 32 -c 8 -o gpurun_out/${R}_head -f $B > gpurun_out/${R}_head.log 2>&1; tail -1 gpurun_out/${R}_head.log | cut -c1-200
  summ head ;;
tail)
  echo "=== full: tensor-core tail stand-alone"
  MODE=mma ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_mma -s 2 -c 1 -o gpurun_out/${R}_tail_mma -f python tools/bench_tail.py > gpurun_out/${R}_tail_mma.log 2>&1; tail -1 gpurun_out/${R}_tail_mma.log | cut -c1-200
  summ tail_mma 1 ;;
rcu)
  echo "=== full: head 3x3 convs with residual adds (RCU conv2, EPI 8) and with a ReLU epilogue (RCU conv1, EPI 7), first forward"
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tcgen05_kernel<.int.256, .int.2, .int.8>" -c 7 -o gpurun_out/${R}_rcu_resid -f $B > gpurun_out/${R}_rcu_resid.log 2>&1; tail -1 gpurun_out/${R}_rcu_resid.log | cut -c1-200
  summ rcu_resid 6 7
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tcgen05_kernel<.int.256, .int.2, .int.7>" -c 7 -o gpurun_out/${R}_rcu_relu -f $B > gpurun_out/${R}_rcu_relu.log 2>&1; tail -1 gpurun_out/${R}_rcu_relu.log | cut -c1-200
  summ rcu_relu 7 ;;
esac
done
du -sh gpurun_out
