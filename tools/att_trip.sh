#!/bin/bash
# One GPU-box visit for the attention kernels: operator parity for both implementations, stand-alone timing, optional ncu.
mkdir -p gpurun_out
for impl in 1 ${PARITY_OLD:+0}; do
  echo "=== parity ADA_ATT_IMPL=$impl"
  ADA_ATT_IMPL=$impl timeout 900 python tools/gpu_check.py --only attention --out gpurun_out/att_check_impl$impl.json 2>&1 | tail -12
done
echo "=== timing (B=32, N=1370, 16 heads)"
ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
for st in 1 0; do for emu in ${EMUS:-0 2 3 4 5 6}; do ADA_ATT_STAGGER=$st ADA_ATT_IMPL=1 ADA_ATT_EMU=$emu timeout 120 python tools/bench_attention.py; done; done
echo "=== timing (B=4, N=5477)"
B=4 N=5477 ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
for emu in ${EMUS:-0 2 3 4 5 6}; do B=4 N=5477 ADA_ATT_IMPL=1 ADA_ATT_EMU=$emu timeout 120 python tools/bench_attention.py; done
echo "=== timing (B=4 and B=8, N=1370: small grids)"
for b in 4 8; do
B=$b N=1370 ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
B=$b N=1370 ADA_ATT_IMPL=1 ADA_ATT_EMU=${BEST_EMU:-4} timeout 120 python tools/bench_attention.py
done
if [ "$1" = "ncu" ]; then
  echo "=== ncu attention_fa"
  ADA_ATT_IMPL=1 ADA_ATT_EMU=${BEST_EMU:-4} timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fa -s 3 -c 1 -o gpurun_out/r02_att2 -f python tools/bench_attention.py > gpurun_out/r02_att2_ncu.log 2>&1; tail -2 gpurun_out/r02_att2_ncu.log
fi
