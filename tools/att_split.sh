#!/bin/bash
# One GPU-box visit for the two-threads-per-row persistent attention kernel (impl 2): parity + agreement, timings, in-model.
mkdir -p gpurun_out
timeout 900 python tools/gpu_check.py --only attention --out gpurun_out/att_check_split.json 2>&1 | tail -45
for cfg in "32 1370" "4 5477" "4 1370" "8 1370" "1 1370"; do set -- $cfg
  for impl in 0 1 2; do B=$1 N=$2 ADA_ATT_IMPL=$impl timeout 120 python tools/bench_attention.py; done
done
echo "=== in-model"
for impl in 0 2; do ADA_ATT_IMPL=$impl timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400; done
