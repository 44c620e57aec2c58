#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_check.py --only attention0 --out gpurun_out/att_check_impl0.json 2>&1 | tail -11
for shape in "32 1370" "4 5477" "8 1370" "4 1370" "1 1370"; do set -- $shape
  B=$1 N=$2 ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
  B=$1 N=$2 ADA_ATT_IMPL=1 timeout 120 python tools/bench_attention.py
done
inmodel() { timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']['attention_tcgen05']
print('$1', 'img/s %.1f  ms/step %.2f  attention ms/step %.2f (%.0f TFLOP/s)  clocks %s' % (d['value'], d['ms_per_step'], b['ms_per_step'], b['tflops'], d['clocks']['sm_mhz']))"; }
ADA_ATT_IMPL=0 inmodel impl0 ""
ADA_ATT_IMPL=1 inmodel impl1 ""
ADA_ATT_IMPL=0 inmodel "impl0 1036" "--size 1036 --batch 4"
ADA_ATT_IMPL=1 inmodel "impl1 1036" "--size 1036 --batch 4"
