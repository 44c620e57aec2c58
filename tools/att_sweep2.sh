#!/bin/bash
for shape in "4 5477" "8 1370" "4 1370" "1 1370"; do set -- $shape
  echo "=== B=$1 N=$2"
  B=$1 N=$2 ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
  for cfg in "2 0 0" "2 0 4" "0 0 4" "2 1 6"; do set -- $shape; b=$1; n=$2; set -- $cfg
    B=$b N=$n ADA_ATT_WAIT=$1 ADA_ATT_STAGGER=$2 ADA_ATT_IMPL=1 ADA_ATT_EMU=$3 timeout 120 python tools/bench_attention.py
  done
done
inmodel() { timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']['attention_tcgen05']
print('$1', 'img/s %.1f  ms/step %.2f  attention ms/step %.2f (%.0f TFLOP/s)  clocks %s' % (d['value'], d['ms_per_step'], b['ms_per_step'], b['tflops'], d['clocks']['sm_mhz']))"; }
echo "=== in-model 1036 b4"
ADA_ATT_IMPL=0 inmodel impl0 "--size 1036 --batch 4"
ADA_ATT_IMPL=1 ADA_ATT_WAIT=2 ADA_ATT_STAGGER=0 ADA_ATT_EMU=0 inmodel "impl1 w2 s0 e0" "--size 1036 --batch 4"
ADA_ATT_IMPL=1 ADA_ATT_WAIT=2 ADA_ATT_STAGGER=0 ADA_ATT_EMU=4 inmodel "impl1 w2 s0 e4" "--size 1036 --batch 4"
echo "=== in-model 518 b32 again"
ADA_ATT_IMPL=0 inmodel impl0 ""
ADA_ATT_IMPL=1 ADA_ATT_WAIT=2 ADA_ATT_STAGGER=0 ADA_ATT_EMU=0 inmodel "impl1 w2 s0 e0" ""
ADA_ATT_IMPL=1 ADA_ATT_WAIT=2 ADA_ATT_STAGGER=0 ADA_ATT_EMU=4 inmodel "impl1 w2 s0 e4" ""
ADA_ATT_IMPL=1 ADA_ATT_WAIT=0 ADA_ATT_STAGGER=0 ADA_ATT_EMU=4 inmodel "impl1 w0 s0 e4" ""
