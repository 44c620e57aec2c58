#!/bin/bash
# Attention v2 design sweep: wait placement x stagger x exponential split, stand-alone timing at the bench shape, then in-model.
mkdir -p gpurun_out
ADA_ATT_IMPL=1 timeout 600 python tools/gpu_check.py --only attention --out gpurun_out/att_check_impl1.json 2>&1 | tail -10
ADA_ATT_IMPL=0 timeout 120 python tools/bench_attention.py
for w in ${WAITS:-0 2}; do for st in 0 1; do for emu in ${EMUS:-0 4 6}; do
  echo -n "wait=$w "; ADA_ATT_WAIT=$w ADA_ATT_STAGGER=$st ADA_ATT_IMPL=1 ADA_ATT_EMU=$emu timeout 120 python tools/bench_attention.py
done; done; done
echo "=== in-model (bench.py, batch 32)"
inmodel() { timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']['attention_tcgen05']
print('$1', 'img/s %.1f  ms/step %.2f  attention ms/step %.2f (%.0f TFLOP/s)  clocks %s' % (d['value'], d['ms_per_step'], b['ms_per_step'], b['tflops'], d['clocks']['sm_mhz']))"; }
ADA_ATT_IMPL=0 inmodel "impl0"
for cfg in ${INMODEL:-"0 0 0" "0 1 6" "2 0 0" "2 1 6"}; do set -- $cfg
  ADA_ATT_IMPL=1 ADA_ATT_WAIT=$1 ADA_ATT_STAGGER=$2 ADA_ATT_EMU=$3 inmodel "impl1 wait=$1 stagger=$2 emu=$3"
done
