#!/bin/bash
inmodel() { timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']['attention_tcgen05']
print('$1', 'img/s %.1f  ms/step %.2f  attention ms/step %.2f (%.0f TFLOP/s)  clocks %s' % (d['value'], d['ms_per_step'], b['ms_per_step'], b['tflops'], d['clocks']['sm_mhz']))"; }
for rep in 1 2; do
ADA_ATT_IMPL=0 inmodel impl0 ""
ADA_ATT_IMPL=1 inmodel impl1 ""
done
ADA_ATT_IMPL=0 inmodel "impl0 1036" "--size 1036 --batch 4"
ADA_ATT_IMPL=1 inmodel "impl1 1036" "--size 1036 --batch 4"
ADA_ATT_IMPL=0 inmodel "impl0 vitg" "--encoder vitg --batch 8"
ADA_ATT_IMPL=1 inmodel "impl1 vitg" "--encoder vitg --batch 8"
