#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py --only tail_mma --out gpurun_out/tail_mma_check.json 2>&1 | tail -7
for mma in 1 0; do ADA_TAIL_MMA=$mma timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ADA_TAIL_MMA=$mma', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', {k: round(v['ms_per_step'],3) for k,v in d['breakdown'].items()})"; done
