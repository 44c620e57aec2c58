#!/bin/bash
# One GPU-box visit for the tensor-core tail: operator parity (both descriptor field orders on the first visit), model parity, timing.
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py --only tail_mma --out gpurun_out/tail_mma_check.json 2>&1 | tail -12
echo "=== model parity (golden + bench-path tests)"
timeout 1500 python -m pytest tests/test_forward_gpu.py tests/test_bench_paths_gpu.py tests/test_raw_gpu.py -x -q 2>&1 | tail -5
echo "=== bench"
for mma in 0 1; do ADA_TAIL_MMA=$mma timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ADA_TAIL_MMA=$mma', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', {k: round(v['ms_per_step'],3) for k,v in d['breakdown'].items()})"; done
