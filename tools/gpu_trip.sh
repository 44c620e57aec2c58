#!/bin/bash
# One GPU-box visit: smoke, GPU tests, bench, ncu launch list. Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
