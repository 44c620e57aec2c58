#!/bin/bash
mkdir -p gpurun_out
MODE=taps timeout 120 python tools/bench_tail.py
MODE=mma timeout 120 python tools/bench_tail.py
MODE=mma ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_mma -s 2 -c 1 -o gpurun_out/r02_tail_mma -f python tools/bench_tail.py > gpurun_out/r02_tail_mma_ncu.log 2>&1; tail -2 gpurun_out/r02_tail_mma_ncu.log
