#!/bin/bash
# Round-end verification on one B200: GPU test-suite, smoke, the default bench line, then the ncu evidence.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench (default)"; T0=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench wall $(( $(date +%s)-T0 )) s"; tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
if [ "$1" = "prof" ]; then bash tools/final_profile.sh; fi
