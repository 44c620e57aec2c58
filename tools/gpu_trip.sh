#!/bin/bash
# One GPU-box visit: (smoke), GPU tests, bench, optional ncu passes. Outputs land in gpurun_out/.
# usage: gpu_trip.sh [tests|notests] [ncu|noncu] [bench args...]
mkdir -p gpurun_out
T=${1:-tests}; N=${2:-noncu}; shift; shift
nvidia-smi -L
if [ "$T" = "tests" ]; then
  echo "=== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
elif [ "$T" = "quick" ]; then
  echo "=== pytest gpu (quick)"; timeout 900 python -m pytest tests -x -q -m gpu -k "golden or kernel" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
fi
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --detail gpurun_out/detail.json "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/detail.json"))[:28]:
        print(f"{r['ms_per_step']:8.3f} ms/step  x{r['launches']:3d}  avg {r['avg_ms']:.3f} ms  {r['tflops'] or 0:7.1f} TF/s  {r['sig']}")
except Exception as e:
    print("no detail", e)
PY
if [ "$N" = "ncu" ]; then
  echo "=== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
  echo "=== ncu full (gemm)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 40 -c 4 -o gpurun_out/prof_gemm -f \
     python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
  ls -la gpurun_out/
fi
