#!/bin/bash
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline $2 --detail gpurun_out/detail_$3.json 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'img/s %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), {k:round(v['ms_per_step'],2) for k,v in d['breakdown'].items()}, d['clocks']['sm_mhz'])"; }
for b in 1 2 4 8 16; do
ADA_GEMM_TILE_MODEL=0 run "b$b old" "--batch $b --graph" b${b}o
run "b$b new" "--batch $b --graph" b${b}n
done
run "vitb b8 new" "--encoder vitb --batch 8 --graph" vitb
ADA_GEMM_TILE_MODEL=0 run "vitb b8 old" "--encoder vitb --batch 8 --graph" vitbo
python - <<'PY'
import json
for tag in ("b4n",):
    print("==", tag)
    for r in json.load(open(f"gpurun_out/detail_{tag}.json"))[:12]:
        print(f"{r['ms_per_step']:8.3f} ms/step x{r['launches']:3d} avg {r['avg_ms']:.4f} ms {r['tflops'] or 0:7.1f} TF/s  {r['sig']}")
PY
