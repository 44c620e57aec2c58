timeout 900 python -m pytest tests -x -q -m gpu -k "vitg or swiglu or raw" 2>&1 | tail -4
timeout 600 python bench.py --encoder vitg --size 518 --batch 8 --no-cpu-baseline --detail gpurun_out/detail_g.json 2>&1 | tail -1 > gpurun_out/bench_g.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_g.json').read())
print(d['value'], d['ms_per_step'], 'frac', d['model_frac_of_peak'], {k:(round(v['ms_per_step'],2)) for k,v in d['breakdown'].items()})
for r in json.load(open("gpurun_out/detail_g.json"))[:8]:
    print(f"{r['ms_per_step']:8.3f} ms/step  x{r['launches']:3d}  avg {r['avg_ms']:.3f} ms  {r['tflops'] or 0:7.1f} TF/s  {r['sig']}")
PY
