for c in gemm_resid_f32_small gemm_resid_f32_ragged gemm_resid_f32 gemm_resid_f32_cg2; do timeout 120 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "$c" 2>&1 | tail -3; done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --detail gpurun_out/detail.json 2>&1 | tail -1 > gpurun_out/bench_w.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w.json').read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:(round(v['ms_per_step'],2)) for k,v in d['breakdown'].items()}, d['clocks'])
for r in json.load(open("gpurun_out/detail.json"))[:10]:
    print(f"{r['ms_per_step']:8.3f} ms/step  x{r['launches']:3d}  avg {r['avg_ms']:.3f} ms  {r['tflops'] or 0:7.1f} TF/s  {r['sig']}")
PY
