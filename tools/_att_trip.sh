timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ADA_RESID_EPI=1 timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "oracle" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_w.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w.json').read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:(round(v['ms_per_step'],2)) for k,v in d['breakdown'].items()}, d['clocks'])
PY
