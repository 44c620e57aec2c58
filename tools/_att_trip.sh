for rep in 1 2; do for v in 0 4 5; do
ADA_ATT_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --steps 6 2>&1 | tail -1 > gpurun_out/bench_w.json; python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_w.json').read())
print('ATT', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), 'att', round(d['breakdown']['attention_tcgen05']['ms_per_step'],2), d['clocks']['sm_mhz'] if d['clocks'] else None)
PY
done; done
