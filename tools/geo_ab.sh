#!/bin/bash
# A/B of the general conv pixel-tile shapes (ADA_CONV_GEO) on the bench workload, interleaved on one box.
mkdir -p gpurun_out
for c in ${CASES:-0 1 0 1}; do
  ADA_CONV_GEO=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/geo_ab_$c.json 2>/dev/null
  python - "$c" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/geo_ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
b = d["breakdown"]
st = d.get("ms_steps_rank0") or [0]
print(f"geo {sys.argv[1]}: {d['ms_per_step']:.2f} ms/step (max {max(st):.1f})  e2e {d['e2e']['ms_per_step']:.2f}  conv {b['gemm_tcgen05_conv3x3']['ms_per_step']:.2f}  "
      f"linear {b['gemm_tcgen05_linear']['ms_per_step']:.2f}  att {b['attention_tcgen05']['ms_per_step']:.2f}  LN {b['layernorm']['ms_per_step']:.2f}  clk {d['clocks']['sm_mhz']}", flush=True)
PY
done
