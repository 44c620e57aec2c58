#!/bin/bash
# A/B of the L2 carry-over knobs on the bench workload, interleaved on one box: "LN_FLAGS:ASTREAM" pairs.
mkdir -p gpurun_out
for c in ${CASES:-0:0 3:0 3:3 7:3 0:0 3:0 3:3 7:3}; do
  f=${c%%:*}; s=${c##*:}
  ADA_LN_FLAGS=$f ADA_GEMM_ASTREAM=$s timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/l2_ab.json 2>/dev/null
  python - "$c" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/l2_ab.json").read().strip().splitlines()[-1])
b = d["breakdown"]
st = d.get("ms_steps_rank0") or []
print(f"ln:astream {sys.argv[1]}: {d['ms_per_step']:.2f} ms/step (max step {max(st) if st else 0:.1f})  e2e {d['e2e']['ms_per_step']:.2f}  "
      f"LN {b['layernorm']['ms_per_step']:.3f}  linear {b['gemm_tcgen05_linear']['ms_per_step']:.2f}  "
      f"att {b['attention_tcgen05']['ms_per_step']:.2f}  conv {b['gemm_tcgen05_conv3x3']['ms_per_step']:.2f}  clk {d['clocks']['sm_mhz']}", flush=True)
PY
done
