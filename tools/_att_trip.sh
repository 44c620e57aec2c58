cp amodal-depth-anything_b200/libamodal_b200.so /tmp/lib_keep.so
for rep in 1 2; do for v in s1 s2; do
cp tools/ab/lib_$v.so amodal-depth-anything_b200/libamodal_b200.so
timeout 300 python bench.py --no-cpu-baseline --steps 8 --detail gpurun_out/detail.json 2>&1 | tail -1 > gpurun_out/bench_w.json; python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_w.json').read())
rows=json.load(open("gpurun_out/detail.json"))
def f(sub):
    r=[x for x in rows if sub in x['sig']]
    return round(r[0]['ms_per_step'],2) if r else None
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), 'lin', round(d['breakdown']['gemm_tcgen05_linear']['ms_per_step'],2), 'conv', round(d['breakdown']['gemm_tcgen05_conv3x3']['ms_per_step'],2), 'fc2', f('N=1024 K=4096'), 'proj', f('M=43840 N=1024 K=1024'), 'fc1', f('N=4096 K=1024'), 'qkv', f('N=3072 K=1024'), d['clocks']['sm_mhz'] if d['clocks'] else None)
PY
done; done
cp /tmp/lib_keep.so amodal-depth-anything_b200/libamodal_b200.so
