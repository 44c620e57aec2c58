import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
iv = lambda x: int(x) if x.isdigit() else 0
stalls = [h for h in hdr if h.startswith("stall_")]
tot = sum(iv(r[ix["# Samples"]]) for r in data)
agg = {s: sum(iv(r[ix[s]]) for r in data) for s in stalls}
print("total", tot, sorted(agg.items(), key=lambda kv: -kv[1])[:10])
# print in program order the instructions with samples > threshold
thr = tot * 0.004
for i, r in enumerate(data):
    n = iv(r[ix["# Samples"]])
    if n >= thr:
        st = sorted(((iv(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"{i:5d} {n:6d} {r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
