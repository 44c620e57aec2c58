"""Per-kernel shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X): run anywhere, no GPU."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], errors="replace") if l.startswith('"')]
rd = csv.reader(rows)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"^void ", "", name)
    name = re.split(r"[<(]", name)[0].replace("ada::", "")
    if name.startswith("at::"):
        name = "at:: (torch, input synthesis / checks outside the forward)"
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 3 --no-extras "
      "(ViT-L 518x518, batch 32)")
print("per-launch times are cold-cache and serialised: compare SHARES")
print(f"total {tot:.1f} ms over {n} launches\n")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} launches {c:5d}  sum {ms:9.2f} ms  share {100 * ms / tot:5.1f} %  avg {1e3 * ms / c:8.1f} us")
