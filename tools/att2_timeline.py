"""Bring-up: clock64 timeline of the attention2 kernel (needs a library built with -DADA_BRINGUP, see tools/att_timeline.sh).
Prints, per KV tile of the first two work units of CTA 0, the phase durations of one thread of each softmax group and the
issuer's three stamps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import _lib as L, ops
B, N, H = 32, 1370, 16
qkv = (torch.randn(B, N, 3, H, 64, device="cuda")).bfloat16()
for _ in range(3):
    ops.attention(qkv, B, N, H)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
L.check(L.load().ada_debug_timeline(buf, 512))
v = list(buf)
t0 = min(x for x in v[:374] if x > 0)
names = ["wait_S", "ldtm", "max", "wait_O", "math", "st+arrive"]
for t in (0, 1):
    print(f"--- group {t}: start(rel) | " + " ".join(f"{n:>9s}" for n in names) + " | tile period")
    prev = None
    for j in range(22):
        s = v[t * 154 + j * 7: t * 154 + j * 7 + 7]
        if not s[0]:
            continue
        d = [s[k + 1] - s[k] for k in range(6)]
        per = s[0] - prev if prev else 0
        prev = s[0]
        print(f"tile {j:2d}: {s[0] - t0:8d} | " + " ".join(f"{x:9d}" for x in d) + f" | {per}")
print("--- issuer: iteration start(rel) | issue_s | issue_pv")
for i in range(22):
    s = v[308 + 3 * i: 308 + 3 * i + 3]
    if s[0]:
        print(f"it {i:2d}: {s[0] - t0:8d} | {s[1] - s[0]:8d} | {s[2] - s[1]:8d}")
