"""Dumps the per-phase clock64 timeline of one softmax thread (ADA_ATT_VARIANT=10)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ADA_ATT_VARIANT"] = "10"
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import ops, _lib as L
B, N, H = 32, 1370, 16
qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
for _ in range(3):
    ops.attention(qkv, B, N, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention(qkv, B, N, H); e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
L.check(L.load().ada_debug_timeline(buf, 512))
t = list(buf)
print("kernel ms", e0.elapsed_time(e1), "pad", os.environ.get("ADA_ATT_PAD", "0"))
names = ["loop_top", "s_full", "ldtm", "max+xchg", "exps", "o_full", "sttm"]
base = t[0]
print("softmax thread 64 (one of the two owners of a row), cycles per phase and 128-key tile")
for j in range(11):
    r = t[j * 8:j * 8 + 7]
    d = [r[0] - base] + [r[k] - r[k - 1] for k in range(1, 7)]
    print(f"tile {j:2d}: start {d[0]:7d}  " + "  ".join(f"{names[k]}+{d[k]:5d}" for k in range(1, 7)))
inames = ["top", "s_free", "S(j+1) issued", "p_full", "PV(j) issued"]
print("issuer thread, absolute cycles")
for j in range(11):
    r = t[128 + j * 8:128 + j * 8 + 5]
    print(f"issuer {j:2d}: " + "  ".join(f"{inames[k]}@{r[k] - base:6d}" for k in range(5) if r[k]))
