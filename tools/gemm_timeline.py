"""clock64 timeline of one epilogue warp (CTA 0) for a GEMM shape: python tools/gemm_timeline.py M N K"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ADA_GEMM_TIMELINE"] = "1"
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import ops, _lib as L
M, N, K = [int(v) for v in sys.argv[1:4]]
A = torch.randn(M, K, device="cuda").bfloat16()
W = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.gemm(A, W, out_bf16=out, ldo=N)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.gemm(A, W, out_bf16=out, ldo=N); e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
L.check(L.load().ada_debug_timeline(buf, 512))
t = list(buf)
print("kernel ms", e0.elapsed_time(e1), "TF/s", 2.0 * M * N * K / e0.elapsed_time(e1) / 1e9)
names = ["top", "rowmap", "tfull", "math", "wait_read", "sts+fence", "store"]
base = t[0]
for j in range(2, 14):
    r = t[j * 8:j * 8 + 7]
    d = [r[0] - base] + [r[k] - r[k - 1] for k in range(1, 7)]
    print(f"tile {j:2d}: start {d[0]:7d}  " + "  ".join(f"{names[k]}+{d[k]:5d}" for k in range(1, 7)))
