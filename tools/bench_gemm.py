"""Times one GEMM shape through the C ABI: [BN=.. CG=..] python tools/bench_gemm.py M N K [gelu]; BN / CG force the tile."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import ops, _lib as L
M, N, K = [int(v) for v in sys.argv[1:4]]
act = L.ACT_GELU if len(sys.argv) > 4 and sys.argv[4] == "gelu" else L.ACT_NONE
A = torch.randn(M, K, device="cuda").bfloat16()
W = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
BN, CG = int(os.environ.get("BN", 0)), int(os.environ.get("CG", 0))
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.gemm(A, W, bias=bias, out_bf16=out, ldo=N, act=act, force_bn=BN, force_cg=CG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm(A, W, bias=bias, out_bf16=out, ldo=N, act=act, force_bn=BN, force_cg=CG)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"M": M, "N": N, "K": K, "bn": BN, "cg": CG, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}))
