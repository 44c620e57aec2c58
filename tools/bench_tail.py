"""Times the tail of the DPT head alone (ViT-L shape: 128 channels, 296^2 -> 518^2) through the C ABI:
MODE=mma  : tail_mma_kernel (upsample + output_conv2 on tensor cores, one kernel)
MODE=taps : tap GEMM (N = 288, fp16) + tail_gather_kernel (the older pair)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import ops, _lib as L
B, G, C = int(os.environ.get("B", 32)), int(os.environ.get("G", 37)), int(os.environ.get("C", 128))
MODE = os.environ.get("MODE", "mma")
Hl, H = 8 * G, 14 * G
g = torch.Generator(device="cuda").manual_seed(3)
y = torch.randn(B, Hl, Hl, C, generator=g, device="cuda")
w2 = torch.randn(32, C, 3, 3, generator=g, device="cuda") * (1.0 / (3 * C ** 0.5))
b2 = torch.randn(32, generator=g, device="cuda") * 0.1
aux = torch.randn(33, generator=g, device="cuda") * 0.3
if MODE == "mma":
    yh, wpk = y.half(), ops.pack_tail_mma(w2)
    run = lambda: ops.tail_mma(yh, wpk, b2, aux, H, H, 1)
else:
    yb, wt = y.bfloat16().reshape(B * Hl * Hl, C), ops.pack_tail_taps(w2)
    V = torch.zeros(B, Hl, Hl, 288, dtype=torch.float16, device="cuda")
    def run():
        ops.gemm(yb, wt, epi=L.EPI_F16, out_bf16=V, ldo=288)
        return ops.tail_gather(V, b2, aux, H, H, True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("ITERS", 10))
e0.record()
for _ in range(n):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"mode": MODE, "B": B, "grid": G, "C": C, "ms": ms, "conv_tflops": 2.0 * B * H * H * 32 * 9 * C / ms / 1e9}))
