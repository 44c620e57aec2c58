"""Times the attention kernel alone (ViT-L shape) through the C ABI; ADA_ATT_VARIANT selects measurement variants."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amodal_depth_anything_b200  # noqa
from amodal_depth_anything_b200 import ops
B, N, H = int(os.environ.get("B", 32)), int(os.environ.get("N", 1370)), 16
IMPL = int(os.environ.get("ADA_ATT_IMPL", -1))
qkv = (torch.randn(B, N, 3, H, 64, device="cuda") * 1.0).bfloat16()
for _ in range(3):
    ops.attention(qkv, B, N, H, IMPL)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(qkv, B, N, H, IMPL)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 4.0 * B * H * N * N * 64
print(json.dumps({"impl": os.environ.get("ADA_ATT_IMPL", "0"), "emu": os.environ.get("ADA_ATT_EMU", ""), "stagger": os.environ.get("ADA_ATT_STAGGER", ""), "B": B, "N": N, "ms": ms, "tflops": fl / ms / 1e9}))
