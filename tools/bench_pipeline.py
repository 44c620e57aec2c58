"""Latency of the device-resident single-image pipeline (AmodalInference = infer.py:72-103 numeric path) with the
reference's real model pair: un-guided ViT-G (features 384, out_channels 1536 x4, infer.py:59) + guided ViT-L, 518x518,
random-init weights, one image per call. Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amodal_depth_anything_b200 as pkg

enc_raw = os.environ.get("RAW", "vitg")
cfg = pkg.MODEL_CONFIGS[enc_raw]
torch.manual_seed(0)
raw = pkg.DepthAnythingV2(encoder=enc_raw, features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
am = pkg.AmodalDAv2(guide_type="mask+observation", encoder="vitl", pretrained=False).cuda().eval()
with torch.no_grad():
    am.encoder.pretrained.patch_embed_guidance.proj.weight.normal_(std=0.02)
    # random-init head: keep the un-guided output off the ReLU floor, otherwise min == max and the reference's min-max
    # normalisation (infer.py:22) is 0/0
    getattr(raw.depth_head.scratch.output_conv2, "2").bias.fill_(1.0)
raw.repack()
pipe = pkg.AmodalInference(raw, am, cuda_graph=bool(int(os.environ.get('GRAPH', '1'))))
rng = np.random.default_rng(0)
img = rng.integers(0, 256, size=(518, 518, 3), dtype=np.uint8)
mask = np.zeros((518, 518), np.uint8); mask[100:400, 150:420] = 255
for _ in range(3):
    out = pipe(img, mask)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    out = pipe(img, mask)
    agg = out["depth_agg"].cpu()      # the caller's read of the result, as infer.py:105 does
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
assert torch.isfinite(agg).all()
print(json.dumps({"cuda_graph": os.environ.get("GRAPH", "1"), "pipeline": f"{enc_raw} un-guided + vitl guided, 518x518, 1 image", "ms_per_image_device": e0.elapsed_time(e1) / n,
                  "ms_per_image_wall": (t1 - t0) / n * 1e3, "launches_raw": raw.launch_count(), "launches_amodal": am.launch_count()}))
