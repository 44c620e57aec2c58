"""Do single steps of the timed loop stall, and does the clock sampler have a part in it? Per-step CUDA-event times of the
bench workload (graph replay, like bench.py) in alternating segments with and without the in-process NVML sampler."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import amodal_depth_anything_b200 as pkg

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = bench.make_model(pkg, torch, "vitl", dev)
model.set_graph(os.environ.get("GRAPH", "1") == "1")
x, mask, obs = bench.make_inputs(torch, 32, 518, 518, dev, 1234)
step = lambda: model(x, guide_rgb=None, guide_mask=mask, observation=obs)


def per_step(n, label):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(n):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    srt = sorted(ms)
    slow = [(i, round(m, 1)) for i, m in enumerate(ms) if m > 1.15 * srt[n // 2]]
    print(f"{label:26s} mean {sum(ms)/n:6.2f}  median {srt[n//2]:6.2f}  max {srt[-1]:6.1f}  slow steps {slow}", flush=True)


N = int(os.environ.get("N", "30"))
for _ in range(5):
    step()
uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
for rep in range(int(os.environ.get("REPS", "3"))):
    per_step(N, "no sampler")
    s = bench.NvmlSampler(uuid)
    s.start()
    s.mark_begin()
    per_step(N, "in-process NVML sampler")
    s.mark_end()
    print("   ", s.stop(), flush=True)
