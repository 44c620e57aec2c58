"""Does the clock sampler disturb the timed loop? Per-step CUDA-event times of the bench workload with no sampler, with the
nvidia-smi -lms sampler of bench.py, and with an in-process NVML sampler."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import amodal_depth_anything_b200 as pkg

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = bench.make_model(pkg, torch, "vitl", dev)
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
    print(f"{label:28s} mean {sum(ms)/n:6.2f}  " + " ".join(f"{m:5.1f}" for m in ms), flush=True)


def per_step(n, label):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(n):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    print(f"{label:28s} mean {sum(ms)/n:6.2f}  median {ms[n//2]:6.2f}  max3 " + " ".join(f"{m:5.1f}" for m in ms[-3:]), flush=True)


N = int(os.environ.get("N", "30"))
for _ in range(3):
    step()
uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
for rep in range(2):
    per_step(N, "no sampler")
    s = bench.NvmlSampler(uuid)
    s.start()
    time.sleep(0.3)
    s.mark_begin()
    per_step(N, "in-process NVML 50 ms")
    s.mark_end()
    print("   ", s.stop(), flush=True)
    per_step(N, "no sampler")
    s = bench.ClockSampler(uuid)
    s.start()
    time.sleep(1.0)
    s.mark_begin()
    per_step(N, "nvidia-smi -lms 100")
    s.mark_end()
    print("   ", s.stop(), flush=True)
