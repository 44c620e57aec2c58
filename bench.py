"""bench.py -- images/sec of the Amodal-DAv2 forward pass (BASELINE.json metric) on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W              # this repo's CUDA path
  torchrun ... bench.py --gpus N --steps K --warmup W        # N ranks, images sharded, no data-path collective
  python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host CPU cores (oracle port)

A step = one forward pass over one synthetic batch (ViT-L, 518x518, guide = mask+observation, 32 images per GPU).
`value` is device-timed (CUDA events on the launch stream, inputs resident in HBM); `e2e` goes through the public model
API with pinned host buffers, H2D and D2H inside the timed region. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic FLOPs per image (2*MAC over every GEMM / conv / QK^T / PV)
GFLOP_PER_IMAGE = {("vits", 518): 119.4, ("vitb", 518): 396.3, ("vitl", 518): 1389.6, ("vitl", 1036): 7764.3,
                   ("vitg", 518): 5771.4}
PROF_CLASSES = ["gemm_tcgen05_linear", "gemm_tcgen05_conv3x3", "attention_tcgen05", "layernorm", "channel_ln_relu",
                "upsample_bilinear", "gather"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.rows, self.proc, self.uuid = [], None, uuid

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(encoder, size, steps, warmup):
    """Times the oracle (CPU restatement of the reference forward) on all host cores: `steps` images, one per step."""
    import torch
    from oracle import amodal_oracle as O
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    gt = "mask+observation"
    sd = synth.make_state_dict(encoder, gt, 0)
    inp = synth.make_inputs(1, size, size, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward(sd, encoder, gt, inp["x"], None, inp["guide_mask"], inp["observation"])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return 1.0 / sec, sec, torch.get_num_threads()


def cpu_context():
    """SURVEY.md section 8d asks for the CPU model and the ViT-S 518x518 single-image time next to the baseline."""
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    _, sec_s, _ = cpu_reference_rate("vits", 518, 5, 1)
    return {"cpu_model": model, "os_cpu_count": os.cpu_count(), "vits_518_b1_s_per_image": sec_s}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, sec, cores = cpu_reference_rate(a.encoder, a.size, a.steps, a.warmup)
    sample = f"{a.steps} steps x 1 image ({a.encoder} {a.size}x{a.size}, fp32, torch CPU) after {a.warmup} warm-up"
    line = {
        "impl": "reference", "metric": "images/sec", "value": rate, "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a, world):
    per_gpu = a.batch if a.scaling == "weak" else max(a.batch // world, 1)
    in_mb = per_gpu * 5 * a.size * a.size * 4 / 1e6
    # fp32 residual stream alone: tokens x D x 4 B, rewritten every block
    d = {"vits": 384, "vitb": 768, "vitl": 1024, "vitg": 1536}[a.encoder]
    x_mb = per_gpu * ((a.size // 14) ** 2 + 1) * d * 4 / 1e6
    l2 = (f"per step {in_mb:.0f} MB of inputs and a {x_mb:.0f} MB fp32 token stream (plus GBs of bf16 activations) stream through "
          f"the 126 MB L2" if in_mb + x_mb > 126 else
          f"small-batch configuration: inputs ({in_mb:.0f} MB) and token stream ({x_mb:.0f} MB) fit the 126 MB L2 and are NOT "
          f"flushed between steps -- latency figure, not the headline metric")
    return {"workload": f"AmodalDAv2 {a.encoder} {a.size}x{a.size} guide=mask+observation forward, batch {per_gpu}/GPU "
                        f"(BASELINE.json configs[2])",
            "encoder": a.encoder, "height": a.size, "width": a.size, "per_gpu_batch": per_gpu,
            "global_batch": per_gpu * world, "parallelism": f"image-sharded x{world}, weights replicated, no collective",
            "cuda_graph": bool(getattr(a, "graph", False)), "l2": l2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--encoder", default="vitl")
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (weak) / total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the forward as a CUDA graph (launch-bound small batches)")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--detail", default="", help="write per-launch-signature timings of the profile pass to this JSON file")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    import amodal_depth_anything_b200 as pkg
    from amodal_depth_anything_b200.shard import max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config(a, world)
    B, H, W = cfg["per_gpu_batch"], a.size, a.size

    # random-init weights of the named architecture (no checkpoints offline); the guidance conv is zero-initialised by the
    # reference (dav2.py:55-61) -- randomise it so the guide path does real work
    torch.manual_seed(0)
    model = pkg.AmodalDAv2(guide_type="mask+observation", encoder=a.encoder, pretrained=False)
    with torch.no_grad():
        model.encoder.pretrained.patch_embed_guidance.proj.weight.normal_(std=0.02)
        model.encoder.pretrained.patch_embed_guidance.proj.bias.uniform_(-0.05, 0.05)
    model = model.to(dev).eval()
    if a.graph:
        model.set_graph(True)

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand(B, 3, H, W, device=dev, generator=g)
    low = torch.rand(B, 1, H // 37, W // 37, device=dev, generator=g)
    mask = (torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear") > 0.5).float() * 2 - 1
    obs = torch.rand(B, 1, H, W, device=dev, generator=g) * 2 - 1

    def step():
        return model(x, guide_rgb=None, guide_mask=mask, observation=obs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = step()
    barrier()
    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms, dev)
    launches = model.launch_count()
    assert torch.isfinite(out).all()

    # ---- end to end through the public API: pinned host inputs -> H2D -> forward -> D2H, every step
    hx, hm, ho = x.cpu().pin_memory(), mask.cpu().pin_memory(), obs.cpu().pin_memory()
    hout = torch.empty(B, 1, H, W).pin_memory()

    from amodal_depth_anything_b200.pipeline import StreamedInference
    runner = StreamedInference(model, dev)
    houts = [hout, torch.empty_like(hout).pin_memory()]

    def e2e_run(n):
        # every step: H2D of that step's pinned inputs, forward through the public model call, D2H of its result;
        # copies of neighbouring steps overlap the kernels (copy stream), all inside the timed region
        runner.run(((hx, hm, ho) for _ in range(n)), [houts[i & 1] for i in range(n)])

    e2e_run(2)
    barrier()
    e0.record()
    e2e_run(a.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / a.steps, dev)
    h2d = (hx.numel() + hm.numel() + ho.numel()) * 4
    d2h = hout.numel() * 4

    # ---- per-kernel-class CUDA-event breakdown (separate pass: the events add gaps, so it is not the headline number)
    import ctypes
    from amodal_depth_anything_b200 import _lib as L
    lib = L.load()
    breakdown, roof = {}, None
    peaks = load_peaks()
    if a.profile_steps > 0:
        lib.ada_set_profile(model._handle, 1)
        for _ in range(a.profile_steps):
            step()
        if a.detail and rank == 0:
            cap = 4096
            meta, rms = (ctypes.c_int32 * (5 * cap))(), (ctypes.c_double * cap)()
            nrec = lib.ada_profile_records(model._handle, cap, meta, rms)
            agg = {}
            for i in range(max(nrec, 0)):
                c, m_, n_, k_, tag = meta[5 * i:5 * i + 5]
                key = f"{PROF_CLASSES[c]} M={m_} N={n_} K={k_} epi={tag & 15} act={(tag >> 4) & 15} bn={(tag >> 8) & 0xfff} cg={tag >> 20}"
                e = agg.setdefault(key, {"ms": 0.0, "n": 0, "flops": 2.0 * m_ * n_ * k_})
                e["ms"] += rms[i]
                e["n"] += 1
            rows = []
            for key, e in agg.items():
                avg = e["ms"] / e["n"]
                rows.append({"sig": key, "launches": e["n"] // a.profile_steps, "avg_ms": avg,
                             "ms_per_step": e["ms"] / a.profile_steps,
                             "tflops": e["flops"] / avg / 1e9 if e["flops"] else None})
            rows.sort(key=lambda r: -r["ms_per_step"])
            json.dump(rows, open(a.detail, "w"), indent=1)
        n = len(PROF_CLASSES)
        msv, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
        ln = (ctypes.c_int32 * n)()
        L.check(lib.ada_profile_read(model._handle, n, msv, fl, by, ln))
        lib.ada_set_profile(model._handle, 0)
        tot = sum(msv) or 1.0
        for i, name in enumerate(PROF_CLASSES):
            if ln[i]:
                breakdown[name] = {"ms_per_step": msv[i] / a.profile_steps, "share": msv[i] / tot,
                                   "launches_per_step": ln[i] // a.profile_steps,
                                   "tflops": fl[i] / msv[i] / 1e9 if fl[i] else None,
                                   "gbs": by[i] / msv[i] / 1e6 if not fl[i] else None}
        # dominant kernel = gemm_tcgen05_kernel (linear + implicit-conv launches of the same kernel)
        g_ms, g_fl, g_n = msv[0] + msv[1], fl[0] + fl[1], ln[0] + ln[1]
        ach = g_fl / g_ms / 1e9
        traffic, traffic_note = None, None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload (never measured live)
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")))
            if (a.encoder, a.size, a.batch) == ("vitl", 518, 32):
                traffic = tj["mean_dram_bytes_per_launch"]
                traffic_note = (f"mean over the 4 linear GEMMs of one encoder block (96 of {g_n // a.profile_steps} GEMM launches/"
                                f"step), algorithmic {tj['mean_algorithmic_bytes_per_launch']} B/launch; {tj['source']}")
        except Exception:  # noqa: BLE001
            pass
        roof = {"kernel": "gemm_tcgen05_kernel", "bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["src"] + ", sustained bf16 (kernel timed inside a long step)",
                "flops_per_launch": g_fl / g_n, "avg_launch_ms": g_ms / g_n, "launches_per_step": g_n // a.profile_steps,
                "share_of_step": g_ms / tot}

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        rate, sec, cores = cpu_reference_rate(a.encoder, a.size, 3, 1)
        cpu_base = {"context": cpu_context(), "value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                    "sample": f"3 images of the same workload ({a.encoder} {a.size}x{a.size}, batch 1 per step, fp32 torch CPU "
                              f"oracle) after 1 warm-up, {sec:.2f} s/image"}

    if rank == 0:
        gb = cfg["global_batch"]
        value = gb / (ms / 1e3)
        gf = GFLOP_PER_IMAGE.get((a.encoder, a.size))
        line = {
            "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": gb / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
                    "how": "StreamedInference: pinned host -> H2D -> AmodalDAv2.forward -> D2H per step, copies on a side stream"},
            "gpu_launches": launches * a.steps,
            "clocks": clocks,
            "roofline": roof,
            "model_tflops_per_gpu": value * gf / 1e3 / world if gf else None,
            "model_frac_of_peak": (value * gf / 1e3 / world) / peaks["tf_sustained"] if gf else None,
            "model_frac_of_burst_peak": (value * gf / 1e3 / world) / peaks["tf_burst"] if gf else None,
            "breakdown": breakdown,
            "cpu_baseline": cpu_base,
            "workspace_gb": model.workspace_bytes() / 2 ** 30,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
