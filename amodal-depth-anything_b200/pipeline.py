"""Host<->device streaming around AmodalDAv2.forward: inputs arrive in pinned host memory, results are returned to pinned
host memory, and the copies of batch k+1 / k-1 overlap the kernels of batch k (an H2D stream, a D2H stream and the compute
stream, double-buffered device inputs; with a single copy stream the upload of batch k+1 queued behind the download of
batch k, which waits for forward k -- measured: 3.7 ms of un-hidden H2D per 32-image step). This is plumbing only (torch streams/events); every batch still goes through the public
model call. Mirrors what infer.py / the eval loop do per sample (H2D at infer.py:89-92, D2H at infer.py:94)."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch


class StreamedInference:
    def __init__(self, model, device=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.h2d_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        self._slots = [None, None]

    def _slot(self, i, host_batch):
        if self._slots[i] is None or any(d.shape != h.shape for d, h in zip(self._slots[i]["dev"], host_batch)):
            self._slots[i] = {
                "dev": [torch.empty(h.shape, dtype=torch.float32, device=self.device) for h in host_batch],
                "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
            }
            self._slots[i]["free"].record(torch.cuda.current_stream(self.device))
        return self._slots[i]

    @torch.no_grad()
    def run(self, host_batches: Iterable[Sequence[torch.Tensor]], host_outs: List[torch.Tensor]) -> None:
        """host_batches: iterable of (x, guide_mask, observation) pinned CPU tensors; host_outs[k]: pinned [B,1,H,W]."""
        compute = torch.cuda.current_stream(self.device)
        pending = []
        for k, hb in enumerate(host_batches):
            s = self._slot(k & 1, hb)
            with torch.cuda.stream(self.h2d_stream):
                self.h2d_stream.wait_event(s["free"])           # the forward that last read this slot is done
                for d, h in zip(s["dev"], hb):
                    d.copy_(h, non_blocking=True)
                s["ready"].record(self.h2d_stream)
            compute.wait_event(s["ready"])
            x, m, o = s["dev"]
            out = self.model(x, guide_rgb=None, guide_mask=m, observation=o)
            s["free"].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            out.record_stream(self.d2h_stream)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(done)
                host_outs[k].copy_(out, non_blocking=True)
            pending.append(out)
        compute.wait_stream(self.d2h_stream)                    # results are in host memory once `compute` drains
