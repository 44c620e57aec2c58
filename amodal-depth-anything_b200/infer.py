"""Device-resident version of the numeric path of the reference's single-image inference (infer.py:16-28,72-103; SURVEY.md
section 8 row f2): image + amodal mask -> observation depth (un-guided model) -> amodal depth (guided model) -> blended
depth with a smoothed seam. In the reference every step between the two networks goes through the host
(`.detach().cpu()` at infer.py:19,94, numpy blend + cv2.blur at :30-44); here the image and the mask are uploaded once
and everything else runs as kernels on the current stream. Colour maps / PNG writing (infer.py:24-27,105-121) stay with
the caller: they are presentation, not part of the depth result.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class AmodalInference:
    def __init__(self, model_raw, depth_amodal_model, size: int = 518, cuda_graph: bool = True):
        self.model_raw = model_raw.set_graph(cuda_graph)      # one image per call is launch bound: replay as CUDA graphs
        self.model = depth_amodal_model.set_graph(cuda_graph)
        self.size = size
        self.device = next(depth_amodal_model.parameters()).device

    def _upload_u8(self, a: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device, non_blocking=True)

    @torch.no_grad()
    def predict_base_depth(self, image_u8_hwc: np.ndarray):
        """infer.py:16-23 -> (base01 [S,S], observation = base01*2-1 [1,1,S,S]) on the device."""
        S = self.size
        img = image_u8_hwc
        if img.shape[:2] != (S, S):  # infer.py:17 resizes the uint8 image on the host with cv2 (bilinear, fixed point)
            import cv2
            img = cv2.resize(img, (S, S))
        x = ops.image_nearest(self._upload_u8(img), S, S, normalize=True)       # infer.py:18 (identity sampling)
        depth_raw = self.model_raw(x)                                            # [1,S,S], infer.py:19
        base, obs = ops.minmax_normalize(depth_raw)                              # infer.py:20-22
        return base[0], obs.unsqueeze(1)

    @torch.no_grad()
    def __call__(self, image_u8_hwc: np.ndarray, amodal_mask: np.ndarray):
        """infer.py:72-103. image: uint8 [H0,W0,3] as cv2.imread returns it; amodal_mask: [H0,W0], non-zero = object.
        Returns device tensors: base_depth [S,S] in [0,1], pred [1,1,S,S], depth_agg [S,S]."""
        S = self.size
        base, obs = self.predict_base_depth(image_u8_hwc)
        rgb = ops.image_nearest(self._upload_u8(image_u8_hwc), S, S)                                # infer.py:84-86
        m01, guide = ops.mask_nearest(self._upload_u8((np.asarray(amodal_mask) > 0).astype(np.uint8)), S, S)   # :80-87
        pred = self.model(rgb, guide_rgb=None, guide_mask=guide, observation=obs)                   # infer.py:88-93
        agg = ops.blend_seam(base, pred[0, 0], m01[0, 0])                                           # infer.py:97-103
        return dict(base_depth=base, pred=pred, depth_agg=agg)
