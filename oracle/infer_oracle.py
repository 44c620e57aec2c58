"""ORACLE -- TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

CPU restatement of the numeric path of the reference's single-image inference, /root/reference/infer.py:16-28,30-44,
72-103: what happens to the image, the mask and the two depth maps around the two networks.
Pinned against the reference's own functions (torchvision Resize(NEAREST), F.interpolate, infer.py:median_filter_blend run
in the build container) by tests/golden/post/*.npz, see tests/golden/make_golden_post.py. cv2.blur is restated in numpy
(BORDER_REFLECT_101, fp32) so the oracle does not need OpenCV on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .amodal_oracle import normalize_rgb


def image_to_tensor_nearest(img_u8_hwc: np.ndarray, size: int = 518) -> torch.Tensor:
    """infer.py:84-86: rgb_ts = tensor(img).unsqueeze(0).permute(0,3,1,2) / 255; Resize((518,518), NEAREST). torchvision's
    tensor path is F.interpolate(mode='nearest') (checked against torchvision itself by the golden)."""
    t = torch.tensor(img_u8_hwc).unsqueeze(0).permute(0, 3, 1, 2) / 255
    return F.interpolate(t, size=(size, size), mode="nearest")


def mask_to_tensor_nearest(mask: np.ndarray, size: int = 518) -> torch.Tensor:
    """infer.py:80-87 (network input) and :100-101 (blend mask): (mask > 0).float() -> nearest resize -> > 0. [1,1,S,S] in {0,1}."""
    m = torch.tensor(np.asarray(mask) > 0).float().unsqueeze(0).unsqueeze(0)
    return (F.interpolate(m, size=(size, size), mode="nearest") > 0).float()


def normalize_base_depth(depth_raw: torch.Tensor) -> torch.Tensor:
    """infer.py:19-23: depth_raw [1,H,W] -> nearest to 518 (identity at 518) -> (d - min) / (max - min) -> [H,W]."""
    d = depth_raw.unsqueeze(1)
    d = F.interpolate(d, (518, 518), mode="nearest") if tuple(d.shape[-2:]) != (518, 518) else d
    d = (d - d.min()) / (d.max() - d.min())
    return d.squeeze()


def box_blur3_reflect101(a: np.ndarray) -> np.ndarray:
    """cv2.blur(a, (3,3)) for fp32: BORDER_REFLECT_101 padding, mean of the 3x3 window."""
    p = np.pad(a.astype(np.float32), 1, mode="reflect")
    rows = (p[:, :-2] + p[:, 1:-1]) + p[:, 2:]
    return (((rows[:-2] + rows[1:-1]) + rows[2:]) * np.float32(1.0 / 9.0)).astype(np.float32)


def median_filter_blend(depth_amodal_post: torch.Tensor, depth_agg: torch.Tensor, mask: np.ndarray, filter_width: int = 3):
    """infer.py:30-44 (despite its name a 3x3 box blur of the seam)."""
    assert filter_width == 3
    mask_t = torch.tensor(mask)
    blended = depth_agg.clone()
    blended[mask_t > 0] = depth_amodal_post[mask_t > 0]
    kernel = torch.ones((1, 1, 3, 3))
    dilated = F.conv2d(mask_t.float().unsqueeze(0).unsqueeze(0), kernel, padding=1)
    border = ((dilated > 0) & (dilated < 9)).squeeze().numpy()
    out = blended.numpy().copy()
    out[border] = box_blur3_reflect101(blended.numpy())[border]
    return torch.tensor(out)


@torch.no_grad()
def infer_single_image(img_u8_hwc: np.ndarray, img518_u8_hwc: np.ndarray, mask: np.ndarray, raw_fn, amodal_fn):
    """infer.py:72-103 without file IO / colour maps. img518 = cv2.resize(img, (518,518)) (host, infer.py:17); raw_fn /
    amodal_fn are the two networks (normalised image -> [1,518,518]; (rgb01, guide_mask, observation) -> [1,1,518,518])."""
    x = torch.tensor(img518_u8_hwc).permute(2, 0, 1).unsqueeze(0) / 255        # infer.py:18
    base = normalize_base_depth(raw_fn(normalize_rgb(x)))                       # infer.py:19-23
    rgb = image_to_tensor_nearest(img_u8_hwc)                                   # infer.py:84-86
    m01 = mask_to_tensor_nearest(mask)                                          # infer.py:87
    pred = amodal_fn(rgb.float(), m01 * 2 - 1, base.unsqueeze(0).unsqueeze(0) * 2 - 1)   # infer.py:88-93
    agg = median_filter_blend(pred.squeeze(), base.clone(), m01.squeeze().numpy())        # infer.py:97-103
    return dict(base_depth=base, pred=pred, depth_agg=agg)
