"""Pins oracle/infer_oracle.py (pre/post-processing of infer.py) against what the reference's
own code produced in the build container (tests/golden/post/, make_golden_post.py): torchvision Resize(NEAREST) /
F.interpolate for the resizes (bit exact) and infer.py:median_filter_blend (cv2.blur: 1e-6)."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import infer_oracle as IO

POST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post")


def synth_image_mask(h0, w0, seed):
    """Must match tests/golden/make_golden_post.py."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h0, w0, 3), dtype=np.uint8)
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 1, max(h0 // 40, 2), max(w0 // 40, 2), generator=g)
    mask = (torch.nn.functional.interpolate(low, size=(h0, w0), mode="bilinear", align_corners=False) > 0.5)[0, 0].numpy()
    return img, mask


def synth_blend(h, w, seed, touch_border):
    g = torch.Generator().manual_seed(seed)
    raw = torch.rand(h, w, generator=g)
    amodal = torch.rand(h, w, generator=g)
    _, mask = synth_image_mask(h, w, seed)
    mask = mask.astype(np.float64)
    if touch_border:
        mask[0, :7] = 1.0
        mask[-1, -5:] = 1.0
    return raw, amodal, mask


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(POST, "nearest_*.npz"))))
def test_nearest_resizes_bit_exact(path):
    z = np.load(path)
    img, mask = synth_image_mask(int(z["h0"]), int(z["w0"]), int(z["seed"]))
    rgb = IO.image_to_tensor_nearest(img)
    u8 = (rgb[0] * 255).round().to(torch.uint8).numpy()
    assert np.array_equal(u8.astype(np.float32) / np.float32(255), rgb[0].numpy())
    assert hashlib.sha256(u8.tobytes()).hexdigest() == str(z["rgb_sha256"])
    m = IO.mask_to_tensor_nearest(mask)[0, 0].to(torch.uint8).numpy()
    assert np.array_equal(np.packbits(m), z["mask_in"]) and np.array_equal(np.packbits(m), z["mask_post"])


def test_blend_matches_reference_function():
    z = np.load(os.path.join(POST, "blend_140x154.npz"))
    raw, amodal, mask = synth_blend(int(z["h"]), int(z["w"]), int(z["seed"]), True)
    out = IO.median_filter_blend(amodal, raw, mask).numpy()
    assert np.abs(out - z["out"]).max() < 1e-6
    seam = out != np.where(mask > 0, amodal.numpy(), raw.numpy())
    assert 100 < seam.sum() < out.size // 2  # the fixture does exercise the seam


def test_normalize_base_depth():
    d = torch.rand(1, 518, 518) * 7 + 3
    b = IO.normalize_base_depth(d)
    assert b.shape == (518, 518) and b.min() == 0 and b.max() == 1
