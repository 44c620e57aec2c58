"""Generates tests/golden/post/*.npz: outputs of the reference's own pre/post-processing code (build container only):
torchvision Resize(NEAREST) and F.interpolate as infer.py:84-87,97-101 call them, and infer.py:median_filter_blend itself
(imported from /root/reference/infer.py with `src.models`, `timm` and `matplotlib` stubbed -- none is used by it).
Re-run: python tests/golden/make_golden_post.py"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_infer():
    for name in ("timm", "matplotlib"):
        sys.modules.setdefault(name, types.ModuleType(name))
    for name, path in (("src", f"{REF}/src"), ("src.models", f"{REF}/src/models")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sys.modules["src.models"].get_model = lambda *a, **k: None
    sys.path.insert(0, REF)
    import infer  # noqa: E402
    return infer


def blob_mask(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 1, max(h // 40, 2), max(w // 40, 2), generator=g)
    return (torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) > 0.5)[0, 0].numpy()


def main():
    infer = import_infer()
    from torchvision.transforms import InterpolationMode, Resize
    out_dir = os.path.join(ROOT, "tests", "golden", "post")
    # ---- nearest resizes exactly as infer.py:84-87 and :100 perform them
    for name, (h0, w0), seed in (("nearest_300x401", (300, 401), 31), ("nearest_777x518", (777, 518), 32),
                                 ("nearest_1080x1920", (1080, 1920), 33)):
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, size=(h0, w0, 3), dtype=np.uint8)
        mask = blob_mask(h0, w0, seed)
        rgb_ts = torch.tensor(img).unsqueeze(dim=0).permute(0, 3, 1, 2) / 255
        tf = Resize(size=(518, 518), interpolation=InterpolationMode.NEAREST)
        rgb_r = tf(rgb_ts)
        m_r = (tf(torch.tensor(mask).float().unsqueeze(0).unsqueeze(0)) > 0).float()
        m_post = torch.nn.functional.interpolate(torch.tensor(mask).float().squeeze().unsqueeze(0).unsqueeze(0), (518, 518)).squeeze()
        rgb_u8 = (rgb_r[0] * 255).round().to(torch.uint8).numpy()   # exact: values are k/255
        assert np.array_equal(rgb_u8.astype(np.float32) / np.float32(255), rgb_r[0].numpy())
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), seed=seed, h0=h0, w0=w0,
                            rgb_sha256=hashlib.sha256(rgb_u8.tobytes()).hexdigest(),     # bit-exact comparison by digest
                            rgb_sample=rgb_u8[:, ::37, ::41],
                            mask_in=np.packbits(m_r[0, 0].to(torch.uint8).numpy()),
                            mask_post=np.packbits((m_post > 0).to(torch.uint8).numpy()))
        print(name, rgb_r.shape, float(m_r.mean()))
    # ---- infer.py:median_filter_blend on seeded fields
    for name, (h, w), seed in (("blend_140x154", (140, 154), 41),):
        g = torch.Generator().manual_seed(seed)
        raw = torch.rand(h, w, generator=g)
        amodal = torch.rand(h, w, generator=g)
        mask = blob_mask(h, w, seed).astype(np.float64)   # infer.py:103 passes amodal_mask/255 (float64 numpy)
        if name == "blend_140x154":
            mask[0, :7] = 1.0  # touch the image border: exercises zero padding of the mask and reflect-101 of the blur
            mask[-1, -5:] = 1.0
        out = infer.median_filter_blend(amodal.clone(), raw.clone(), mask)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), seed=seed, h=h, w=w, out=out.numpy().astype(np.float32))
        print(name, float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
