"""Generates tests/golden/eval/*.npz by running the reference's own evaluation functions (build container only):
src/util/alignment.py:align_depth_least_square and the ten metric functions of src/util/metric.py, called exactly as
discriminative_trainer.py:542-613 calls them, on seeded samples from oracle.eval_oracle.synth_sample.
`skimage` (imported by metric.py for an unrelated edge metric) is stubbed. Re-run: python tests/golden/make_golden_eval.py"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [("kitti_like_518", 61, 518, 518, 375, 1242), ("small_126x98", 62, 126, 98, 200, 333), ("up_70", 63, 70, 70, 518, 518)]


def main():
    sk = types.ModuleType("skimage")
    skf = types.ModuleType("skimage.feature")
    skf.canny = None
    sys.modules.update({"skimage": sk, "skimage.feature": skf})
    for name, path in (("src", f"{REF}/src"), ("src.util", f"{REF}/src/util")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    from src.util import metric as M
    from src.util.alignment import align_depth_least_square
    from oracle.eval_oracle import METRICS, synth_sample
    for name, seed, h, w, H, W in CASES:
        s = synth_sample(seed, h, w, H, W)
        pred = F.interpolate(s["pred"], size=(H, W), mode="nearest").squeeze()
        depth_align, scale, shift = align_depth_least_square(gt_arr=s["depth_obs"].numpy(), pred_arr=pred.numpy(),
                                                             valid_mask_arr=s["visible_mask"].bool(), return_scale_shift=True,
                                                             max_resolution=None)
        depth_align = torch.tensor(depth_align)
        vals = {"scale": float(scale[0]), "shift": float(shift[0])}
        for met in METRICS:
            f = getattr(M, met)
            vals["pred_" + met] = float(f(pred + 1e-5, s["depth_gt"] + 1e-5, s["object_mask"]))
            vals["aligned_" + met] = float(f(depth_align + 1e-5, s["depth_gt"] + 1e-5, s["object_mask"]))
        np.savez(os.path.join(ROOT, "tests", "golden", "eval", name + ".npz"), seed=seed, h=h, w=w, H=H, W=W,
                 **{k: np.float64(v) for k, v in vals.items()})
        print(name, {k: round(v, 5) for k, v in vals.items() if "abs_rel" in k or k in ("scale", "shift")}, int(s["object_mask"].sum()))


if __name__ == "__main__":
    main()
