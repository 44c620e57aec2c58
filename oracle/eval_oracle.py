"""ORACLE -- TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

CPU restatement of the per-sample evaluation post-ops of the reference's validation loop
(/root/reference/src/trainer/discriminative_trainer.py:542-613): nearest resize of the prediction, least-squares
scale/shift alignment (src/util/alignment.py:7-54) and the ten masked metrics of src/util/metric.py:37-161.
Pinned against the reference's own functions by tests/golden/eval/*.npz (tests/golden/make_golden_eval.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

METRICS = ("abs_relative_difference", "squared_relative_difference", "rmse_linear", "rmse_log", "log10", "delta1_acc",
           "delta2_acc", "delta3_acc", "i_rmse", "silog_rmse")  # config/*.yaml eval.eval_metrics


def align_depth_least_square(gt_arr, pred_arr, valid_mask_arr):
    """alignment.py:7-54 with max_resolution=None: lstsq of [pred, 1] against gt over the valid pixels."""
    gt, pred, valid = gt_arr.squeeze(), pred_arr.squeeze(), valid_mask_arr.squeeze()
    g = gt[valid].reshape((-1, 1))
    p = pred[valid].reshape((-1, 1))
    A = np.concatenate([p, np.ones_like(p)], axis=-1)
    scale, shift = np.linalg.lstsq(A, g, rcond=None)[0]
    return pred_arr * scale + shift, scale, shift


def _masked_mean(x, valid):
    x = x.clone()
    x[~valid] = 0
    return torch.sum(x, (-1, -2)) / valid.sum((-1, -2))


def metric(name, output, target, valid):
    """metric.py:37-161 for one [H, W] sample."""
    if name == "abs_relative_difference":
        return _masked_mean(torch.abs(output - target) / target, valid).mean()
    if name == "squared_relative_difference":
        return _masked_mean(torch.pow(torch.abs(output - target), 2) / target, valid).mean()
    if name == "rmse_linear":
        d = (output - target).clone()
        d[~valid] = 0
        return torch.sqrt(torch.sum(torch.pow(d, 2), (-1, -2)) / valid.sum((-1, -2))).mean()
    if name == "rmse_log":
        d = torch.log(output) - torch.log(target)
        d[~valid] = 0
        return torch.sqrt(torch.sum(torch.pow(d, 2), (-1, -2)) / valid.sum((-1, -2))).mean()
    if name == "log10":
        return torch.abs(torch.log10(output[valid]) - torch.log10(target[valid])).mean()
    if name in ("delta1_acc", "delta2_acc", "delta3_acc"):
        thr = 1.25 ** int(name[5])
        r = torch.max(output / target, target / output)
        bit = torch.where(r < thr, torch.ones_like(r), torch.zeros_like(r))
        bit[~valid] = 0
        return (torch.sum(bit, (-1, -2)) / valid.sum((-1, -2))).mean()
    if name == "i_rmse":
        d = 1.0 / output - 1.0 / target
        d[~valid] = 0
        return torch.sqrt(torch.sum(torch.pow(d, 2), (-1, -2)) / valid.sum((-1, -2))).mean()
    if name == "silog_rmse":
        d = torch.log(output) - torch.log(target)
        d[~valid] = 0
        n = valid.sum((-1, -2))
        first = torch.sum(torch.pow(d, 2), (-1, -2)) / n
        second = torch.pow(torch.sum(d, (-1, -2)), 2) / (n ** 2)
        return torch.sqrt(torch.mean(first - second)) * 100
    raise KeyError(name)


def evaluate_sample(pred, depth_gt, depth_obs, visible_mask, object_mask):
    """discriminative_trainer.py:542-613 for one sample. pred: [1,1,h,w] network output; the rest [H,W]."""
    pred = F.interpolate(pred, size=tuple(depth_gt.shape[-2:]), mode="nearest").squeeze()                 # :542
    aligned, scale, shift = align_depth_least_square(depth_obs.numpy(), pred.numpy(), visible_mask.bool().numpy())   # :546-551
    aligned = torch.tensor(aligned)
    res = {"scale": float(scale[0]), "shift": float(shift[0]), "pred": {}, "aligned": {}}
    obj = object_mask.bool()
    for name in METRICS:                                                                                  # :584-613
        res["pred"][name] = float(metric(name, pred + 1e-5, depth_gt + 1e-5, obj))
        res["aligned"][name] = float(metric(name, aligned + 1e-5, depth_gt + 1e-5, obj))
    return res


def synth_sample(seed, h=518, w=518, H=375, W=1242):
    """Seeded sample shaped like the validation data: depth in (0,1], observation = noisy affine of the true depth,
    object (invisible) and visible masks as blobs."""
    g = torch.Generator().manual_seed(seed)

    def smooth(hh, ww, lo, hi):
        low = torch.rand(1, 1, max(hh // 30, 2), max(ww // 30, 2), generator=g)
        return F.interpolate(low, size=(hh, ww), mode="bilinear", align_corners=False)[0, 0] * (hi - lo) + lo
    gt = smooth(H, W, 0.05, 1.0)
    obs = (gt + 0.02 * torch.randn(H, W, generator=g)).clamp_min(1e-3)
    pred_full = (0.7 * gt + 0.1 + 0.03 * torch.randn(H, W, generator=g)).clamp_min(1e-3)
    pred = F.interpolate(pred_full[None, None], size=(h, w), mode="bilinear", align_corners=False)
    visible = smooth(H, W, 0, 1) > 0.55
    obj = (smooth(H, W, 0, 1) > 0.6) & ~visible
    return dict(pred=pred, depth_gt=gt, depth_obs=obs, visible_mask=visible, object_mask=obj)
