"""ORACLE SUPPORT -- TEST INFRASTRUCTURE. Deterministic synthetic state dicts and inputs for the parity tests.

`make_state_dict` enumerates the reference's state-dict template (verified by loading the result
into the unmodified reference model with strict=True in tests/golden/make_golden.py) and fills it from a seeded CPU
generator, so the same weights exist on the build box and on the GPU box without shipping a checkpoint.
Every parameter is given a non-degenerate value (the reference zero-initialises the guidance conv, dav2.py:55-61, which
would hide bugs in the guide path -- SURVEY.md fact 3).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .amodal_oracle import CONFIGS, GUIDE_CHANNELS, POS_GRID


def state_dict_shapes(encoder: str, guide_type: str, features=None, out_channels=None,
                      input_projection: bool = True) -> "OrderedDict[str, tuple]":
    """features / out_channels override the wrapper's per-encoder table (dav2.py:31-34) -- the un-guided model takes them
    as constructor arguments (depth_anything_v2_raw/dpt.py:155-162) and has no input_projection."""
    c = CONFIGS[encoder]
    D, F, C = c["embed_dim"], features or c["features"], list(out_channels or c["out_channels"])
    s: "OrderedDict[str, tuple]" = OrderedDict()
    p = "encoder.pretrained."
    s[p + "cls_token"] = (1, 1, D)
    s[p + "pos_embed"] = (1, 1 + POS_GRID * POS_GRID, D)
    s[p + "mask_token"] = (1, D)
    s[p + "patch_embed.proj.weight"] = (D, 3, 14, 14)
    s[p + "patch_embed.proj.bias"] = (D,)
    cg = GUIDE_CHANNELS[guide_type]
    if cg:
        s[p + "patch_embed_guidance.proj.weight"] = (D, cg, 14, 14)
        s[p + "patch_embed_guidance.proj.bias"] = (D,)
    for i in range(c["depth"]):
        b = p + f"blocks.{i}."
        s[b + "norm1.weight"] = (D,)
        s[b + "norm1.bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D)
        s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D)
        s[b + "attn.proj.bias"] = (D,)
        s[b + "ls1.gamma"] = (D,)
        s[b + "norm2.weight"] = (D,)
        s[b + "norm2.bias"] = (D,)
        if c["ffn"] == "mlp":
            s[b + "mlp.fc1.weight"] = (c["hidden"], D)
            s[b + "mlp.fc1.bias"] = (c["hidden"],)
            s[b + "mlp.fc2.weight"] = (D, c["hidden"])
            s[b + "mlp.fc2.bias"] = (D,)
        else:
            s[b + "mlp.w12.weight"] = (2 * c["hidden"], D)
            s[b + "mlp.w12.bias"] = (2 * c["hidden"],)
            s[b + "mlp.w3.weight"] = (D, c["hidden"])
            s[b + "mlp.w3.bias"] = (D,)
        s[b + "ls2.gamma"] = (D,)
    s[p + "norm.weight"] = (D,)
    s[p + "norm.bias"] = (D,)
    h = "encoder.depth_head."
    for i in range(4):
        s[h + f"projects.{i}.weight"] = (C[i], D, 1, 1)
        s[h + f"projects.{i}.bias"] = (C[i],)
    s[h + "resize_layers.0.weight"] = (C[0], C[0], 4, 4)
    s[h + "resize_layers.0.bias"] = (C[0],)
    s[h + "resize_layers.1.weight"] = (C[1], C[1], 2, 2)
    s[h + "resize_layers.1.bias"] = (C[1],)
    s[h + "resize_layers.3.weight"] = (C[3], C[3], 3, 3)
    s[h + "resize_layers.3.bias"] = (C[3],)
    for i in range(4):
        s[h + f"scratch.layer{i + 1}_rn.weight"] = (F, C[i], 3, 3)
    for k in range(1, 5):
        r = h + f"scratch.refinenet{k}."
        s[r + "out_conv.weight"] = (F, F, 1, 1)
        s[r + "out_conv.bias"] = (F,)
        for u in (1, 2):
            for cv in (1, 2):
                s[r + f"resConfUnit{u}.conv{cv}.weight"] = (F, F, 3, 3)
                s[r + f"resConfUnit{u}.conv{cv}.bias"] = (F,)
    s[h + "scratch.output_conv1.weight"] = (F // 2, F, 3, 3)
    s[h + "scratch.output_conv1.bias"] = (F // 2,)
    s[h + "scratch.output_conv2.0.weight"] = (32, F // 2, 3, 3)
    s[h + "scratch.output_conv2.0.bias"] = (32,)
    s[h + "scratch.output_conv2.2.weight"] = (1, 32, 1, 1)
    s[h + "scratch.output_conv2.2.bias"] = (1,)
    for i in range(4 if input_projection else 0):
        s[h + f"input_projection.{i}.0.weight"] = (C[i], C[i], 3, 3)
        s[h + f"input_projection.{i}.0.bias"] = (C[i],)
        s[h + f"input_projection.{i}.1.weight"] = (C[i],)
        s[h + f"input_projection.{i}.1.bias"] = (C[i],)
    return s


def make_state_dict_raw(encoder: str, features: int, out_channels, seed: int = 0, shift: float = 0.25):
    """Seeded state dict of the un-guided DepthAnythingV2 (depth_anything_v2_raw/dpt.py:154-175): keys `pretrained.*` /
    `depth_head.*` (no `encoder.` prefix, no guidance, no input_projection). The last bias is shifted so that most of the
    ReLU output is positive, as a trained depth model's is (an all-zero output would make the relative bar vacuous)."""
    sd = make_state_dict(encoder, "none", seed, False, features=features, out_channels=out_channels,
                         input_projection=False)
    sd = OrderedDict((k[len("encoder."):], v) for k, v in sd.items())
    sd["depth_head.scratch.output_conv2.2.bias"] += shift
    return sd


def make_state_dict(encoder: str, guide_type: str = "mask+observation", seed: int = 0, stress: bool = False,
                    features=None, out_channels=None, input_projection: bool = True):
    """Seeded fp32 CPU state dict. Distributions follow the reference's init in spirit (Linear ~N(0,.02),
    convs uniform(+-1/sqrt(fan_in)) = torch's default, dinov2.py:359-364) but biases, LayerNorm affines, LayerScale and
    the guidance conv are randomised so that every tensor influences the output. `stress=True` scales the last conv so
    the sigmoid output spans most of (0,1) (tolerance regime)."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = OrderedDict()
    for k, shp in state_dict_shapes(encoder, guide_type, features, out_channels, input_projection).items():
        if k.endswith("mask_token"):
            t = torch.zeros(shp)
        elif k.endswith("cls_token") or k.endswith("pos_embed"):
            t = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif ("norm" in k or "input_projection" in k and k.split(".")[-2] == "1") and k.endswith("weight") and len(shp) == 1:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            t = 0.02 * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            t = torch.randn(shp, generator=g) * 0.02
        else:  # conv / conv-transpose weights
            fan_in = shp[1] * shp[2] * shp[3]
            if "patch_embed" in k:
                t = torch.randn(shp, generator=g) * 0.02
            else:
                bound = (1.0 / fan_in) ** 0.5  # torch's default conv init (kaiming_uniform, a=sqrt(5))
                t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        sd[k] = t
    if stress:
        sd["encoder.depth_head.scratch.output_conv2.2.weight"] *= 40.0
    return sd


def make_inputs(B: int, H: int, W: int, seed: int = 0):
    """SURVEY.md section 8d synthetic inputs: rgb uniform [0,1]; mask = smooth blob in {-1,+1}; observation uniform
    [-1,1]; guide_rgb uniform [-1,1]. Also returns the boolean mask used for the AbsRel bar."""
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.rand(B, 3, H, W, generator=g)
    low = torch.rand(B, 1, max(H // 37, 2), max(W // 37, 2), generator=g)
    mask01 = (torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) > 0.5).float()
    observation = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    guide_rgb = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    return dict(x=x, guide_rgb=guide_rgb, guide_mask=mask01 * 2 - 1, observation=observation, mask01=mask01 > 0.5)
