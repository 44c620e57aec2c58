"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU fp32 restatement of the reference's discriminative forward pass (AmodalDAv2.forward) as plain functional torch
ops over a state dict, each function citing the reference file:line it follows (paths relative to
/root/reference/src/models/amodalsynthdrive/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product (amodal-depth-anything_b200/) never does.

Where the arithmetic lives: the reference is pure Python over third-party PyTorch (ATen / mkldnn on CPU), pinned by the
reference at pytorch=2.0.1 (environment.yaml:251). This restatement calls the same ATen ops (conv2d, conv_transpose2d,
linear, layer_norm, softmax, gelu(erf), silu, bilinear/bicubic interpolate) from the torch in this image (2.11), whose
semantics for these arguments are unchanged.

Pinning: the reference ships no tests / golden vectors for this path, so the oracle is pinned
against outputs of the reference itself, generated in the build container by tests/golden/make_golden.py (imports the
unmodified reference from /root/reference, loads the same seeded state dict with strict=True, records outputs and
intermediates) and committed under tests/golden/*.npz. tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

# dav2.py:31-34 (+ the ViT-G head sizes used by infer.py:60 / app.py:42, which dav2.py's table lacks),
# dpt.py:213-218 (taps), dinov2.py:366-427 (encoder sizes), swiglu_ffn.py:57 (hidden = (int(4D*2/3)+7)//8*8 = 4096)
CONFIGS = {
    "vits": dict(embed_dim=384, depth=12, num_heads=6, ffn="mlp", hidden=1536, taps=[2, 5, 8, 11], features=64,
                 out_channels=[48, 96, 192, 384]),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12, ffn="mlp", hidden=3072, taps=[2, 5, 8, 11], features=128,
                 out_channels=[96, 192, 384, 768]),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16, ffn="mlp", hidden=4096, taps=[4, 11, 17, 23], features=256,
                 out_channels=[256, 512, 1024, 1024]),
    "vitg": dict(embed_dim=1536, depth=40, num_heads=24, ffn="swiglu", hidden=4096, taps=[9, 19, 29, 39], features=384,
                 out_channels=[1536, 1536, 1536, 1536]),
}
# dinov2.py:110-125
GUIDE_CHANNELS = {"image+mask+observation": 5, "image+mask": 4, "image+observation": 4, "mask+observation": 2,
                  "mask": 1, "observation": 1, "none": 0}
PIXEL_MEAN = (0.485, 0.456, 0.406)  # dav2.py:50
PIXEL_STD = (0.229, 0.224, 0.225)   # dav2.py:51
PATCH = 14
POS_GRID = 37                       # dinov2.py:437 img_size=518 / 14
INTERPOLATE_OFFSET = 0.1            # dinov2.py:446


def build_guide(guide_type: str, guide_rgb, guide_mask, observation):
    """dav2.py:67-82."""
    if guide_type == "image+mask+observation":
        return torch.cat([guide_rgb, guide_mask, observation], dim=1)
    if guide_type == "image+mask":
        return torch.cat([guide_rgb, guide_mask], dim=1)
    if guide_type == "image+observation":
        return torch.cat([guide_rgb, observation], dim=1)
    if guide_type == "mask+observation":
        return torch.cat([guide_mask, observation], dim=1)
    if guide_type == "observation":
        return observation
    if guide_type == "mask":
        return guide_mask
    if guide_type == "none":
        return None
    raise NotImplementedError


def normalize_rgb(x):
    """dav2.py:65 with the buffers of dav2.py:50-51."""
    mean = torch.tensor(PIXEL_MEAN, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    std = torch.tensor(PIXEL_STD, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    return (x - mean) / std


def patch_embed(x, w, b):
    """dinov2_layers/patch_embed.py:69-82: conv k14 s14, flatten(2).transpose(1,2)."""
    _, _, H, W = x.shape
    assert H % PATCH == 0, f"Input image height {H} is not a multiple of patch height {PATCH}"
    assert W % PATCH == 0, f"Input image width {W} is not a multiple of patch width: {PATCH}"
    return F.conv2d(x, w, b, stride=PATCH).flatten(2).transpose(1, 2)


def interpolate_pos_encoding(pos_embed, npatch: int, w: int, h: int):
    """dinov2.py:199-230. `w`, `h` are x.shape[2], x.shape[3] as the reference names them (dinov2.py:233)."""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    pos_embed = pos_embed.float()
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    w0, h0 = w // PATCH + INTERPOLATE_OFFSET, h // PATCH + INTERPOLATE_OFFSET
    sqrt_n = math.sqrt(N)
    sx, sy = float(w0) / sqrt_n, float(h0) / sqrt_n
    patch_pos = F.interpolate(patch_pos.reshape(1, int(sqrt_n), int(sqrt_n), dim).permute(0, 3, 1, 2),
                              scale_factor=(sx, sy), mode="bicubic", antialias=False)
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def prepare_tokens(sd: Dict[str, torch.Tensor], x, guide):
    """dinov2.py:232-258 (masks=None, no register tokens)."""
    p = "encoder.pretrained."
    _, _, w, h = x.shape
    t = patch_embed(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"])
    if guide is not None:
        t = t + patch_embed(guide, sd[p + "patch_embed_guidance.proj.weight"], sd[p + "patch_embed_guidance.proj.bias"])
    t = torch.cat((sd[p + "cls_token"].expand(t.shape[0], -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd[p + "pos_embed"], t.shape[1] - 1, w, h)


def attention(x, wqkv, bqkv, wproj, bproj, heads: int):
    """dinov2_layers/attention.py:49-62 (the no-xformers branch of MemEffAttention, attention.py:66-69)."""
    B, N, C = x.shape
    qkv = F.linear(x, wqkv, bqkv).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    return F.linear((attn @ v).transpose(1, 2).reshape(B, N, C), wproj, bproj)


def ffn(sd, pre: str, x, kind: str):
    """dinov2_layers/mlp.py:35-41 (fc1, exact-erf GELU, fc2) or swiglu_ffn.py:29-33 (w12, silu(x1)*x2, w3)."""
    if kind == "mlp":
        return F.linear(F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"])), sd[pre + "fc2.weight"],
                        sd[pre + "fc2.bias"])
    x1, x2 = F.linear(x, sd[pre + "w12.weight"], sd[pre + "w12.bias"]).chunk(2, dim=-1)
    return F.linear(F.silu(x1) * x2, sd[pre + "w3.weight"], sd[pre + "w3.bias"])


def block(sd, i: int, x, cfg):
    """dinov2_layers/block.py:82-107, eval branch (105-106); LayerScale layer_scale.py:27-28; LN eps 1e-6 dinov2.py:96."""
    b = f"encoder.pretrained.blocks.{i}."
    D = cfg["embed_dim"]
    h = F.layer_norm(x, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)
    x = x + sd[b + "ls1.gamma"] * attention(h, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"],
                                            sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"], cfg["num_heads"])
    h = F.layer_norm(x, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
    return x + sd[b + "ls2.gamma"] * ffn(sd, b + "mlp.", h, cfg["ffn"])


def intermediate_layers(sd, cfg, x, guide, inter: Optional[dict] = None):
    """dinov2.py:298-308 + 324-349: outputs after the tapped blocks, shared final norm, cls split off."""
    t = prepare_tokens(sd, x, guide)
    if inter is not None:
        inter["tokens"] = t
    outs = []
    for i in range(cfg["depth"]):
        t = block(sd, i, t, cfg)
        if i in cfg["taps"]:
            outs.append(t)
    p = "encoder.pretrained."
    outs = [F.layer_norm(o, (cfg["embed_dim"],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6) for o in outs]
    return [(o[:, 1:], o[:, 0]) for o in outs]


def channel_layernorm(x, w, b, eps=1e-6):
    """dpt.py:56-61 (channels_first)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def residual_conv_unit(sd, pre: str, x):
    """util/blocks.py:57-80 (bn=False, groups=1; activation = nn.ReLU(False), dpt.py:15)."""
    out = F.relu(x)
    out = F.conv2d(out, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return out + x


def feature_fusion(sd, pre: str, xs: List[torch.Tensor], size=None):
    """util/blocks.py:123-148 (align_corners=True, dpt.py:18)."""
    out = xs[0]
    if len(xs) == 2:
        out = out + residual_conv_unit(sd, pre + "resConfUnit1.", xs[1])
    out = residual_conv_unit(sd, pre + "resConfUnit2.", out)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def dpt_head(sd, cfg, feats, patch_h: int, patch_w: int, sigmoid: bool, inter: Optional[dict] = None,
             input_projection: bool = True):
    """dpt.py:161-197 (use_clstoken=False). input_projection=False: the un-guided head of
    depth_anything_v2_raw/dpt.py:118-151 (identical but for the missing input_projection levels)."""
    h = "encoder.depth_head."
    out = []
    for i, (x, _cls) in enumerate(feats):
        x = x.permute(0, 2, 1).reshape(x.shape[0], x.shape[-1], patch_h, patch_w)       # dpt.py:171
        x = F.conv2d(x, sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"])    # dpt.py:172
        if i == 0:                                                                       # dpt.py:89-107,173
            x = F.conv_transpose2d(x, sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        out.append(x)
    layers = []
    for i, x in enumerate(out):                                                          # dpt.py:153-159,178-179
        if input_projection:
            x = F.conv2d(x, sd[h + f"input_projection.{i}.0.weight"], sd[h + f"input_projection.{i}.0.bias"], padding=1)
            x = F.relu(channel_layernorm(x, sd[h + f"input_projection.{i}.1.weight"],
                                         sd[h + f"input_projection.{i}.1.bias"]))
        layers.append(x)
    rn = [F.conv2d(layers[i], sd[h + f"scratch.layer{i + 1}_rn.weight"], None, padding=1) for i in range(4)]  # 184-187
    s = h + "scratch."
    path_4 = feature_fusion(sd, s + "refinenet4.", [rn[3]], size=rn[2].shape[2:])        # dpt.py:189-192
    path_3 = feature_fusion(sd, s + "refinenet3.", [path_4, rn[2]], size=rn[1].shape[2:])
    path_2 = feature_fusion(sd, s + "refinenet2.", [path_3, rn[1]], size=rn[0].shape[2:])
    path_1 = feature_fusion(sd, s + "refinenet1.", [path_2, rn[0]])
    o = F.conv2d(path_1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)           # dpt.py:193
    o = F.interpolate(o, (int(patch_h * 14), int(patch_w * 14)), mode="bilinear", align_corners=True)     # dpt.py:194
    o = F.relu(F.conv2d(o, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))    # dpt.py:146-151
    logits = F.conv2d(o, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])
    if inter is not None:
        for i in range(4):
            inter[f"layer{i + 1}"] = layers[i]
            inter[f"layer{i + 1}_rn"] = rn[i]
        inter.update(path_4=path_4, path_3=path_3, path_2=path_2, path_1=path_1, logits=logits)
    return torch.sigmoid(logits) if sigmoid else logits


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], encoder: str, guide_type: str, x, guide_rgb=None, guide_mask=None,
            observation=None, loss_stategy: str = "invisible_part", inter: Optional[dict] = None):
    """AmodalDAv2.forward (dav2.py:64-85) -> DepthAnythingV2.forward (dpt.py:225-231). Returns [B,1,H,W] fp32."""
    cfg = CONFIGS[encoder]
    x = normalize_rgb(x)
    guide = build_guide(guide_type, guide_rgb, guide_mask, observation)
    patch_h, patch_w = x.shape[-2] // 14, x.shape[-1] // 14
    feats = intermediate_layers(sd, cfg, x, guide, inter)
    if inter is not None:
        for i, (f, _c) in enumerate(feats):
            inter[f"tap{i}"] = f
    return dpt_head(sd, cfg, feats, patch_h, patch_w, sigmoid=("ssi" not in loss_stategy), inter=inter)


@torch.no_grad()
def forward_raw(sd: Dict[str, torch.Tensor], encoder: str, x, inter: Optional[dict] = None):
    """Un-guided DepthAnythingV2.forward (depth_anything_v2_raw/dpt.py:176-184): x is the ImageNet-normalised image
    (the caller normalises, infer.py:18); the head's Sequential ends in ReLU (:109-116) and forward applies ReLU again
    (:182). sd uses the raw model's keys (`pretrained.*`, `depth_head.*`). Returns [B,H,W] fp32."""
    sd = {"encoder." + k: v for k, v in sd.items()}
    cfg = CONFIGS[encoder]
    patch_h, patch_w = x.shape[-2] // 14, x.shape[-1] // 14
    feats = intermediate_layers(sd, cfg, x, None, inter)
    logits = dpt_head(sd, cfg, feats, patch_h, patch_w, sigmoid=False, inter=inter, input_projection=False)
    return F.relu(F.relu(logits)).squeeze(1)


def abs_relative_difference(output, target, valid_mask=None):
    """src/util/metric.py:37-47 (AbsRel used in the parity bar)."""
    actual = output
    abs_rel = torch.abs(actual - target) / target
    if valid_mask is not None:
        abs_rel[~valid_mask] = 0
        n = valid_mask.sum((-1, -2))
    else:
        n = output.shape[-1] * output.shape[-2]
    abs_rel = torch.sum(abs_rel, (-1, -2)) / n
    return abs_rel.mean()
