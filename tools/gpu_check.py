"""Bring-up harness: runs each kernel check in its own subprocess (a trapped kernel poisons its CUDA context) and writes
gpurun_out/gpu_check.json. `python tools/gpu_check.py` runs all; `--one NAME` runs a single check in-process."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _imports():
    import torch
    import amodal_depth_anything_b200  # noqa: F401
    from amodal_depth_anything_b200 import _lib as L, ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch, L, ops


def _cmp(name, got, ref, atol, rtol):
    import torch
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    res = {"max_abs": err.max().item(), "max_ref": ref.abs().max().item(), "bad": bad, "n": ref.numel(),
           "nan": int(torch.isnan(got).sum().item())}
    res["ok"] = bad == 0 and res["nan"] == 0
    return res


def chk_gemm(M, N, K, bn, epi_name, cg=0):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(N, K, generator=g, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    ref = A.float() @ Wt.float().t() + bias
    if epi_name == "bias":
        out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
        ops.gemm(A, Wt, bias=bias, out_bf16=out, ldo=N, force_bn=bn, force_cg=cg)
    elif epi_name == "gelu":
        out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
        ops.gemm(A, Wt, bias=bias, out_bf16=out, ldo=N, act=L.ACT_GELU, force_bn=bn, force_cg=cg)
        ref = torch.nn.functional.gelu(ref)
    elif epi_name == "resid":  # LayerScale'd branch output (bf16); the residual add itself is fused into the next LN
        gamma = torch.randn(N, generator=g, device="cuda")
        ref = gamma * ref
        out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
        ops.gemm(A, Wt, bias=bias, gamma=gamma, out_bf16=out, ldo=N, force_bn=bn, force_cg=cg)
    elif epi_name == "swiglu":
        Hd = N // 2
        # interleave rows in 32-chunks exactly like ada_finalize does
        idx = torch.arange(N, device="cuda")
        chunk, within = idx // 64, idx % 64
        src = torch.where(within < 32, chunk * 32 + within, Hd + chunk * 32 + within - 32)
        out = torch.zeros(M, Hd, dtype=torch.bfloat16, device="cuda")
        ops.gemm(A, Wt[src].contiguous(), epi=L.EPI_SWIGLU, bias=bias[src].contiguous(), out_bf16=out, ldo=Hd, force_bn=bn, force_cg=cg)
        x1, x2 = ref[:, :Hd], ref[:, Hd:]
        ref = torch.nn.functional.silu(x1) * x2
    torch.cuda.synchronize()
    return _cmp("gemm", out, ref, 2e-2, 1e-2)


def chk_embed():
    torch, L, ops = _imports()
    B, P, D, K = 2, 1369, 384, 1024
    g = torch.Generator(device="cuda").manual_seed(2)
    A = (torch.randn(B * P, K, generator=g, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(D, K, generator=g, device="cuda") * 0.05).bfloat16()
    posb = torch.randn(P, D, generator=g, device="cuda")
    x = torch.full((B * (P + 1), D), 7.0, device="cuda")
    ops.gemm(A, Wt, epi=L.EPI_EMBED, aux=posb, out_f32=x, ldo=D, P=P)
    torch.cuda.synchronize()
    ref = (A.float() @ Wt.float().t()).view(B, P, D) + posb
    got = x.view(B, P + 1, D)
    r = _cmp("embed", got[:, 1:], ref, 1e-2, 1e-3)
    r["cls_untouched"] = bool((got[:, 0] == 7.0).all().item())
    r["ok"] = r["ok"] and r["cls_untouched"]
    return r


def chk_conv(B, H, W, Cin, Cout, mode, cg=0):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(B, Cin, H, W, generator=g, device="cuda")).bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, generator=g, device="cuda") * (1.0 / (3 * Cin ** 0.5))
    bias = torch.randn(Cout, generator=g, device="cuda") * 0.1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    wp = ops.pack_conv3x3(w)
    wq = w.bfloat16().float()
    ref = torch.nn.functional.conv2d(x.float(), wq, bias, padding=1)
    if mode == "plain":
        out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
        ops.gemm(x_nhwc, wp, conv=(B, H, W, Cin), N=Cout, bias=bias, out_bf16=out, ldo=Cout, force_cg=cg)
        torch.cuda.synchronize()
        return _cmp("conv", out.permute(0, 3, 1, 2), ref, 3e-2, 1e-2)
    if mode == "rcu":  # relu act + two residuals + relu copy
        r1 = torch.randn(B, H, W, Cout, generator=g, device="cuda").bfloat16()
        r2 = torch.randn(B, H, W, Cout, generator=g, device="cuda").bfloat16()
        out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
        outr = torch.zeros_like(out)
        ops.gemm(x_nhwc, wp, conv=(B, H, W, Cin), N=Cout, bias=bias, act=L.ACT_RELU, resid1=r1, resid2=r2, out_bf16=out,
                 out_relu=outr, ldo=Cout, force_cg=cg)
        torch.cuda.synchronize()
        ref2 = torch.relu(ref).permute(0, 2, 3, 1) + r1.float() + r2.float()
        a = _cmp("conv", out, ref2, 3e-2, 1e-2)
        b = _cmp("conv", outr, torch.relu(ref2), 3e-2, 1e-2)
        a["relu_copy_ok"] = b["ok"]
        a["ok"] = a["ok"] and b["ok"]
        return a
    if mode == "chln":  # conv + channel LayerNorm + ReLU in the epilogue (dpt.py:153-159)
        lw = 1.0 + 0.2 * torch.randn(Cout, generator=g, device="cuda")
        lb = 0.3 * torch.randn(Cout, generator=g, device="cuda")
        out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
        ops.gemm(x_nhwc, wp, conv=(B, H, W, Cin), N=Cout, epi=L.EPI_BF16_CHLN, bias=bias, gamma=lw, aux=lb, out_bf16=out,
                 ldo=Cout, force_cg=cg)
        torch.cuda.synchronize()
        u = ref.mean(1, keepdim=True)
        sv = (ref - u).pow(2).mean(1, keepdim=True)
        ref2 = torch.relu((ref - u) / torch.sqrt(sv + 1e-6) * lw.view(1, -1, 1, 1) + lb.view(1, -1, 1, 1))
        return _cmp("conv_chln", out.permute(0, 3, 1, 2), ref2, 3e-2, 1e-2)
    if mode == "tail":
        assert Cout == 32
        w2 = torch.randn(32, generator=g, device="cuda") * 0.3
        b2 = torch.randn(1, generator=g, device="cuda") * 0.1
        aux = torch.cat([w2, b2]).contiguous()
        out = torch.zeros(B, H, W, dtype=torch.float32, device="cuda")
        ops.gemm(x_nhwc, wp, conv=(B, H, W, Cin), N=32, epi=L.EPI_TAIL, bias=bias, aux=aux, out_f32=out, sigmoid=1)
        torch.cuda.synchronize()
        z = (torch.relu(ref) * w2.view(1, 32, 1, 1)).sum(1) + b2
        return _cmp("tail", out, torch.sigmoid(z), 5e-3, 0)


def chk_convT(ks, Cin=96, Cout=96, B=2, H=37, W=37, pix=False):
    """pix: A operand on 8x16 pixel tiles, pixel shuffle stored through 5-D TMA boxes (Cin, Cout % 64 == 0)."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(B, Cin, H, W, generator=g, device="cuda").bfloat16()
    w = torch.randn(Cin, Cout, ks, ks, generator=g, device="cuda") * 0.1
    bias = torch.randn(Cout, generator=g, device="cuda") * 0.1
    ref = torch.nn.functional.conv_transpose2d(x.float(), w.bfloat16().float(), bias, stride=ks)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    wp = ops.pack_convT(w, ks)
    out = torch.zeros(B, H * ks, W * ks, Cout, dtype=torch.bfloat16, device="cuda")
    if pix:
        ops.gemm(x_nhwc, wp, conv=(B, H, W, Cin), conv_taps=1, N=ks * ks * Cout, epi=L.EPI_CONVT, bias=bias, out_bf16=out, ks=ks,
                 cout=Cout)
    else:
        ops.gemm(x_nhwc.reshape(B * H * W, Cin), wp, epi=L.EPI_CONVT, bias=bias, out_bf16=out, ks=ks, cout=Cout, H=H, W=W)
    torch.cuda.synchronize()
    return _cmp("convT", out.permute(0, 3, 1, 2), ref, 3e-2, 1e-2)


def chk_attention(B, N, heads, grow=False, impl=-1):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(5)
    D = heads * 64
    qkv = torch.randn(B, N, 3, heads, 64, generator=g, device="cuda") * 1.5
    if grow:  # key norms grow along the sequence -> row maxima jump by > 2^8 between KV tiles (exercises the O rescale)
        ramp = 1.0 + 7.0 * (torch.arange(N, device="cuda") // 128).float() / max((N - 1) // 128, 1)
        qkv[:, :, 1] *= ramp.view(1, N, 1, 1)
    qkv = qkv.bfloat16()
    out = ops.attention(qkv, B, N, heads, impl)
    torch.cuda.synchronize()
    q, k, v = [t.float().permute(0, 2, 1, 3) for t in qkv.unbind(2)]  # [B,h,N,64]
    att = torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    return _cmp("attention", out, ref, 2e-2, 2e-2)


def chk_attention_impls_agree(B, N, heads, grow=False):
    """The two attention kernels are selected by grid size; an image must not change with the batch it is processed in, so
    they have to agree bit for bit (same exponential evaluation per column, same summation order of the row sums)."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(15)
    qkv = torch.randn(B, N, 3, heads, 64, generator=g, device="cuda") * 1.5
    if grow:
        ramp = 1.0 + 7.0 * (torch.arange(N, device="cuda") // 128).float() / max((N - 1) // 128, 1)
        qkv[:, :, 1] *= ramp.view(1, N, 1, 1)
    qkv = qkv.bfloat16()
    a = ops.attention(qkv, B, N, heads, 0)
    worst, n_diff, ok = 0.0, 0, True
    for impl in (1, 2):
        b = ops.attention(qkv, B, N, heads, impl)
        torch.cuda.synchronize()
        diff = (a.float() - b.float()).abs()
        ok = ok and bool(torch.equal(a, b))
        worst = max(worst, diff.max().item())
        n_diff += int((diff > 0).sum().item())
    return {"ok": ok, "max_abs": worst, "n_diff": n_diff, "n": a.numel()}


def chk_layernorm(D, drop):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(6)
    B, n_tok = 3, 50
    x = torch.randn(B * n_tok, D, generator=g, device="cuda") * 2 + 0.5
    w = torch.randn(D, generator=g, device="cuda")
    b = torch.randn(D, generator=g, device="cuda")
    out = ops.layernorm(x, w, b, 1e-6, n_tok, drop)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)
    if drop:
        ref = ref.view(B, n_tok, D)[:, 1:].reshape(-1, D)
    return _cmp("ln", out, ref, 2e-2, 1e-2)


def chk_layernorm_delta():
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(16)
    rows, D = 300, 1024
    x = torch.randn(rows, D, generator=g, device="cuda") * 2
    d = torch.randn(rows, D, generator=g, device="cuda").bfloat16()
    w = torch.randn(D, generator=g, device="cuda")
    b = torch.randn(D, generator=g, device="cuda")
    x1 = x.clone()
    out = ops.layernorm(x1, w, b, 1e-6, 1, False, delta=d, write_x=True)
    x2 = x.clone()
    out2 = ops.layernorm(x2, w, b, 1e-6, 1, False, delta=d, write_x=False)
    torch.cuda.synchronize()
    xs = x + d.float()
    ref = torch.nn.functional.layer_norm(xs, (D,), w, b, 1e-6)
    r = _cmp("ln", out, ref, 2e-2, 1e-2)
    r["x_written"] = bool(torch.equal(x1, xs))
    r["x_untouched"] = bool(torch.equal(x2, x))
    r["ok"] = r["ok"] and r["x_written"] and r["x_untouched"] and bool(torch.equal(out, out2))
    # two pending branches: (x + d) + d2, the order the encoder relies on (norm2 sees x + d, the next norm1 adds d2)
    d2 = torch.randn(rows, D, generator=g, device="cuda").bfloat16()
    x3 = x.clone()
    out3 = ops.layernorm(x3, w, b, 1e-6, 1, False, delta=d, write_x=True, delta2=d2)
    torch.cuda.synchronize()
    xs2 = (x + d.float()) + d2.float()
    r2 = _cmp("ln2", out3, torch.nn.functional.layer_norm(xs2, (D,), w, b, 1e-6), 2e-2, 1e-2)
    r["x_written_2"] = bool(torch.equal(x3, xs2))
    r["ok"] = r["ok"] and r2["ok"] and r["x_written_2"]
    return r


def chk_layernorm_tap():
    """norm1 of the next block and the tap of the previous one in one pass: same statistics, two affines, the second written
    without the cls rows (dinov2.py:337-340)."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(26)
    B, n_tok, D = 3, 50, 1024
    x = torch.randn(B * n_tok, D, generator=g, device="cuda") * 2 + 0.3
    d1 = torch.randn(B * n_tok, D, generator=g, device="cuda").bfloat16()
    d2 = torch.randn(B * n_tok, D, generator=g, device="cuda").bfloat16()
    w, b, w2, b2 = (torch.randn(D, generator=g, device="cuda") for _ in range(4))
    x1 = x.clone()
    out, tap = ops.layernorm(x1, w, b, 1e-6, n_tok, False, delta=d1, write_x=True, delta2=d2, tap_w=w2, tap_b=b2)
    torch.cuda.synchronize()
    xs = (x + d1.float()) + d2.float()
    r = _cmp("ln", out, torch.nn.functional.layer_norm(xs, (D,), w, b, 1e-6), 2e-2, 1e-2)
    ref2 = torch.nn.functional.layer_norm(xs, (D,), w2, b2, 1e-6).view(B, n_tok, D)[:, 1:].reshape(-1, D)
    r2 = _cmp("tap", tap, ref2, 2e-2, 1e-2)
    # the separate tap pass (what the last block still uses) must give the same bits
    x2 = x.clone()
    tap_alone = ops.layernorm(x2, w2, b2, 1e-6, n_tok, True, delta=d1, write_x=False, delta2=d2)
    torch.cuda.synchronize()
    r["tap_ok"] = r2["ok"]
    r["tap_same_bits_as_separate_pass"] = bool(torch.equal(tap, tap_alone))
    r["x_written"] = bool(torch.equal(x1, xs))
    r["ok"] = r["ok"] and r2["ok"] and r["tap_same_bits_as_separate_pass"] and r["x_written"]
    return r


def chk_channel_ln(C):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn(2, 9, 11, C, generator=g, device="cuda") * 2 + 0.3).bfloat16()
    w = torch.randn(C, generator=g, device="cuda")
    b = torch.randn(C, generator=g, device="cuda")
    out = ops.channel_ln_relu(x, w, b)
    torch.cuda.synchronize()
    xf = x.float()
    u = xf.mean(-1, keepdim=True)
    s = (xf - u).pow(2).mean(-1, keepdim=True)
    ref = torch.relu((xf - u) / torch.sqrt(s + 1e-6) * w + b)
    return _cmp("cln", out, ref, 2e-2, 1e-2)


def chk_upsample(Hi, Ho, C=64):
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(2, Hi, Hi, C, generator=g, device="cuda").bfloat16()
    out = ops.upsample(x, Ho, Ho)
    torch.cuda.synchronize()
    ref = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), (Ho, Ho), mode="bilinear", align_corners=True)
    return _cmp("up", out.permute(0, 3, 1, 2), ref, 2e-2, 1e-2)


def chk_fused_tail(gh, gw):
    """tap GEMM at low res + gather == conv2(interp(y)) chain of dpt.py:194-195."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(17)
    B, Cm = 2, 64
    Hl, Wl, H, W = 8 * gh, 8 * gw, 14 * gh, 14 * gw
    y = torch.randn(B, Cm, Hl, Wl, generator=g, device="cuda").bfloat16()
    w2 = torch.randn(32, Cm, 3, 3, generator=g, device="cuda") * (1.0 / (3 * Cm ** 0.5))
    b2 = torch.randn(32, generator=g, device="cuda") * 0.1
    w3 = torch.randn(32, generator=g, device="cuda") * 0.3
    b3 = torch.randn(1, generator=g, device="cuda") * 0.1
    up = torch.nn.functional.interpolate(y.float(), (H, W), mode="bilinear", align_corners=True)
    z = torch.relu(torch.nn.functional.conv2d(up, w2.bfloat16().float(), b2, padding=1))
    ref = torch.sigmoid((z * w3.view(1, 32, 1, 1)).sum(1) + b3)
    wt = ops.pack_tail_taps(w2)
    ya = y.permute(0, 2, 3, 1).reshape(B * Hl * Wl, Cm).contiguous()
    V = torch.zeros(B, Hl, Wl, 288, dtype=torch.float16, device="cuda")
    ops.gemm(ya, wt, epi=L.EPI_F16, out_bf16=V, ldo=288)
    out = ops.tail_gather(V, b2, torch.cat([w3, b3]).contiguous(), H, W, True)
    torch.cuda.synchronize()
    return _cmp("fused_tail", out, ref, 5e-3, 0)


def chk_tail_mma(gh, gw, Cm=64, B=2, sigmoid=1):
    """upsample + output_conv2 on tensor cores (tail_mma_kernel) == conv2(interp(y)) chain of dpt.py:194-195."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(19)
    Hl, Wl, H, W = 8 * gh, 8 * gw, 14 * gh, 14 * gw
    y = (torch.randn(B, Cm, Hl, Wl, generator=g, device="cuda") * 1.5).half()
    w2 = torch.randn(32, Cm, 3, 3, generator=g, device="cuda") * (1.0 / (3 * Cm ** 0.5))
    b2 = torch.randn(32, generator=g, device="cuda") * 0.1
    w3 = torch.randn(32, generator=g, device="cuda") * 0.3
    b3 = torch.randn(1, generator=g, device="cuda") * 0.1
    up = torch.nn.functional.interpolate(y.float(), (H, W), mode="bilinear", align_corners=True)
    z = torch.relu(torch.nn.functional.conv2d(up, w2.half().float(), b2, padding=1))
    ref = (z * w3.view(1, 32, 1, 1)).sum(1) + b3
    ref = torch.sigmoid(ref) if sigmoid == 1 else torch.relu(ref) if sigmoid == 2 else ref
    wpk = ops.pack_tail_mma(w2)
    ya = y.permute(0, 2, 3, 1).contiguous()
    out = ops.tail_mma(ya, wpk, b2, torch.cat([w3, b3]).contiguous(), H, W, sigmoid)
    torch.cuda.synchronize()
    return _cmp("tail_mma", out, ref, 5e-3, 0)


def chk_patch_gather():
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, W = 2, 70, 98
    rgb = torch.rand(B, 3, H, W, generator=g, device="cuda")
    m = torch.rand(B, 1, H, W, generator=g, device="cuda") * 2 - 1
    o = torch.rand(B, 1, H, W, generator=g, device="cuda") * 2 - 1
    out = ops.patch_gather(rgb, [m, o], 1024)
    torch.cuda.synchronize()
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    x = torch.cat([(rgb - mean) / std, m, o], 1)
    ref = torch.nn.functional.unfold(x, 14, stride=14).transpose(1, 2).reshape(B * (H // 14) * (W // 14), 5 * 196)
    r = _cmp("gather", out[:, :980], ref, 1e-6, 2 ** -8)
    r["pad_zero"] = bool((out[:, 980:] == 0).all().item())
    r["ok"] = r["ok"] and r["pad_zero"]
    return r


def chk_conv_s2(B, H, W, Cin, Cout):
    """3x3, stride 2, pad 1 (resize_layers[3], dpt.py:102-107) as an implicit GEMM: the tensor map walks the input with
    element stride 2; odd and even input sizes, ragged output tiles."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(10)
    x = torch.randn(B, Cin, H, W, generator=g, device="cuda").bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, generator=g, device="cuda") * (1.0 / (3 * Cin ** 0.5))
    bias = torch.randn(Cout, generator=g, device="cuda") * 0.1
    ref = torch.nn.functional.conv2d(x.float(), w.bfloat16().float(), bias, stride=2, padding=1)
    Ho, Wo = ref.shape[2:]
    out = torch.zeros(B, Ho, Wo, Cout, dtype=torch.bfloat16, device="cuda")
    ops.gemm(x.permute(0, 2, 3, 1).contiguous(), ops.pack_conv3x3(w), conv=(B, H, W, Cin), N=Cout, bias=bias, out_bf16=out,
             ldo=Cout, conv_stride=2)
    torch.cuda.synchronize()
    return _cmp("conv_s2", out.permute(0, 3, 1, 2), ref, 3e-2, 1e-2)


def chk_gemm_resid_f32(M, N, K, gamma=True, cg=0):
    """EPI_RESID_F32: x (fp32, in place) += gamma * (A W^T + b) -- the residual update of block.py:105-106."""
    torch, L, ops = _imports()
    g = torch.Generator(device="cuda").manual_seed(23)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(N, K, generator=g, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    gam = (1.0 + 0.1 * torch.randn(N, generator=g, device="cuda")) if gamma else None
    x = torch.randn(M, N, generator=g, device="cuda") * 2
    ref = A.float() @ Wt.float().t() + bias
    if gamma:
        ref = ref * gam
    ref = x + ref
    out = x.clone()
    ops.gemm(A, Wt, epi=L.EPI_RESID_F32, bias=bias, gamma=gam, out_f32=out, ldo=N, force_cg=cg)
    torch.cuda.synchronize()
    return _cmp("resid_f32", out, ref, 2e-3, 2e-3)


CHECKS = {
    "gemm_resid_f32": lambda: chk_gemm_resid_f32(1370 * 3, 1024, 1024),
    "gemm_resid_f32_cg2": lambda: chk_gemm_resid_f32(1370 * 3, 1024, 4096, cg=2),
    "gemm_resid_f32_ragged": lambda: chk_gemm_resid_f32(777, 384, 1536, gamma=False),
    "gemm_resid_f32_small": lambda: chk_gemm_resid_f32(100, 96, 64),
    "gemm_small_bn128": lambda: chk_gemm(300, 256, 128, 128, "bias"),
    "gemm_small_bn256": lambda: chk_gemm(300, 256, 128, 256, "bias"),
    "gemm_small_bn64": lambda: chk_gemm(300, 256, 192, 64, "bias"),
    "gemm_qkv_shape": lambda: chk_gemm(1370 * 2, 3072, 1024, 0, "bias"),
    "gemm_fc1_gelu": lambda: chk_gemm(1370, 4096, 1024, 0, "gelu"),
    "gemm_fc2_resid": lambda: chk_gemm(1370, 1024, 4096, 0, "resid"),
    "gemm_ragged_n": lambda: chk_gemm(500, 48, 384, 0, "bias"),
    "gemm_swiglu": lambda: chk_gemm(700, 1024, 256, 0, "swiglu"),
    "gemm_embed": chk_embed,
    "gemm_cg2_small": lambda: chk_gemm(300, 256, 128, 256, "bias", cg=2),
    "gemm_cg2_bn128": lambda: chk_gemm(700, 384, 192, 128, "bias", cg=2),
    "gemm_cg2_qkv": lambda: chk_gemm(1370 * 2, 3072, 1024, 0, "bias", cg=2),
    "gemm_cg2_gelu": lambda: chk_gemm(1370, 4096, 1024, 0, "gelu", cg=2),
    "gemm_cg2_gamma": lambda: chk_gemm(1370, 1024, 4096, 0, "resid", cg=2),
    "gemm_cg2_swiglu": lambda: chk_gemm(700, 1024, 256, 0, "swiglu", cg=2),
    "conv_cg2_plain": lambda: chk_conv(2, 37, 45, 128, 256, "plain", cg=2),
    "conv_cg2_rcu": lambda: chk_conv(2, 19, 33, 64, 128, "rcu", cg=2),
    "conv_plain_37": lambda: chk_conv(2, 37, 37, 128, 64, "plain"),
    "conv_c48": lambda: chk_conv(1, 20, 33, 48, 48, "plain"),
    "conv_rcu": lambda: chk_conv(2, 19, 19, 64, 64, "rcu"),
    "conv_tail": lambda: chk_conv(1, 70, 84, 64, 32, "tail"),
    "convT_k4": lambda: chk_convT(4),
    "convT_k2": lambda: chk_convT(2),
    # Cout % 64 == 0: pixel shuffle through the 5-D TMA store (warps whose 32 pixels span 1, 2 and up to 7 image rows)
    "convT_k4_c256_tma": lambda: chk_convT(4, 256, 256, pix=True),
    "convT_k2_c512_tma": lambda: chk_convT(2, 128, 512, pix=True),
    "convT_k4_c64_w5_tma": lambda: chk_convT(4, 64, 64, B=3, H=7, W=5, pix=True),
    "convT_k2_c128_w49_tma": lambda: chk_convT(2, 192, 128, B=1, H=37, W=49, pix=True),
    "convT_k2_c256_linear": lambda: chk_convT(2, 256, 256),
    # both attention kernels on every shape (impl 0 = attention.cuh, 1 / 2 = attention2.cuh one / two threads per row), then the auto-selected one
    **{f"attention{i}_small": (lambda i=i: chk_attention(1, 128, 1, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_ragged": (lambda i=i: chk_attention(1, 200, 2, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_1370": (lambda i=i: chk_attention(2, 1370, 6, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_rescale": (lambda i=i: chk_attention(1, 1370, 2, grow=True, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_5477": (lambda i=i: chk_attention(1, 5477, 2, impl=i)) for i in (0, 1, 2)},
    # persistent kernel: several work units per CTA, full (two query tiles) and short (one tile) units interleaved
    **{f"attention{i}_multiunit_1370": (lambda i=i: chk_attention(4, 1370, 16, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_multiunit_300": (lambda i=i: chk_attention(8, 300, 16, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_multiunit_128": (lambda i=i: chk_attention(10, 128, 16, impl=i)) for i in (0, 1, 2)},
    **{f"attention{i}_multiunit_rescale": (lambda i=i: chk_attention(3, 700, 16, grow=True, impl=i)) for i in (0, 1, 2)},
    "attention_auto_b32": lambda: chk_attention(32, 1370, 16),
    "attention_impls_agree_1370": lambda: chk_attention_impls_agree(3, 1370, 16),
    "attention_impls_agree_rescale": lambda: chk_attention_impls_agree(2, 700, 16, grow=True),
    "attention_impls_agree_5477": lambda: chk_attention_impls_agree(1, 5477, 4),
    "layernorm_384": lambda: chk_layernorm(384, False),
    "layernorm_1024_drop": lambda: chk_layernorm(1024, True),
    "layernorm_1536": lambda: chk_layernorm(1536, False),
    "layernorm_delta": chk_layernorm_delta,
    "layernorm_tap_fused": chk_layernorm_tap,
    "channel_ln_48": lambda: chk_channel_ln(48),
    "channel_ln_1024": lambda: chk_channel_ln(1024),
    "upsample_19_37": lambda: chk_upsample(19, 37),
    "upsample_37_74": lambda: chk_upsample(37, 74),
    "upsample_148_296_c256": lambda: chk_upsample(148, 296, 256),
    "upsample_40_70_c128": lambda: chk_upsample(40, 70, 128),   # not a 2x map: strips cross source rows irregularly
    "channel_ln_256": lambda: chk_channel_ln(256),
    "channel_ln_512": lambda: chk_channel_ln(512),
    "channel_ln_1536": lambda: chk_channel_ln(1536),
    "fused_tail_5x7": lambda: chk_fused_tail(5, 7),
    "fused_tail_9x9": lambda: chk_fused_tail(9, 9),
    "tail_mma_5x7_c64": lambda: chk_tail_mma(5, 7, 64),
    "tail_mma_9x9_c128": lambda: chk_tail_mma(9, 9, 128),
    "tail_mma_3x11_c32_logit": lambda: chk_tail_mma(3, 11, 32, B=3, sigmoid=0),
    "tail_mma_37x37_c128": lambda: chk_tail_mma(37, 37, 128, B=2),
    "tail_mma_1x1_c64_relu": lambda: chk_tail_mma(1, 1, 64, B=5, sigmoid=2),
    "patch_gather": chk_patch_gather,
    "conv_chln_256": lambda: chk_conv(2, 37, 45, 256, 256, "chln"),
    "conv_chln_256_cg2": lambda: chk_conv(2, 40, 70, 256, 256, "chln", cg=2),
    "conv_chln_48": lambda: chk_conv(2, 20, 33, 48, 48, "chln"),
    "conv_chln_192": lambda: chk_conv(1, 37, 37, 192, 192, "chln"),
    "conv_s2_37": lambda: chk_conv_s2(2, 37, 37, 128, 192),
    "conv_s2_74x50": lambda: chk_conv_s2(1, 74, 50, 64, 64),
    "conv_s2_1024": lambda: chk_conv_s2(2, 37, 37, 1024, 1024),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_check.json"))
    a = ap.parse_args()
    if a.one:
        try:
            r = CHECKS[a.one]()
        except Exception as e:  # noqa: BLE001
            r = {"ok": False, "exception": repr(e)[:500]}
            try:
                _, L, _ = _imports()
                r["device_error"] = [hex(v) for v in L.device_error()]
            except Exception:  # noqa: BLE001
                pass
        print("RESULT " + json.dumps(r))
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results = {}
    names = [n for n in CHECKS if a.only in n]
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True,
                               timeout=180)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            r = json.loads(line[-1][7:]) if line else {"ok": False, "rc": p.returncode, "stderr": p.stderr[-800:]}
        except subprocess.TimeoutExpired:
            r = {"ok": False, "timeout": True}
        r["sec"] = round(time.time() - t0, 1)
        results[n] = r
        print(n, json.dumps(r), flush=True)
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)
    nok = sum(1 for r in results.values() if r.get("ok"))
    print(f"SUMMARY {nok}/{len(results)} ok")


if __name__ == "__main__":
    main()
