"""Numerics model of the tensor-core path (TEST INFRASTRUCTURE).

The oracle (kl_f8.py) with 16-bit rounding inserted at exactly the points where
the CUDA path stores a GEMM operand (DESIGN.md "precision plan"), everything else
exact.  Operand classes and their formats per precision mode:

    class   what                                               bf16   fp16   mixed
    w       all GEMM weights                                   bf16   fp16   fp16 (x 2^k per layer, exact)
    act     GroupNorm(+SiLU) outputs                           bf16   fp16   fp16
    h       conv1 outputs (norm2 inputs)                       bf16   fp16   fp16
    xc      16-bit copies of x feeding downsample / nin        bf16   fp16   (x itself, below)
    x       the residual stream                                fp32   fp32   fp16 of x * 2^-6
    qk v p o  q|k, V^T, softmax probabilities, attention out   bf16   fp16   bf16
              (and proj_out's weights)

Accumulation, bias, residual stream and GroupNorm statistics stay fp32 (fp64
stats in the kernels).  The model separates *kernel bugs* (CUDA result far from
this model) from the *inherent operand-rounding floor* (this model vs the
oracle).  With seeded random-init weights the floor on the latent mean is
0.9-1.6e-2 for bf16 (it straddles the north-star 1e-2 gate by itself: weights,
GroupNorm outputs, conv1 outputs and x copies each contribute 4-9e-3), 1.7-1.9e-3
for fp16 and 2.1-2.4e-3 for mixed (the bf16 attention operands add < 6e-4, the
16-bit residual stream ~4e-4).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kl_f8

CLASSES = ("w", "act", "h", "xc", "qk", "v", "p", "o")
XC_SCALE = 2.0 ** -6


def formats(mode):
    """mode name or {class: "bf16"|"fp16"|"f32"} -> full per-class dict."""
    if isinstance(mode, dict):
        return {k: mode.get(k, "f32") for k in CLASSES}
    if mode in ("bf16", "fp16", "f32"):
        return {k: mode for k in CLASSES}
    if mode == "mixed":
        d = {k: "fp16" for k in CLASSES}
        d.update(qk="bf16", v="bf16", p="bf16", o="bf16")
        return d
    raise ValueError(mode)


def rounder(fmt, scale=1.0):
    if fmt == "f32":
        return lambda t: t
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    if scale == 1.0:
        return lambda t: t.to(dt).float()
    return lambda t: (t * scale).to(dt).float() / scale


def encode_moments(x, sd, fmt="bf16", stream16=None):
    """stream16: the residual stream x itself is stored in 16 bit (fp16 of x * 2^-6; default: on in mixed mode, as in
    the CUDA path); GroupNorm statistics of x still come from the producer's fp32 values."""
    f = formats(fmt)
    r = {k: rounder(f[k], XC_SCALE if (k == "xc" and fmt == "mixed") else 1.0) for k in CLASSES}
    if stream16 is None:
        stream16 = fmt == "mixed"
    rs = rounder("fp16", XC_SCALE) if stream16 else (lambda t: t)
    xc = (lambda t: t) if stream16 else r["xc"]          # with a 16-bit stream x is its own operand copy
    p = "encoder."

    def conv(a16, name, stride=1, pad=0):
        return F.conv2d(a16, r["w"](sd[name + ".weight"]), sd[name + ".bias"], stride=stride, padding=pad)

    def gn(t, name, silu=True, stats_from=None):
        # statistics come from the producer's fp32 values (conv epilogue), the normalised tensor
        # is what was stored (16-bit for conv1's output and for a 16-bit stream)
        src = t if stats_from is None else stats_from
        b_, c_ = src.shape[:2]
        g = src.reshape(b_, 32, -1).double()
        mean = g.mean(-1); var = (g * g).mean(-1) - mean * mean
        rstd = (1.0 / torch.sqrt(var + 1e-6)).float().repeat_interleave(c_ // 32, 1)[:, :, None, None]
        mean = mean.float().repeat_interleave(c_ // 32, 1)[:, :, None, None]
        sc = sd[name + ".weight"][None, :, None, None] * rstd
        y = t * sc + (sd[name + ".bias"][None, :, None, None] - mean * sc)
        return r["act"](y * torch.sigmoid(y) if silu else y)

    def res(t, t32, n):
        """t: the stored stream, t32: its fp32 value as produced (statistics source) -> (stored, fp32) of the output"""
        h32 = conv(gn(t, n + ".norm1", stats_from=t32), n + ".conv1", 1, 1)
        h = conv(gn(r["h"](h32), n + ".norm2", stats_from=h32), n + ".conv2", 1, 1)
        skip = t
        if (n + ".nin_shortcut.weight") in sd:
            skip = conv(xc(t), n + ".nin_shortcut")
        o32 = skip + h
        return rs(o32), o32

    with torch.no_grad():
        h32 = kl_f8.conv(x, sd, p + "conv_in", 1, 1)                     # exact integer operand, hi+lo split weights
        h = rs(h32)
        for lvl in range(4):
            for b in range(2):
                h, h32 = res(h, h32, p + f"down.{lvl}.block.{b}")
            if lvl != 3:
                h32 = conv(F.pad(xc(h), (0, 1, 0, 1)), p + f"down.{lvl}.downsample.conv", 2, 0)
                h = rs(h32)
        h, h32 = res(h, h32, p + "mid.block_1")
        a = p + "mid.attn_1"
        hn = gn(h, a + ".norm", silu=False, stats_from=h32)
        q = r["qk"](conv(hn, a + ".q")); k = r["qk"](conv(hn, a + ".k"))
        vT = r["v"](F.conv2d(hn, r["w"](sd[a + ".v.weight"]), None))     # bias added after P V
        b_, c, hh, ww = q.shape
        s = torch.bmm(q.reshape(b_, c, -1).permute(0, 2, 1), k.reshape(b_, c, -1)) * (int(c) ** -0.5)
        pr = r["p"](F.softmax(s, dim=2))
        o = torch.bmm(vT.reshape(b_, c, -1), pr.permute(0, 2, 1)) + sd[a + ".v.bias"][None, :, None]
        o = r["o"](o).reshape(b_, c, hh, ww)
        # proj_out's weights share the attention output's format (one 16-bit format per tcgen05 GEMM)
        h32 = h + F.conv2d(o, r["o"](sd[a + ".proj_out.weight"]), sd[a + ".proj_out.bias"])
        h = rs(h32)
        h, h32 = res(h, h32, p + "mid.block_2")
        # conv_out with quant_conv folded into the weights (exact in fp64, then rounded once)
        wq = sd["quant_conv.weight"].double().reshape(8, 8)
        wf = torch.einsum("om,mikl->oikl", wq, sd[p + "conv_out.weight"].double()).float()
        bf = (wq @ sd[p + "conv_out.bias"].double() + sd["quant_conv.bias"].double()).float()
        return F.conv2d(gn(h, p + "norm_out", stats_from=h32), r["w"](wf), bf, padding=1)
