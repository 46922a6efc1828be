"""Numerics model of the tensor-core path (TEST INFRASTRUCTURE).

The oracle (kl_f8.py) with 16-bit rounding inserted at exactly the points where
the CUDA path stores a GEMM operand (DESIGN.md "precision plan"): GroupNorm+SiLU
outputs, conv1 outputs, the 16-bit copy of x that feeds downsample / nin_shortcut,
q|k, V^T, softmax probabilities, attention output, and all GEMM weights.
Accumulation, bias, residual stream and GroupNorm statistics stay fp32 (fp64
stats in the kernels).  It separates *kernel bugs* (CUDA result far from this
model) from the *inherent operand-rounding floor* (this model vs the oracle):
with seeded random-init weights the bf16 floor is 0.8-1.4e-2 relative L2 on the
latent mean, i.e. it straddles the north-star 1e-2 gate by itself, while fp16
operands (same tcgen05 kind::f16 rate) sit near 2e-3.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kl_f8


def rounder(fmt):
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    return lambda t: t.to(dt).float()


def encode_moments(x, sd, fmt="bf16"):
    r = rounder(fmt)
    p = "encoder."

    def conv(a16, name, stride=1, pad=0):
        return F.conv2d(a16, r(sd[name + ".weight"]), sd[name + ".bias"], stride=stride, padding=pad)

    def gn(t, name, silu=True, stats_from=None):
        # statistics come from the producer's fp32 values (conv epilogue), the normalised tensor
        # is what was stored (16-bit for conv1's output)
        src = t if stats_from is None else stats_from
        b_, c_ = src.shape[:2]
        g = src.reshape(b_, 32, -1).double()
        mean = g.mean(-1); var = (g * g).mean(-1) - mean * mean
        rstd = (1.0 / torch.sqrt(var + 1e-6)).float().repeat_interleave(c_ // 32, 1)[:, :, None, None]
        mean = mean.float().repeat_interleave(c_ // 32, 1)[:, :, None, None]
        sc = sd[name + ".weight"][None, :, None, None] * rstd
        y = t * sc + (sd[name + ".bias"][None, :, None, None] - mean * sc)
        return r(y * torch.sigmoid(y) if silu else y)

    def res(t, n):
        h32 = conv(gn(t, n + ".norm1"), n + ".conv1", 1, 1)
        h = conv(gn(r(h32), n + ".norm2", stats_from=h32), n + ".conv2", 1, 1)
        if (n + ".nin_shortcut.weight") in sd:
            t = conv(r(t), n + ".nin_shortcut")
        return t + h

    with torch.no_grad():
        h = kl_f8.conv(x, sd, p + "conv_in", 1, 1)                       # fp32 CUDA-core kernel
        for lvl in range(4):
            for b in range(2):
                h = res(h, p + f"down.{lvl}.block.{b}")
            if lvl != 3:
                h = conv(F.pad(r(h), (0, 1, 0, 1)), p + f"down.{lvl}.downsample.conv", 2, 0)
        h = res(h, p + "mid.block_1")
        a = p + "mid.attn_1"
        hn = gn(h, a + ".norm", silu=False)
        q = r(conv(hn, a + ".q")); k = r(conv(hn, a + ".k"))
        vT = r(F.conv2d(hn, r(sd[a + ".v.weight"]), None))               # bias added after P V
        b_, c, hh, ww = q.shape
        s = torch.bmm(q.reshape(b_, c, -1).permute(0, 2, 1), k.reshape(b_, c, -1)) * (int(c) ** -0.5)
        pr = r(F.softmax(s, dim=2))
        o = torch.bmm(vT.reshape(b_, c, -1), pr.permute(0, 2, 1)) + sd[a + ".v.bias"][None, :, None]
        o = r(o).reshape(b_, c, hh, ww)
        h = h + conv(o, a + ".proj_out")
        h = res(h, p + "mid.block_2")
        # conv_out with quant_conv folded into the weights (exact in fp64, then rounded once)
        wq = sd["quant_conv.weight"].double().reshape(8, 8)
        wf = torch.einsum("om,mikl->oikl", wq, sd[p + "conv_out.weight"].double()).float()
        bf = (wq @ sd[p + "conv_out.bias"].double() + sd["quant_conv.bias"].double()).float()
        return F.conv2d(gn(h, p + "norm_out"), r(wf), bf, padding=1)
