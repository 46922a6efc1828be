"""Oracle: explicit-loop numpy restatements of the third-party primitives the
reference calls (PyTorch ATen: conv2d, group_norm, softmax attention via bmm,
LSTM cell).  TEST INFRASTRUCTURE; small cases only.  They follow the operators'
published definitions (torch.nn.Conv2d / GroupNorm / LSTM documentation) and pin
the ``torch.nn.functional`` calls that ``kl_f8.py`` / ``rbvae.py`` are written on.
"""
from __future__ import annotations

import numpy as np


def conv2d(x, w, b, stride=1, pad=(0, 0, 0, 0)):
    """x [N,C,H,W], w [O,C,kh,kw]; pad = (left, right, top, bottom) zeros (F.pad order).
    out[n,o,y,x] = b[o] + sum_{c,r,s} w[o,c,r,s] * xpad[n,c,y*stride+r,x*stride+s]"""
    x = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad[2], pad[3]), (pad[0], pad[1])))
    N, C, H, W = x.shape
    O, _, kh, kw = w.shape
    Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    out = np.zeros((N, O, Ho, Wo))
    for r in range(kh):
        for s in range(kw):
            patch = x[:, :, r:r + stride * Ho:stride, s:s + stride * Wo:stride]     # [N,C,Ho,Wo]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, r, s].astype(np.float64))
    return out + b.astype(np.float64)[None, :, None, None]


def group_norm(x, gamma, beta, groups=32, eps=1e-6):
    """y = (x - mean_g) / sqrt(var_g + eps) * gamma + beta, biased variance over (C/G, H, W)."""
    N, C, H, W = x.shape
    xg = x.astype(np.float64).reshape(N, groups, -1)
    mean = xg.mean(-1, keepdims=True)
    var = xg.var(-1, keepdims=True)
    y = ((xg - mean) / np.sqrt(var + eps)).reshape(N, C, H, W)
    return y * gamma[None, :, None, None] + beta[None, :, None, None]


def silu(x):
    return x / (1.0 + np.exp(-x))


def attention(q, k, v):
    """q,k,v [N,C,L] as in AttnBlock.forward (model.py:185-198): out[n,c,j] = sum_i v[n,c,i] P[n,j,i],
    P = softmax_i(q[:, :, j] . k[:, :, i] / sqrt(C))."""
    C = q.shape[1]
    s = np.einsum("ncj,nci->nji", q.astype(np.float64), k.astype(np.float64)) * (int(C) ** -0.5)
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(-1, keepdims=True)
    return np.einsum("nci,nji->ncj", v.astype(np.float64), p)


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTM cell, gate order (i, f, g, o)."""
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))
    gates = x @ w_ih.T + b_ih + h @ w_hh.T + b_hh
    i, f, g, o = np.split(gates, 4, axis=-1)
    c2 = sig(f) * c + sig(i) * np.tanh(g)
    return sig(o) * np.tanh(c2), c2
