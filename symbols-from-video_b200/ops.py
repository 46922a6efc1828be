"""Single-operator entry points (sfv_op_* in include/sfv.h): each hot kernel
callable on its own, so the parity tests can check every distinct layer shape
against the oracle (SURVEY 8c "per-layer-shape known answers")."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _host_f32(t):
    return np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())


def conv2d_nhwc(x, weight, bias, stride=1, pad=(1, 1), residual=None, relu=False, precision="fp32"):
    """x fp32 NHWC cuda, weight OIHW (host or device), bias [O] -> y fp32 NHWC."""
    _lib.require_cuda(x, "x")
    N, H, W, Cin = x.shape
    Cout, _, ks, _ = weight.shape
    Ho = (H + pad[0] + pad[1] - ks) // stride + 1
    Wo = (W + pad[0] + pad[1] - ks) // stride + 1
    y = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32, device=x.device)
    w = _host_f32(weight); b = _host_f32(bias)
    x = x.contiguous()
    if residual is not None:
        residual = residual.contiguous()
    _lib.run(_lib.lib().sfv_op_conv2d, x, _lib.ptr(x), w.ctypes.data, b.ctypes.data, _lib.ptr(residual), _lib.ptr(y),
                                        N, H, W, Cin, Cout, ks, stride, pad[0], pad[1], int(relu),
                                        _lib.PRECISIONS[precision])
    return y


def conv_in_u8(frames, weight, bias, precision="fp32"):
    """uint8 [N,H,W,3] cuda, conv_in weight [128,3,3,3] / bias [128] -> y fp32 NHWC [N,H,W,128]
    (= conv3x3(2*u/255-1) + bias)."""
    _lib.require_cuda(frames, "frames")
    N, H, W, _ = frames.shape
    y = torch.empty(N, H, W, 128, dtype=torch.float32, device=frames.device)
    w = _host_f32(weight); b = _host_f32(bias)
    _lib.run(_lib.lib().sfv_op_conv_in_u8, frames, _lib.ptr(frames.contiguous()), w.ctypes.data, b.ctypes.data, _lib.ptr(y),
                                            N, H, W, _lib.PRECISIONS[precision])
    return y


def group_norm_nhwc(x, gamma, beta, groups=32, eps=1e-6, silu=False):
    """x fp32 [N,HW,C] (or [N,H,W,C]) cuda -> same shape."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    y = torch.empty_like(x)
    g = gamma.to(x.device, torch.float32).contiguous(); b = beta.to(x.device, torch.float32).contiguous()
    _lib.run(_lib.lib().sfv_op_group_norm, x, _lib.ptr(x), _lib.ptr(g), _lib.ptr(b), _lib.ptr(y), N, HW, C, groups,
                                            float(eps), int(silu))
    return y


def attention(q, k, v, scale=None, precision="fp32"):
    """q,k,v fp32 [N,L,C] cuda -> softmax(q k^T scale) v."""
    _lib.require_cuda(q, "q")
    N, L, C = q.shape
    scale = C ** -0.5 if scale is None else scale
    out = torch.empty_like(q)
    _lib.run(_lib.lib().sfv_op_attention, q, _lib.ptr(q.contiguous()), _lib.ptr(k.contiguous()), _lib.ptr(v.contiguous()),
                                           _lib.ptr(out), N, L, C, float(scale), _lib.PRECISIONS[precision])
    return out


def resize_lanczos(frames, H, W, want_float=False):
    """uint8 [B,Hs,Ws,3] cuda -> uint8 [B,H,W,3] (PIL LANCZOS arithmetic), optionally also
    the normalised fp32 NCHW tensor 2*(v/255)-1."""
    import ctypes as C
    _lib.require_cuda(frames, "frames")
    frames = frames.contiguous()
    B, Hs, Ws, _ = frames.shape
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=frames.device)
    f = torch.empty(B, 3, H, W, dtype=torch.float32, device=frames.device) if want_float else None
    nb = C.c_size_t()
    _lib.check(_lib.lib().sfv_resize_workspace_bytes(B, Hs, Ws, H, W, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=frames.device)
    _lib.run(_lib.lib().sfv_resize_normalise, frames, _lib.ptr(frames), B, Hs, Ws, H, W, _lib.ptr(f), _lib.ptr(out),
                                               _lib.ptr(ws), nb.value)
    return (out, f) if want_float else out
