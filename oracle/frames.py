"""Oracle: frame preparation.  TEST INFRASTRUCTURE.

``load_img`` restates src/stable-diffusion/get_percep_embeddings.py:48-71
(= load_img_for_sd, scripts/evaluation/state_consistency_eval/embedding_matching.py:318-338)
with PIL itself.  ``lanczos_resize_u8`` is an independent numpy restatement of
the third-party arithmetic behind ``Image.resize(..., LANCZOS)`` -- Pillow
(pillow==10.2.0 pinned at reference requirements.txt:113; algorithm published in
Pillow's src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc / Vertical_8bpc) -- which the tests pin against
PIL on random and real frames.  The CUDA kernel (csrc/resize.cu) must match it
bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter
    (support 3).  Returns bounds [out,2] (xmin, count) and int32 coeffs [out,ksize]."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, out_size, axis):
    """One 8-bit pass along `axis` of an [H,W,C] uint8 image."""
    in_size = img.shape[axis]
    bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)           # [in, ...]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        xmin, n = bounds[o]
        acc = np.tensordot(kk[o, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def lanczos_resize_u8(img: np.ndarray, H: int, W: int) -> np.ndarray:
    """[Hs,Ws,3] uint8 -> [H,W,3] uint8 exactly as PIL Image.resize((W,H), LANCZOS):
    horizontal pass first, uint8 rounding between passes, a pass is skipped when
    that dimension is unchanged."""
    out = img
    if img.shape[1] != W:
        out = _pass(out, W, 1)
    if img.shape[0] != H:
        out = _pass(out, H, 0)
    return out


def load_img_pil(image, target=(1280, 720)):
    """get_percep_embeddings.py:54-71 on an in-memory PIL image: RGB, LANCZOS to
    1280x720, LANCZOS again to multiples of 32 (1280x704), /255, NCHW, 2x-1.
    Returns (float32 [1,3,H,W], uint8 [H,W,3])."""
    import PIL
    import torch
    image = image.convert("RGB")
    image = image.resize(target, resample=PIL.Image.LANCZOS)
    w, h = target
    w, h = map(lambda x: x - x % 32, (w, h))
    if (w, h) != target:
        image = image.resize((w, h), resample=PIL.Image.LANCZOS)
    u8 = np.array(image)
    arr = u8.astype(np.float32) / 255.0
    arr = arr[None].transpose(0, 3, 1, 2)
    return 2. * torch.from_numpy(arr) - 1., u8


def normalise_u8(frames_u8):
    """uint8 [N,H,W,3] -> float32 [N,3,H,W], the arithmetic of get_percep_embeddings.py:67-71."""
    import torch
    arr = np.asarray(frames_u8).astype(np.float32) / 255.0
    return 2. * torch.from_numpy(arr.transpose(0, 3, 1, 2).copy()) - 1.


def synthetic_frames(n, H, W, seed=1234, smooth=False):
    """SURVEY 8d synthetic inputs: i.i.d. uniform uint8, or a low-pass-filtered variant."""
    import torch
    g = torch.Generator().manual_seed(seed)
    f = torch.randint(0, 256, (n, H, W, 3), generator=g, dtype=torch.uint8)
    if smooth:
        import torch.nn.functional as F
        x = f.permute(0, 3, 1, 2).float()
        x = F.avg_pool2d(F.pad(x, (4, 4, 4, 4), mode="reflect"), 9, 1)
        x = (x - x.amin()) / (x.amax() - x.amin()) * 255.0
        f = x.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    return f.numpy()
