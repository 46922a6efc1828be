"""Forward values of the reference's RBVAE training losses, computed on the device through the C ABI.

Mirrors models/percep_RBVAE/percep_RBVAE_train.py:27-107 (the contrastive trainer defines the same functions):
same names, arguments and defaults; every function returns a 0-dim float32 CUDA tensor.  Forward only -- the
reference differentiates through them, this package does not train.
"""
from __future__ import annotations

import torch

from . import _lib


def _f32(t, what):
    _lib.require_cuda(t, what)
    return t.to(torch.float32).contiguous()


def _out(ref):
    return torch.empty((), dtype=torch.float32, device=ref.device)


def l1_loss(q_logits, lamb):
    """:27-29  lamb * torch.norm(q_logits, p=1)"""
    q = _f32(q_logits, "q_logits")
    out = _out(q)
    _lib.run(_lib.lib().sfv_loss_l1, q, _lib.ptr(q), q.numel(), float(lamb), _lib.ptr(out))
    return out


def recon_loss(x_recon, x):
    """:32-33  F.mse_loss(x_recon, x)"""
    a, b = _f32(x_recon, "x_recon"), _f32(x, "x")
    if a.shape != b.shape:
        raise ValueError(f"shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    out = _out(a)
    _lib.run(_lib.lib().sfv_loss_mse, a, _lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(out))
    return out


def triplet_loss(anchor, pos, neg, margin=1.0, p=2.0, eps=1e-08, swap=True, size_average=None, reduce=None,
                 reduction="mean"):
    """:35-49  F.triplet_margin_loss(...); p = 2 and reduction = 'mean' (the reference's own defaults) only."""
    if p != 2.0 or reduction != "mean" or size_average is not None or reduce is not None:
        raise NotImplementedError("triplet_loss: only p=2, reduction='mean' (the reference's defaults)")
    a, po, n = _f32(anchor, "anchor"), _f32(pos, "pos"), _f32(neg, "neg")
    if not (a.dim() == 2 and a.shape == po.shape == n.shape):
        raise ValueError("triplet_loss expects three [N, D] tensors of one shape")
    out = _out(a)
    _lib.run(_lib.lib().sfv_loss_triplet, a, _lib.ptr(a), _lib.ptr(po), _lib.ptr(n), a.shape[0], a.shape[1], float(margin),
             float(eps), int(bool(swap)), _lib.ptr(out))
    return out


def kl_binary_concrete(q_logits, p=0.5, eps=1e-8):
    """:52-77  KL(Bernoulli(sigmoid(q_logits)) || Bernoulli(p)), summed over the last dim, mean over the rest."""
    q = _f32(q_logits, "q_logits")
    L = q.shape[-1]
    out = _out(q)
    _lib.run(_lib.lib().sfv_loss_kl_binary_concrete, q, _lib.ptr(q), q.numel() // L, L, float(p), float(eps), _lib.ptr(out))
    return out


def contrast_loss(x1, x2, label, margin: float = 1.0, dist="euclidean"):
    """:80-107  mean((1 - label) d^2 + label clamp(margin - d, 0)^2), d = pairwise_distance or 1 - cosine_similarity."""
    if dist not in ("euclidean", "cosine"):
        raise ValueError("dist must be 'euclidean' or 'cosine'")
    a, b = _f32(x1, "x1"), _f32(x2, "x2")
    if not (a.dim() == 2 and a.shape == b.shape):
        raise ValueError("contrast_loss expects two [N, D] tensors of one shape")
    lb = _f32(torch.as_tensor(label, device=a.device).expand(a.shape[0]) if not torch.is_tensor(label) else label, "label")
    lb = lb.reshape(-1)
    if lb.numel() != a.shape[0]:
        raise ValueError("label must hold one value per row")
    out = _out(a)
    _lib.run(_lib.lib().sfv_loss_contrast, a, _lib.ptr(a), _lib.ptr(b), _lib.ptr(lb), a.shape[0], a.shape[1], float(margin),
             int(dist == "cosine"), _lib.ptr(out))
    return out
