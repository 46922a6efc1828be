"""Oracle: restatement of the RBVAE training losses (forward values).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows models/percep_RBVAE/percep_RBVAE_train.py:27-107 line by
line; the torch.nn.functional calls the reference makes are written out from their published definitions
(``F.mse_loss``, ``F.pairwise_distance``: ||x1 - x2 + eps||_p, ``F.cosine_similarity``:
x1.x2 / (max(||x1||, eps) max(||x2||, eps)), ``F.triplet_margin_loss``: mean(max(d(a,p) - d(a,n) + margin, 0)) with
d(a,n) := min(d(a,n), d(p,n)) under ``swap``).  Pinned against the unmodified functions in tests/test_decoder_losses.py.
"""
from __future__ import annotations

import math

import torch


def l1_loss(q_logits, lamb):
    """:27-29"""
    return lamb * q_logits.abs().sum()


def recon_loss(x_recon, x):
    """:32-33  F.mse_loss, reduction='mean'"""
    return ((x_recon - x) ** 2).mean()


def pairwise_distance(x1, x2, eps=1e-6):
    return ((x1 - x2 + eps) ** 2).sum(-1).sqrt()


def triplet_loss(anchor, pos, neg, margin=1.0, eps=1e-08, swap=True):
    """:35-49  F.triplet_margin_loss(p=2, reduction='mean')"""
    d_ap = pairwise_distance(anchor, pos, eps)
    d_an = pairwise_distance(anchor, neg, eps)
    if swap:
        d_an = torch.minimum(d_an, pairwise_distance(pos, neg, eps))
    return torch.clamp(margin + d_ap - d_an, min=0.0).mean()


def kl_binary_concrete(q_logits, p=0.5, eps=1e-8):
    """:52-77"""
    q = torch.sigmoid(q_logits).clamp(eps, 1.0 - eps)
    log_p, log_1_minus_p = math.log(p), math.log(1.0 - p)
    kl = q * (torch.log(q + eps) - log_p) + (1.0 - q) * (torch.log((1.0 - q) + eps) - log_1_minus_p)
    return kl.sum(dim=-1).mean()


def contrast_loss(x1, x2, label, margin=1.0, dist="euclidean"):
    """:80-107"""
    if dist == "cosine":
        n1 = x1.norm(dim=1).clamp_min(1e-8)
        n2 = x2.norm(dim=1).clamp_min(1e-8)
        d = 1 - (x1 * x2).sum(1) / (n1 * n2)
    else:
        d = pairwise_distance(x1, x2)
    loss = (1 - label) * d ** 2 + label * torch.clamp(margin - d, min=0.0) ** 2
    return loss.mean()
