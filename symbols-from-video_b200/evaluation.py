"""State-consistency evaluation on packed codes and the robustness perturbations, on the device.

Host mirror of scripts/evaluation/state_consistency_eval/embedding_matching.py:
  assign_label                    :195-206
  add_gaussian_noise              :141-161   (+ T.ToPILImage() at :243)
  add_occlusion                   :165-193   (+ T.ToPILImage() at :243)
  calculate_state_consistency     :208-297   (same block in percep_RBVAE_train.py:455-497)

The reference runs one frame at a time (PIL -> tensor -> PIL -> LANCZOS -> SD encoder -> RBVAE ->
``.cpu()``) and ends with ``np.unique(axis=0)`` over float {0,1} vectors.  Here frames stay uint8 on the
device, the perturbation, the LANCZOS resize, both models and the per-state "share of frames equal to
the most common code" all run in libsfv kernels on whole batches; random draws are made on the host from
the same global generators, in the reference's per-frame order, so a seeded run consumes the RNGs
identically.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import _lib, ops
from .autoencoder import SCALE_FACTOR, _scaled_sample


def assign_label(frame_index: int, flags) -> int:
    """embedding_matching.py:195-206: number of flags <= frame_index."""
    label = 0
    for f in flags:
        if frame_index >= f:
            label += 1
        else:
            break
    return label


def labels_from_flags(indices, flags) -> torch.Tensor:
    """Vectorised assign_label for sorted flags -> int32 [N] (host)."""
    idx = np.asarray(indices, dtype=np.int64)
    return torch.from_numpy(np.searchsorted(np.asarray(flags, dtype=np.int64), idx, side="right").astype(np.int32))


def state_consistency(codes: torch.Tensor, labels: torch.Tensor, n_states: int):
    """embedding_matching.py:275-297 on packed codes.

    codes int32/uint32-bits [N, words] cuda, labels int [N] -> (weighted_avg, percentages list, counts list);
    a state with no frames gets 0.0 like the reference (:283-285)."""
    _lib.require_cuda(codes, "codes")
    codes = codes.contiguous()
    labels = labels.to(device=codes.device, dtype=torch.int32).contiguous()
    if codes.dim() != 2 or labels.shape[0] != codes.shape[0]:
        raise ValueError(f"expected codes [N,words] and labels [N], got {tuple(codes.shape)} / {tuple(labels.shape)}")
    best = torch.empty(n_states, dtype=torch.int32, device=codes.device)
    count = torch.empty(n_states, dtype=torch.int32, device=codes.device)
    _lib.run(_lib.lib().sfv_state_consistency, codes, _lib.ptr(codes), _lib.ptr(labels), codes.shape[0], codes.shape[1],
                                                n_states, _lib.ptr(best), _lib.ptr(count))
    _lib.check_async_error(codes.device)     # the .cpu() below synchronises anyway: surface device-side errors here
    best = best.cpu().numpy().astype(np.int64); count = count.cpu().numpy().astype(np.int64)
    pct = [float(b) / float(c) if c > 0 else 0.0 for b, c in zip(best, count)]
    total = int(count.sum())
    weighted = float(np.dot(pct, count) / total) if total > 0 else 0
    return weighted, pct, [int(c) for c in count]


def perturb_frames(frames: torch.Tensor, noise=None, mean=0.0, std=0.1, occ_xy=None, occ_size=0, out=None):
    """uint8 [B,H,W,3] cuda -> uint8 [B,H,W,3]: gaussian noise (noise fp32 [B,3,H,W]) and/or grey squares
    (occ_xy int [B,2] = (x,y) per frame), quantised exactly as the reference's ToTensor/ToPILImage round trip."""
    _lib.require_cuda(frames, "frames")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected uint8 [B,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    B, H, W, _ = frames.shape
    dev = frames.device
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).reshape(B, 3, H, W).contiguous()
    if occ_xy is not None:
        occ_xy = torch.as_tensor(occ_xy).to(device=dev, dtype=torch.int32).reshape(B, 2).contiguous()
        if occ_size > 0 and B > 0:
            lim = torch.tensor([W - occ_size, H - occ_size], device=dev, dtype=torch.int32)
            if bool((occ_xy < 0).any()) or bool((occ_xy > lim).any()):
                raise ValueError("occlusion square outside the frame")
    out = torch.empty_like(frames) if out is None else out
    _lib.run(_lib.lib().sfv_perturb_frames, frames, _lib.ptr(frames), _lib.ptr(out), B, H, W, _lib.ptr(noise), float(mean),
                                             float(std), _lib.ptr(occ_xy), int(occ_size))
    return out


def add_gaussian_noise(frames: torch.Tensor, mean=0.0, std=0.1):
    """embedding_matching.py:141-161 on uint8 frames: per frame one ``torch.randn(1,3,H,W)`` from the
    global CPU generator (what ``randn_like`` of the reference's CPU tensor draws), clamp to [0,1]."""
    B, H, W, _ = frames.shape
    noise = torch.cat([torch.randn(1, 3, H, W) for _ in range(B)]) if B else torch.empty(0, 3, H, W)
    return perturb_frames(frames, noise=noise, mean=mean, std=std)


def occlusion_size(H: int, W: int, coverage: float) -> int:
    return int(np.sqrt(coverage * H * W))                      # :180


def add_occlusion(frames: torch.Tensor, coverage=0.2):
    """embedding_matching.py:165-193: a grey square covering `coverage` of each frame at a position drawn
    per frame with ``random.randint`` (x first, then y), value 0.5."""
    B, H, W, _ = frames.shape
    s = occlusion_size(H, W, coverage)
    xy = [(random.randint(0, W - s), random.randint(0, H - s)) for _ in range(B)]
    return perturb_frames(frames, occ_xy=torch.tensor(xy, dtype=torch.int32).reshape(B, 2), occ_size=s)


@torch.no_grad()
def calculate_state_consistency(model, frames, flags, test_indices, device="cuda", sd_model=None, temperature=0.5,
                                noise_ratio=0.1, perturbation=None, perturbation_params=None,
                                target_size=(1280, 720), batch=16, sample_posterior=True, return_codes=False):
    """embedding_matching.py:208-297 for the perceptual model, batched.

    model        sfv_b200.Seq2SeqBinaryVAE (percep), built for the latent of the resized frame
    frames       uint8 [N,Hs,Ws,3] (host or device): the raw video frames (the reference's ``test_dataset.frames``)
    flags        sorted transition frame indices (transition_flags.txt)
    test_indices frame indices to evaluate, in the order the reference visits them
    sd_model     sfv_b200.FirstStage / AutoencoderKL (the reference's ``sd_model``)
    perturbation None | add_gaussian_noise | add_occlusion (or "gaussian" / "occlusion")
    target_size  load_img_for_sd's (W,H) = (1280,720), each rounded down to a multiple of 32 (:318-338)
    Per frame the host draws, in the reference's order: the perturbation's randoms, ``randn`` for
    posterior.sample() (ddpm.py:542-549 via generate_perceptual_embedding :131-136), ``rand`` for the binary
    concrete (percep_RBVAE_model.py:33).  sample_posterior=False uses mode() instead (deterministic runs).
    Returns (weighted_avg, percentages) -- plus the packed codes and labels if return_codes."""
    vae = getattr(sd_model, "first_stage_model", sd_model)
    if vae is None:
        raise ValueError("sd_model (the KL-f8 first stage) is required for the perceptual model")
    scale = getattr(sd_model, "scale_factor", SCALE_FACTOR)
    if isinstance(frames, np.ndarray):
        frames = torch.from_numpy(frames)
    params = dict(perturbation_params or {})
    kind = {add_gaussian_noise: "gaussian", add_occlusion: "occlusion"}.get(perturbation, perturbation)
    if kind not in (None, "gaussian", "occlusion"):
        raise ValueError(f"unknown perturbation {perturbation!r}")
    W, H = (v - v % 32 for v in target_size)
    lh, lw = H // 8, W // 8
    L = model.latent_dim
    idx = [int(i) for i in test_indices]
    Hs, Ws = frames.shape[1:3]
    occ = occlusion_size(Hs, Ws, params.get("coverage", 0.2)) if kind == "occlusion" else 0
    codes_all = []
    for s in range(0, len(idx), batch):
        sel = idx[s:s + batch]
        n = len(sel)
        noise_img, xy, noise_lat, U = [], [], [], []
        for _ in sel:                                            # the reference's per-frame draw order
            if kind == "gaussian":
                noise_img.append(torch.randn(1, 3, Hs, Ws))
            elif kind == "occlusion":
                xy.append((random.randint(0, Ws - occ), random.randint(0, Hs - occ)))
            if sample_posterior:
                noise_lat.append(torch.randn(1, 4, lh, lw))
            U.append(torch.rand(1, L))
        fr = frames[torch.as_tensor(sel)].to(device, non_blocking=True)
        if kind == "gaussian":
            fr = perturb_frames(fr, noise=torch.cat(noise_img), mean=params.get("mean", 0.0), std=params.get("std", 0.1))
        elif kind == "occlusion":
            fr = perturb_frames(fr, occ_xy=torch.tensor(xy, dtype=torch.int32), occ_size=occ)
        if (Hs, Ws) != (H, W):
            if (W, H) != tuple(target_size):
                # load_img_for_sd resizes twice when the target is not a multiple of 32 (:328-332)
                fr = ops.resize_lanczos(fr, target_size[1], target_size[0])
            fr = ops.resize_lanczos(fr, H, W)
        post = vae.encode_uint8(fr)
        lat = _scaled_sample(post, torch.cat(noise_lat).to(device) if sample_posterior else None, scale)
        codes, _ = model.encode_codes(lat.unsqueeze(1), temperature=temperature, noise_ratio=noise_ratio,
                                      U=torch.cat(U))
        codes_all.append(codes)
    codes = torch.cat(codes_all) if codes_all else torch.empty(0, (L + 31) // 32, dtype=torch.int32, device=device)
    labels = labels_from_flags(idx, flags)
    weighted, pct, _ = state_consistency(codes, labels, len(flags) + 1)
    if return_codes:
        return weighted, pct, codes, labels
    return weighted, pct
