"""Batched, sharded frame -> KL-f8 latent -> binary-code driver.

Replaces the reference's per-frame loops (batch 1, one H2D and one D2H sync per
frame): ``get_percep_embeddings.main`` (src/stable-diffusion/get_percep_embeddings.py:76-114)
and ``calculate_state_consistency`` (scripts/evaluation/state_consistency_eval/
embedding_matching.py:235-265).  The latent stays on the device between the two
models; only uint8 frames go up and (latents, packed codes) come down.

Multi-GPU: frames are independent at T=1, so rank r of G encodes the
contiguous range [floor(r N / G), floor((r+1) N / G)) with no data-path
collective; one all_gather of the packed codes and latents at the end.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from .autoencoder import SCALE_FACTOR, AutoencoderKL, _scaled_sample
from .rbvae import Seq2SeqBinaryVAE


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous frame range of `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


@dataclass
class EncodeResult:
    latents: torch.Tensor      # fp32 [N,4,h,w] = scale_factor * (mode or sample)
    codes: torch.Tensor        # int32 (uint32 bits) [N, ceil(L/32)]
    h: torch.Tensor            # fp32 [N,L] LSTM hidden state that was thresholded


class FramePipeline:
    """uint8 frames [N,H,W,3] -> (latents, packed codes)."""

    def __init__(self, vae: AutoencoderKL, rbvae: Seq2SeqBinaryVAE | None, batch: int = 64,
                 scale_factor: float = SCALE_FACTOR, device="cuda"):
        self.vae, self.rbvae, self.batch, self.scale = vae, rbvae, batch, scale_factor
        self.device = torch.device(device)
        self._stage = [None, None]
        self._copy_stream = None
        self._out = {}
        self.host_batch = int(os.environ.get("SFV_HOST_BATCH", getattr(vae, "chunk", None) or 16))

    def _staging(self, i, shape):
        t = self._stage[i]
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.uint8, device=self.device)
            self._stage[i] = t
        return t

    @torch.no_grad()
    def encode_device(self, frames_dev: torch.Tensor, noise=None, noise_ratio=0.0, U=None):
        """One batch already resident in HBM: uint8 [B,H,W,3] cuda -> EncodeResult (on device)."""
        post = self.vae.encode_uint8(frames_dev)
        lat = _scaled_sample(post, noise, self.scale)
        if self.rbvae is None:
            return EncodeResult(lat, None, None)
        codes, h = self.rbvae.encode_codes(lat.unsqueeze(1), noise_ratio=noise_ratio, U=U)
        return EncodeResult(lat, codes, h.squeeze(1))

    def _pinned(self, name, shape, dtype):
        t = self._out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._out[name] = t
        return t

    @torch.no_grad()
    def encode_host(self, frames: torch.Tensor | np.ndarray, reuse_output: bool = False):
        """Host uint8 frames [N,H,W,3] (ideally pinned) -> EncodeResult on the HOST.

        The frames are walked in sub-batches of ``host_batch`` (default: the encoder's chunk): the H2D copy of
        sub-batch i+1 runs on a copy stream under the kernels of sub-batch i, and each sub-batch's results go down
        into pinned host buffers asynchronously, so only the first upload and the last download are exposed.
        ``reuse_output=True`` returns views of the pipeline's pinned buffers (overwritten by the next call)."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
        N, H, W, _ = frames.shape
        hb = max(1, min(self.batch, self.host_batch))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        starts = list(range(0, N, hb))
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        L = self.rbvae.latent_dim if self.rbvae is not None else 0
        lat_out = self._pinned("lat", (N, 4, H // 8, W // 8), torch.float32)
        code_out = self._pinned("codes", (N, (L + 31) // 32), torch.int32) if self.rbvae is not None else None
        h_out = self._pinned("h", (N, L), torch.float32) if self.rbvae is not None else None

        def upload(i):
            s = starts[i]
            chunk = frames[s:s + hb]
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(consumed[i % 2])
                dst = self._staging(i % 2, (hb, H, W, 3))[:chunk.shape[0]]
                dst.copy_(chunk, non_blocking=True)
                ready[i % 2].record(self._copy_stream)
            return dst

        pending = upload(0) if starts else None
        for i in range(len(starts)):
            cur = pending
            main.wait_event(ready[i % 2])
            if i + 1 < len(starts):
                pending = upload(i + 1)
            r = self.encode_device(cur)
            consumed[i % 2].record(main)
            s0, n = starts[i], cur.shape[0]
            lat_out[s0:s0 + n].copy_(r.latents, non_blocking=True)
            if r.codes is not None:
                code_out[s0:s0 + n].copy_(r.codes, non_blocking=True)
                h_out[s0:s0 + n].copy_(r.h, non_blocking=True)
        torch.cuda.synchronize(self.device)
        if reuse_output:
            return EncodeResult(lat_out, code_out, h_out)
        return EncodeResult(lat_out.clone(), None if code_out is None else code_out.clone(),
                            None if h_out is None else h_out.clone())


def all_gather_ragged(local: torch.Tensor, counts: list[int], group=None):
    """All-gather per-rank row blocks of different lengths (contiguous frame ranges):
    pad to the longest block, one all_gather_into_tensor, strip the padding."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)])


def encode_sharded(pipe: FramePipeline, frames, rank: int, world: int, gather=True, group=None):
    """Rank-local encode of this rank's contiguous frame range, then (optionally)
    NCCL all-gather of latents and packed codes so every rank holds the full video."""
    N = len(frames)
    lo, hi = shard_range(N, rank, world)
    res = pipe.encode_host(frames[lo:hi])
    if not gather or world == 1:
        return res, (lo, hi)
    counts = [shard_range(N, r, world)[1] - shard_range(N, r, world)[0] for r in range(world)]
    dev = pipe.device
    lat = all_gather_ragged(res.latents.to(dev), counts, group)
    codes = all_gather_ragged(res.codes.to(dev), counts, group) if res.codes is not None else None
    h = all_gather_ragged(res.h.to(dev), counts, group) if res.h is not None else None
    return EncodeResult(lat, codes, h), (lo, hi)
