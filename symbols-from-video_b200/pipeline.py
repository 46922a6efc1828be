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

from . import _lib
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
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
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
    def encode_device(self, frames_dev: torch.Tensor, noise=None, noise_ratio=0.0, U=None, out: EncodeResult | None = None):
        """One batch already resident in HBM: uint8 [B,H,W,3] cuda -> EncodeResult (on device).

        ``out``: preallocated destinations (latents [B,4,h,w] fp32, codes [B,words] int32, h [B,L] fp32), e.g. this
        rank's slice of the all-gather buffers: the posterior-sample and LSTM+threshold+pack kernels write them
        directly, so no staging copy precedes the collective."""
        post = self.vae.encode_uint8(frames_dev)
        lat = _scaled_sample(post, noise, self.scale, out=None if out is None else out.latents)
        if self.rbvae is None:
            return EncodeResult(lat, None, None)
        B = lat.shape[0]
        codes, h = self.rbvae.encode_codes(lat.unsqueeze(1), noise_ratio=noise_ratio, U=U,
                                           out_codes=None if out is None else out.codes,
                                           out_h=None if out is None or out.h is None else out.h.view(B, 1, -1))
        return EncodeResult(lat, codes, h.squeeze(1))

    def _pinned(self, name, shape, dtype):
        t = self._out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._out[name] = t
        return t

    @torch.no_grad()
    def encode_host(self, frames: torch.Tensor | np.ndarray, reuse_output: bool = False,
                    sample_posterior: bool = False, noise: torch.Tensor | None = None,
                    device_out: EncodeResult | None = None):
        """Host uint8 frames [N,H,W,3] (ideally pinned) -> EncodeResult on the HOST.

        Latents are ``scale * posterior.mode()`` by default.  ``sample_posterior=True`` stores what the
        reference's precompute writes, ``scale * posterior.sample()`` (get_percep_embeddings.py:100-101): the
        noise is ``noise`` ([N,4,H/8,W/8]) when given, else one ``torch.randn(1,4,h,w)`` per frame from the global
        CPU generator, in frame order -- the draws DiagonalGaussianDistribution.sample makes (distributions.py:36).
        ``device_out``: device tensors for all N frames that additionally receive the results (written in place by
        the kernels, sub-batch by sub-batch) -- a rank's slice of the gather buffers in the multi-GPU driver.

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

        lh, lw = H // 8, W // 8
        if noise is not None:
            if tuple(noise.shape) != (N, 4, lh, lw):
                raise ValueError(f"noise must be [N,4,H/8,W/8] = {(N, 4, lh, lw)}, got {tuple(noise.shape)}")
            sample_posterior = True

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
            s0, n = starts[i], cur.shape[0]
            nz = None
            if sample_posterior:
                nz = noise[s0:s0 + n] if noise is not None else torch.cat([torch.randn(1, 4, lh, lw) for _ in range(n)])
                nz = nz.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
            dst = None
            if device_out is not None:
                dst = EncodeResult(device_out.latents[s0:s0 + n],
                                   None if device_out.codes is None else device_out.codes[s0:s0 + n],
                                   None if device_out.h is None else device_out.h[s0:s0 + n])
            r = self.encode_device(cur, noise=nz, out=dst)
            consumed[i % 2].record(main)
            lat_out[s0:s0 + n].copy_(r.latents, non_blocking=True)
            if r.codes is not None:
                code_out[s0:s0 + n].copy_(r.codes, non_blocking=True)
                h_out[s0:s0 + n].copy_(r.h, non_blocking=True)
        # synchronises, and turns a tripped pipeline watchdog / fp16 range check into an exception instead of
        # returning garbage with status 0
        _lib.check_async_error(self.device)
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


def all_gather_slices(buf: torch.Tensor, rank: int, world: int, group=None, async_op: bool = False):
    """In-place all-gather of a [world * rows, ...] buffer whose block `rank` this rank has already filled (its
    kernels wrote there directly): no staging copy, one NCCL call.  Returns the work handle when async_op."""
    import torch.distributed as dist
    rows = buf.shape[0] // world
    mine = buf[rank * rows:(rank + 1) * rows]
    if dist.get_backend(group) != "nccl":
        mine = mine.clone()                      # gloo (CPU tests) does not take aliased send / receive buffers
    return dist.all_gather_into_tensor(buf, mine, group=group, async_op=async_op)


def encode_sharded(pipe: FramePipeline, frames, rank: int, world: int, gather=True, group=None,
                   sample_posterior: bool = False):
    """Rank-local encode of this rank's contiguous frame range, then (optionally) an NCCL all-gather of latents,
    packed codes and h so every rank holds the full video.  The gather buffers are allocated up front and this
    rank's kernels write straight into its block of them (no staging copy before the collective);
    ragged tails are padded to the longest shard and stripped after the gather."""
    N = len(frames)
    lo, hi = shard_range(N, rank, world)
    if not gather or world == 1:
        return pipe.encode_host(frames[lo:hi], sample_posterior=sample_posterior), (lo, hi)
    counts = [shard_range(N, r, world)[1] - shard_range(N, r, world)[0] for r in range(world)]
    mx = max(counts)
    dev = pipe.device
    H, W = int(frames.shape[1]), int(frames.shape[2])
    L = pipe.rbvae.latent_dim if pipe.rbvae is not None else 0
    lat_g = torch.zeros(world * mx, 4, H // 8, W // 8, dtype=torch.float32, device=dev)
    codes_g = torch.zeros(world * mx, (L + 31) // 32, dtype=torch.int32, device=dev) if L else None
    h_g = torch.zeros(world * mx, L, dtype=torch.float32, device=dev) if L else None
    n = hi - lo
    mine = EncodeResult(lat_g[rank * mx:rank * mx + n], None if codes_g is None else codes_g[rank * mx:rank * mx + n],
                        None if h_g is None else h_g[rank * mx:rank * mx + n])
    pipe.encode_host(frames[lo:hi], reuse_output=True, sample_posterior=sample_posterior, device_out=mine)
    works = [all_gather_slices(t, rank, world, group, async_op=True) for t in (lat_g, codes_g, h_g) if t is not None]
    for w in works:
        w.wait()

    def strip(t):
        return None if t is None else torch.cat([t[r * mx:r * mx + counts[r]] for r in range(world)])
    return EncodeResult(strip(lat_g), strip(codes_g), strip(h_g)), (lo, hi)
