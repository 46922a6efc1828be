"""Full-video embedding precompute (BASELINE configs[2]): one command turns a video into the reference's
``*_perceps.npy`` (and, optionally, packed binary codes).

Mirror of ``get_percep_embeddings.main`` (src/stable-diffusion/get_percep_embeddings.py:76-114):

  reference, per frame (batch 1)                     here, per batch
  ------------------------------------------------   -----------------------------------------------------------
  JPEG on disk -> Image.open -> convert("RGB")  :54  cv2 decode threads -> pinned ring -> one async H2D per slot
  resize LANCZOS to 1280x720                    :60  sfv_resize_normalise (PIL's 8-bit two-pass arithmetic, bit exact)
  resize again to multiples of 32 (1280x704)    :63  second sfv_resize_normalise pass, only if the size changes
  /255, HWC->NCHW, 2x-1                         :67  fused into conv_in's operand build
  encode_first_stage                            :99  AutoencoderKL.encode_uint8
  get_first_stage_encoding = 0.18215 * sample() :100 sfv_posterior_sample, noise drawn per frame from the global
                                                     CPU generator in frame order, as sample() does (distributions.py:36)
  embeddings[basename] = latent.cpu().numpy()   :106 same keys, float32 (1,4,h,w)
  np.save(OUTPUT, embeddings)                   :113 save_embeddings_npy (pickled dict) and/or FlatEmbeddingStore
  per-frame exceptions are printed and skipped  :107 decode errors abort the job (a silently shorter file is worse)

Multi-GPU: rank r of G takes the contiguous frame range [floor(rN/G), floor((r+1)N/G)); every rank
decodes only its range; results are all-gathered in place (NCCL) so rank 0 (or every rank) can write the file.
Resumable: with ``part_dir`` every finished block of ``part_frames`` frames is written as
``part-<lo>-<hi>.npz``; a restarted job skips blocks whose part file exists and assembles the final file from
the parts.  With ``sample_posterior=True`` a resumed job draws different noise than an uninterrupted one for
the remaining frames unless ``noise_seed`` is given (per-frame generators seeded ``noise_seed + frame index``).
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, ops
from .autoencoder import SCALE_FACTOR, _scaled_sample
from .embedding_store import FlatEmbeddingStore, save_embeddings_npy
from .feeder import Feeder
from .pipeline import all_gather_slices, shard_range


@dataclass
class PrecomputeResult:
    keys: list
    latents: torch.Tensor            # fp32 [N,4,h,w] (host), N = frames of the whole job if gathered, else this rank's
    codes: torch.Tensor | None       # int32 [N, ceil(L/32)] (host)
    h: torch.Tensor | None           # fp32 [N, L] (host)
    frame_range: tuple
    stats: dict = field(default_factory=dict)


def target_hw(target_size=(1280, 720)):
    """load_img: (w, h) -> each rounded down to a multiple of 32 (get_percep_embeddings.py:63-66)."""
    w, h = target_size
    return h - h % 32, w - w % 32


@torch.no_grad()
def precompute_embeddings(source, vae, rbvae=None, target_size=(1280, 720), batch: int = 32, rank: int = 0,
                          world: int = 1, gather: bool = True, group=None, sample_posterior: bool = True,
                          noise_seed: int | None = None, scale_factor: float = SCALE_FACTOR, fit: str = "resize",
                          n_decoders: int = 2, slots: int = 4, out_npy: str | None = None,
                          out_flat: str | None = None, out_codes: str | None = None, part_dir: str | None = None,
                          part_frames: int = 256, frame_range: tuple | None = None, device=None,
                          rbvae_temperature: float = 0.5) -> PrecomputeResult:
    """See the module docstring.  source: feeder.VideoSource / FrameDirSource / ArraySource.

    fit          "resize": load_img's two LANCZOS passes (1280x720 then 1280x704); "crop": one pass to target_size,
                 then the top-left (h - h%32, w - w%32) window (how the chinchess test fixture was cut)
    frame_range  (lo, hi) sub-range of the source to process (default: everything), sharded over ranks
    out_npy      reference-format pickled dict (written by rank 0 after the gather, or by every rank for its own
                 range when gather=False and world > 1: '<out_npy>.rank<r>')
    """
    if fit not in ("resize", "crop"):
        raise ValueError("fit must be 'resize' or 'crop'")
    if sample_posterior and noise_seed is None:
        n_decoders = 1          # global-RNG draws must follow frame order (as the reference's loop does): in-order decode
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    N0, N1 = (0, len(source)) if frame_range is None else frame_range
    n_job = N1 - N0
    lo, hi = shard_range(n_job, rank, world)
    lo, hi = N0 + lo, N0 + hi
    H, W = target_hw(target_size)
    lh, lw = H // 8, W // 8
    L = rbvae.latent_dim if rbvae is not None else 0
    words = (L + 31) // 32
    counts = [shard_range(n_job, r, world)[1] - shard_range(n_job, r, world)[0] for r in range(world)]
    mx = max(counts) if counts else 0
    do_gather = gather and world > 1
    rows = world * mx if do_gather else hi - lo
    base = rank * mx if do_gather else 0
    # results live on the device (this rank's block of the gather buffers), written in place by the kernels
    lat_g = torch.zeros(rows, 4, lh, lw, dtype=torch.float32, device=dev)
    codes_g = torch.zeros(rows, words, dtype=torch.int32, device=dev) if L else None
    h_g = torch.zeros(rows, L, dtype=torch.float32, device=dev) if L else None

    # ---- resume: blocks of part_frames frames inside this rank's range --------------------------------------------
    blocks = [(a, min(a + part_frames, hi)) for a in range(lo, hi, part_frames)] if part_dir else [(lo, hi)]
    done_blocks = 0
    if part_dir:
        os.makedirs(part_dir, exist_ok=True)

    def part_path(a, b):
        return os.path.join(part_dir, f"part-{a:010d}-{b:010d}.npz")

    Hs, Ws = source.frame_hw
    dummy_u = torch.zeros(batch, max(L, 1), device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    t_start = time.perf_counter()
    gpu_ms = 0.0
    ev_pairs = []
    decode_stats = []
    n_done = 0
    for (a, b) in blocks:
        if part_dir and os.path.exists(part_path(a, b)):
            z = np.load(part_path(a, b))
            lat_g[base + a - lo:base + b - lo].copy_(torch.from_numpy(z["latents"]))
            if L:
                codes_g[base + a - lo:base + b - lo].copy_(torch.from_numpy(z["codes"]))
                h_g[base + a - lo:base + b - lo].copy_(torch.from_numpy(z["h"]))
            done_blocks += 1
            continue
        feeder = Feeder(source, a, b, batch=batch, slots=slots, n_decoders=n_decoders)
        staging = [torch.empty((batch, Hs, Ws, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        staged = [torch.cuda.Event() for _ in range(2)]
        consumed = [None, None]
        k = 0
        try:
            for slot in feeder:
                n, first = slot.n, slot.first
                st = staging[k % 2]
                with torch.cuda.stream(copy_stream):
                    if consumed[k % 2] is not None:
                        copy_stream.wait_event(consumed[k % 2])
                    st[:n].copy_(slot.buf[:n], non_blocking=True)
                    staged[k % 2].record(copy_stream)
                # the slot may be refilled once the copy has left it: hand it back from a callback-free wait
                staged[k % 2].synchronize()
                feeder.release(slot)
                main.wait_event(staged[k % 2])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
                fr = st[:n]
                if (Hs, Ws) != (H, W):
                    tw, th = target_size
                    if (Hs, Ws) != (th, tw):
                        fr = ops.resize_lanczos(fr, th, tw)                  # load_img :60
                    if (th, tw) != (H, W):
                        fr = ops.resize_lanczos(fr, H, W) if fit == "resize" else fr[:, :H, :W].contiguous()   # :63-66
                post = vae.encode_uint8(fr)
                nz = None
                if sample_posterior:
                    if noise_seed is None:
                        nz = torch.cat([torch.randn(1, 4, lh, lw) for _ in range(n)])
                    else:
                        nz = torch.cat([torch.randn(1, 4, lh, lw, generator=torch.Generator().manual_seed(noise_seed + first + i))
                                        for i in range(n)])
                    nz = nz.to(dev, non_blocking=True)
                r0 = base + first - lo
                lat = _scaled_sample(post, nz, scale_factor, out=lat_g[r0:r0 + n])
                if rbvae is not None:
                    # U given (unused at noise_ratio 0) so the call draws nothing from the global RNG: the only
                    # consumer of the generator in this loop is the posterior noise, as in the reference's loop
                    rbvae.encode_codes(lat.unsqueeze(1), temperature=rbvae_temperature, noise_ratio=0.0, U=dummy_u[:n],
                                       out_codes=codes_g[r0:r0 + n], out_h=h_g[r0:r0 + n].view(n, 1, L))
                e1.record(main)
                ev_pairs.append((e0, e1))
                consumed[k % 2] = torch.cuda.Event()
                consumed[k % 2].record(main)
                k += 1
                n_done += n
        finally:
            feeder.close()
        _lib.check_async_error(dev)         # synchronises; a tripped watchdog / range check must not reach the file
        decode_stats.append(feeder.stats())
        if part_dir:
            sl = slice(base + a - lo, base + b - lo)
            tmp = part_path(a, b) + ".tmp.npz"
            np.savez(tmp, latents=lat_g[sl].cpu().numpy(), codes=codes_g[sl].cpu().numpy() if L else np.zeros((0,)),
                     h=h_g[sl].cpu().numpy() if L else np.zeros((0,)))
            os.replace(tmp, part_path(a, b))
    _lib.check_async_error(dev)
    for e0, e1 in ev_pairs:
        gpu_ms += e0.elapsed_time(e1)
    wall = time.perf_counter() - t_start

    if do_gather:
        works = [all_gather_slices(t, rank, world, group, async_op=True) for t in (lat_g, codes_g, h_g) if t is not None]
        for w in works:
            w.wait()

        def strip(t):
            return None if t is None else torch.cat([t[r * mx:r * mx + counts[r]] for r in range(world)]).cpu()
        lat, codes, h = strip(lat_g), strip(codes_g), strip(h_g)
        key_range = (N0, N1)
    else:
        lat = lat_g.cpu()
        codes = None if codes_g is None else codes_g.cpu()
        h = None if h_g is None else h_g.cpu()
        key_range = (lo, hi)
    keys = [source.key(i) for i in range(*key_range)]

    dec_frames = sum(s["frames"] for s in decode_stats)
    dec_busy = sum(s["decode_busy_s"] for s in decode_stats)
    stats = dict(frames_this_rank=hi - lo, frames_encoded_now=n_done, resumed_blocks=done_blocks, wall_s=wall,
                 frames_per_s=(n_done / wall) if wall > 0 and n_done else None,
                 gpu_busy_s=gpu_ms / 1e3, gpu_bound_fps=(n_done / (gpu_ms / 1e3)) if gpu_ms > 0 else None,
                 decode_bound_fps=(dec_frames / dec_busy) if dec_busy > 0 else None,
                 decoders=n_decoders, source_hw=(Hs, Ws), encoded_hw=(H, W), rank=rank, world=world)

    writer = rank == 0 or not do_gather
    suffix = f".rank{rank}" if (world > 1 and not do_gather) else ""
    if writer and out_npy:
        save_embeddings_npy(out_npy + suffix if suffix else out_npy, keys, lat)
    if writer and out_flat:
        FlatEmbeddingStore(keys, lat.numpy()).save(out_flat + suffix if suffix else out_flat)
    if writer and out_codes and codes is not None:
        np.savez(out_codes + suffix if suffix else out_codes, keys=np.array(keys), codes=codes.numpy(), h=h.numpy(),
                 latent_dim=L)
    return PrecomputeResult(keys, lat, codes, h, key_range, stats)


def main(argv=None):
    """python -m sfv_precompute <video or frame folder> --ckpt sd-v1-4.ckpt --out chin_chess_perceps.npy"""
    import argparse
    import torch.distributed as dist
    from .autoencoder import AutoencoderKL
    from .feeder import FrameDirSource, VideoSource
    from .rbvae import Seq2SeqBinaryVAE
    from .weights import init_encoder_state_dict
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("input", help="video file, or a folder of extracted frames (the reference's IMAGE_FOLDER)")
    ap.add_argument("--out", required=True, help="reference-format embeddings file (*_perceps.npy)")
    ap.add_argument("--flat", default=None, help="also write the flat mmap-able store here")
    ap.add_argument("--codes", default=None, help="also write packed binary codes (needs --rbvae-ckpt)")
    ap.add_argument("--ckpt", default=None, help="Stable Diffusion / kl-f8 checkpoint; omit for seeded random weights")
    ap.add_argument("--rbvae-ckpt", default=None, help="best_model_*.pt of a percep RBVAE (['model_state_dict'])")
    ap.add_argument("--latent-dim", type=int, default=25)
    ap.add_argument("--size", default="1280x720", help="load_img's target size WxH")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--decoders", type=int, default=2)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--mode", action="store_true", help="store scale*mode() instead of scale*sample()")
    ap.add_argument("--seed", type=int, default=None, help="per-frame noise generators seeded seed + frame index")
    ap.add_argument("--parts", default=None, help="directory for resumable part files")
    a = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vae = AutoencoderKL(ckpt_path=a.ckpt, precision=a.precision)
    if a.ckpt is None:
        vae.load_state_dict(init_encoder_state_dict(0))
    tw, th = (int(v) for v in a.size.lower().split("x"))
    rb = None
    if a.rbvae_ckpt:
        H, W = target_hw((tw, th))
        rb = Seq2SeqBinaryVAE(4, 4, a.latent_dim, a.latent_dim, input_hw=(H // 8, W // 8))
        rb.load_state_dict(torch.load(a.rbvae_ckpt, map_location="cpu")["model_state_dict"], strict=False)
    src = FrameDirSource(a.input) if os.path.isdir(a.input) else VideoSource(a.input)
    res = precompute_embeddings(src, vae, rb, target_size=(tw, th), batch=a.batch, rank=rank, world=world,
                                sample_posterior=not a.mode, noise_seed=a.seed, n_decoders=a.decoders,
                                out_npy=a.out, out_flat=a.flat, out_codes=a.codes, part_dir=a.parts)
    if rank == 0:
        print(f"\nSaved embeddings for {len(res.keys)} images to {a.out}")
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.stats.items()})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
