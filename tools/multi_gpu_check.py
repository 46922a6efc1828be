"""torchrun --nproc-per-node G tools/multi_gpu_check.py : sfv_b200.encode_sharded over NCCL on the chinchess fixture.
Every rank encodes its contiguous frame range, latents / codes / h are all-gathered (ragged), and every rank checks
the gathered codes against the reference's golden codes (tests/golden/chinchess_480x64x128.npz)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sfv_b200
from oracle import chinchess                    # checker side only (fixture decoding + weights)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "chinchess_480x64x128.npz"))
    u8 = chinchess.frames_from_delta(g["frame_delta"])[:477]          # 477 frames: ragged over 2 / 4 / 8 ranks
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    vae = sfv_b200.AutoencoderKL(precision=prec)
    vae.load_state_dict(sfv_b200.init_encoder_state_dict(int(g["weight_seed"])))
    rsd, _ = chinchess.rbvae_weights()
    H, W = chinchess.HW
    rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, chinchess.L, chinchess.L, input_hw=(H // 8, W // 8), precision=prec)
    rb.load_state_dict(rsd)
    pipe = sfv_b200.FramePipeline(vae, rb, batch=64, device=torch.device("cuda", local))
    res, (lo, hi) = sfv_b200.encode_sharded(pipe, torch.from_numpy(u8), rank, world)
    vae.check_async_error()
    z = sfv_b200.unpack_codes(res.codes, chinchess.L).cpu().numpy()
    ref, h = g["z_hard"][:477], g["h"][:477]
    diff = z != ref
    band = np.abs(h) < 1e-3
    out = dict(rank=rank, world=world, shard=(lo, hi), gathered=int(z.shape[0]), flips_outside=int((diff & ~band).sum()),
               flips_inside=int((diff & band).sum()), h_maxabs=float(np.abs(res.h.cpu().numpy() - h).max()),
               latent_sum_maxabs=float(np.abs(res.latents.double().sum(dim=(1, 2, 3)).cpu().numpy() - g["latent_sum"][:477]).max()))
    print(out, flush=True)
    assert z.shape[0] == 477 and out["flips_outside"] == 0
    if prec == "fp32":
        assert out["h_maxabs"] < 1e-5
    # BASELINE configs[2]: the full-video precompute driver, frame range sharded over the ranks, latents + codes
    # all-gathered in place, rank 0 writes the reference-format .npy (needs the sample video under oracle/_ref)
    from oracle import ref_shim
    video = ref_shim.video_path()
    if video is not None:
        import tempfile
        tmp = tempfile.mkdtemp() if rank == 0 else None
        box = [tmp]
        dist.broadcast_object_list(box, src=0)
        npy = os.path.join(box[0], "chinchess_perceps.npy")
        r = sfv_b200.precompute_embeddings(sfv_b200.VideoSource(video), vae, rb, target_size=(W, 72), batch=64, rank=rank,
                                           world=world, sample_posterior=False, fit="crop", n_decoders=2,
                                           out_npy=npy if rank == 0 else None)
        z2 = sfv_b200.unpack_codes(r.codes, chinchess.L).numpy()
        d2 = z2 != g["z_hard"]
        b2 = np.abs(g["h"]) < 1e-3
        out2 = dict(rank=rank, precompute_frames=len(r.keys), flips_outside=int((d2 & ~b2).sum()), flips_inside=int((d2 & b2).sum()),
                    stats={k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.stats.items() if "fps" in k})
        print(out2, flush=True)
        assert len(r.keys) == 480 and out2["flips_outside"] == 0
        dist.barrier()
        if rank == 0:
            emb = np.load(npy, allow_pickle=True).item()
            assert len(emb) == 480 and emb["0000000479.jpg"].shape == (1, 4, H // 8, W // 8) and emb["0000000000.jpg"].dtype == np.float32
            print("precompute .npy written by rank 0:", len(emb), "keys", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
