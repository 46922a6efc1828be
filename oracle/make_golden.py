"""Generate tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
The reference has no tests or fixtures of its own for this path, so
these vectors -- outputs of the reference's own classes on seeded inputs and
seeded random-init weights (no checkpoint exists offline) -- are what
pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
Weights are regenerated from the seed by oracle.kl_f8.init_state_dict /
oracle.rbvae.init_state_dict, so only inputs' seeds and outputs are stored.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import chinchess, frames, kl_f8, rbvae, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def encoder_case(name, seed, shape, frame_seed, smooth):
    sd = kl_f8.init_state_dict(seed)
    m = ref_shim.autoencoder_kl(sd)
    B, H, W = shape
    u8 = frames.synthetic_frames(B, H, W, seed=frame_seed, smooth=smooth)
    x = frames.normalise_u8(u8)
    with torch.no_grad():
        post = m.encode(x)
        enc_out = m.encoder(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), weight_seed=seed, frame_seed=frame_seed,
                        smooth=smooth, shape=np.array(shape), mean=post.mean.numpy(),
                        logvar=post.logvar.numpy(), std=post.std.numpy(), var=post.var.numpy(),
                        parameters=post.parameters.numpy(), encoder_out_sum=float(enc_out.double().sum()))
    print(name, "mean std", float(post.mean.std()), "logvar range", float(post.logvar.min()), float(post.logvar.max()))


def rbvae_case(name, kind, cin, ch, layers, L, hw, seed, T):
    fh = hw[0]
    for _ in range(3):
        fh = (fh - 1) // 2 + 1
    fw = hw[1]
    for _ in range(3):
        fw = (fw - 1) // 2 + 1
    sd = rbvae.init_state_dict(cin, L, (fh, fw), channels=ch, num_layers=layers, seed=seed)
    m = ref_shim.rbvae(kind, cin, L, sd, feat_hw=(fh, fw))
    g = torch.Generator().manual_seed(seed + 100)
    B = 3
    x = torch.randn(B, T, cin, *hw, generator=g) * (0.18215 * 4 if kind == "percep" else 1.0)
    if kind != "percep":
        x = torch.rand(B, T, cin, *hw, generator=g)
    with torch.no_grad():
        z0 = m.encode(x, temperature=0.5, hard=True, noise_ratio=0.0)
        # h_seq: the reference only exposes it through forward(), whose decoder half is hard-wired to the
        # native shape; recompute it from the reference's own sub-modules exactly as encode() does (:176-186)
        Bq, Tq = x.shape[:2]
        logits = m.encoder_cnn(x.reshape(Bq * Tq, cin, *hw)).reshape(Bq, Tq, L)
        h_seq, _ = m.encoder_rnn(logits)
        torch.manual_seed(777)
        z_noise = m.encode(x, temperature=0.5, hard=True, noise_ratio=0.3)
        torch.manual_seed(777)
        U = torch.rand(Bq * Tq, L)       # the draw binary_concrete_logits made (percep_RBVAE_model.py:33)
        torch.manual_seed(778)
        z_soft = m.encode(x, temperature=0.7, hard=False, noise_ratio=0.1)
        torch.manual_seed(778)
        U_soft = torch.rand(Bq * Tq, L)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, cin=cin, ch=ch, layers=layers, L=L,
                        hw=np.array(hw), seed=seed, T=T, x=x.numpy(), logits=logits.numpy(), h=h_seq.numpy(),
                        z_hard=z0.numpy(), z_noise=z_noise.numpy(), U=U.numpy(), z_soft=z_soft.numpy(),
                        U_soft=U_soft.numpy())
    print(name, "min|h|", float(h_seq.abs().min()), "max|h|", float(h_seq.abs().max()))


def rbvae_forward_case(name, kind, cin, ch, layers, L, hw, seed, B, T):
    """Full Seq2SeqBinaryVAE.forward of the UNMODIFIED reference at its native shape (the decoder's reshape is
    hard-wired to it): x_recon, h_seq, z_seq and the decoder LSTM output, soft and hard."""
    fh, fw = hw
    for _ in range(3):
        fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
    sd = rbvae.init_state_dict(cin, L, (fh, fw), channels=ch, num_layers=layers, seed=seed)
    sd.update(rbvae.init_decoder_state_dict(cin, L, (fh, fw), channels=ch, num_layers=layers, seed=seed))
    m = ref_shim.rbvae(kind, cin, L, sd)
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.randn(B, T, cin, *hw, generator=g) * 0.7 if kind == "percep" else torch.rand(B, T, cin, *hw, generator=g)
    out = {}
    with torch.no_grad():
        for tag, hard, temp, nr, s0 in (("soft", False, 1.0, 0.1, 901), ("hard", True, 0.5, 0.0, 902)):
            torch.manual_seed(s0)
            x_recon, h_seq, z_seq = m(x, temperature=temp, hard=hard, noise_ratio=nr)
            torch.manual_seed(s0)
            U = torch.rand(B * T, L)          # the draw binary_concrete_logits made (percep_RBVAE_model.py:33)
            d_seq, _ = m.decoder_rnn(z_seq)
            out.update({f"x_recon_{tag}": x_recon.numpy(), f"h_{tag}": h_seq.numpy(), f"z_{tag}": z_seq.numpy(),
                        f"d_{tag}": d_seq.numpy(), f"U_{tag}": U.numpy()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, cin=cin, ch=ch, layers=layers, L=L, hw=np.array(hw),
                        seed=seed, B=B, T=T, x=x.numpy(), **out)
    print(name, "x_recon range", float(out["x_recon_soft"].min()), float(out["x_recon_soft"].max()))


def losses_case():
    """The reference's own loss functions (percep_RBVAE_train.py:27-107, cut out unmodified) on seeded inputs."""
    fn = ref_shim.train_loss_functions()
    g = torch.Generator().manual_seed(4242)
    N, L, D = 12, 25, 25
    q = torch.randn(N, L, generator=g) * 3
    q[0, :3] = torch.tensor([40.0, -40.0, 0.0])          # saturated sigmoid: the clamp / eps branches
    a, p, n = (torch.randn(N, D, generator=g) for _ in range(3))
    label = (torch.rand(N, generator=g) > 0.5).float()
    xr, x = torch.rand(2, 3, 4, 8, 16, generator=g), torch.rand(2, 3, 4, 8, 16, generator=g)
    res = dict(l1=fn["l1_loss"](q, 0.01), mse=fn["recon_loss"](xr, x),
               triplet_swap=fn["triplet_loss"](a, p, n), triplet_noswap=fn["triplet_loss"](a, p, n, margin=0.5, swap=False),
               kl_half=fn["kl_binary_concrete"](q), kl_p03=fn["kl_binary_concrete"](q, p=0.3),
               contrast_euclid=fn["contrast_loss"](a, p, label), contrast_cos=fn["contrast_loss"](a, p, label, margin=0.7, dist="cosine"))
    np.savez_compressed(os.path.join(OUT, "losses.npz"), q=q.numpy(), a=a.numpy(), p=p.numpy(), n=n.numpy(), label=label.numpy(),
                        xr=xr.numpy(), x=x.numpy(), **{k: np.float32(v.item()) for k, v in res.items()})
    print("losses", {k: float(v) for k, v in res.items()})


def resize_case():
    """PIL LANCZOS on the first chinchess frame (768x432 -> 1280x720 -> 1280x704), subsampled."""
    import cv2
    from PIL import Image
    cap = cv2.VideoCapture(os.path.join(ref_shim.REF_ROOT, "videos/chinchess_gettyimages-148739276-640_adpp.mp4"))
    ok, fr = cap.read()
    fr = cv2.cvtColor(fr, cv2.COLOR_BGR2RGB)
    _, u8 = frames.load_img_pil(Image.fromarray(fr))
    rng = np.random.default_rng(5)
    small = rng.integers(0, 256, (45, 80, 3), dtype=np.uint8)
    small_up = np.array(Image.fromarray(small).resize((128, 72), resample=Image.LANCZOS))
    small_dn = np.array(Image.fromarray(small).resize((32, 24), resample=Image.LANCZOS))
    np.savez_compressed(os.path.join(OUT, "resize_pil.npz"), frame0=fr, frame0_1280x704_rows=u8[::16],
                        frame0_checksum=int(u8.astype(np.int64).sum()), small=small, small_up=small_up,
                        small_dn=small_dn)
    print("resize_pil checksum", int(u8.astype(np.int64).sum()))


def chinchess_case():
    """The reference's sample video (480 frames; videos/frames/transition_flags.txt: chinese_chess) through
    the UNMODIFIED reference classes: load_img's LANCZOS resize + crop (get_percep_embeddings.py:48-71, at a
    128x72 target so the fixture stays small), AutoencoderKL.encode -> 0.18215*mode()
    (embedding_matching.py:131-136), percep Seq2SeqBinaryVAE.encode(hard=True, noise_ratio=0).
    Frames are stored as wrap-around deltas between consecutive frames (deflate-friendly)."""
    import cv2
    from PIL import Image
    H, W = chinchess.HW
    cap = cv2.VideoCapture(os.path.join(ref_shim.REF_ROOT, "videos/chinchess_gettyimages-148739276-640_adpp.mp4"))
    fr = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        img = Image.fromarray(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)).convert("RGB")
        img = img.resize((W, 72), resample=Image.LANCZOS).crop((0, 0, W, H))
        fr.append(np.array(img))
    u8 = np.stack(fr)
    assert u8.shape == (480, H, W, 3), u8.shape
    L = chinchess.L
    sd = kl_f8.init_state_dict(0)
    m = ref_shim.autoencoder_kl(sd)
    rsd, (fh, fw) = chinchess.rbvae_weights()
    rb = ref_shim.rbvae("percep", 4, L, rsd, feat_hw=(fh, fw))
    lat, hs, zs = [], [], []
    with torch.no_grad():
        for i in range(0, 480, 32):
            x = frames.normalise_u8(u8[i:i + 32])
            z = 0.18215 * m.encode(x).mode()
            lat.append(z)
            xi = z.unsqueeze(1)
            zs.append(rb.encode(xi, temperature=0.5, hard=True, noise_ratio=0.0)[:, 0])
            logits = rb.encoder_cnn(z).reshape(-1, 1, L)
            hs.append(rb.encoder_rnn(logits)[0][:, 0])
    lat = torch.cat(lat).numpy(); hs = torch.cat(hs).numpy(); zs = torch.cat(zs).numpy()
    delta = u8.copy()
    delta[1:] = u8[1:] - u8[:-1]                       # uint8 wrap-around
    np.savez_compressed(os.path.join(OUT, "chinchess_480x64x128.npz"), frame_delta=delta, weight_seed=0,
                        L=L, latent_sum=lat.astype(np.float64).sum(axis=(1, 2, 3)),
                        latent_first=lat[:2], latent_last=lat[-1:], h=hs, z_hard=zs,
                        transitions=np.array([74, 206, 282, 389]), grey_out=10)
    print("chinchess: min|h|", float(np.abs(hs).min()), "in-band", int((np.abs(hs) < 1e-3).sum()),
          "distinct codes", len(np.unique(zs, axis=0)))


def evaluation_case():
    """The reference's own add_gaussian_noise / add_occlusion / calculate_state_consistency
    (embedding_matching.py:141-297, executed unmodified through ref_shim.embedding_matching_functions)
    on small seeded inputs."""
    import random
    import torchvision.transforms as T
    from PIL import Image
    rng = np.random.default_rng(21)
    imgs = rng.integers(0, 256, (3, 24, 40, 3), dtype=np.uint8)
    flags = [50, 120, 200]
    ns = ref_shim.embedding_matching_functions(flags=flags)
    torch.manual_seed(5)
    gauss = np.stack([np.array(T.ToPILImage()(ns["add_gaussian_noise"](T.ToTensor()(Image.fromarray(im)), mean=0.05, std=0.2)))
                      for im in imgs])
    torch.manual_seed(5)
    noise = torch.cat([torch.randn(1, 3, 24, 40) for _ in imgs]).numpy()
    random.seed(7)
    occ = np.stack([np.array(T.ToPILImage()(ns["add_occlusion"](T.ToTensor()(Image.fromarray(im)), coverage=0.3)))
                    for im in imgs])
    random.seed(7)
    size = int(np.sqrt(0.3 * 24 * 40))
    xy = np.array([(random.randint(0, 40 - size), random.randint(0, 24 - size)) for _ in imgs], dtype=np.int32)
    # state consistency: 300 frames, 70-bit codes drawn from a few prototypes per state with bit noise
    L = 70
    protos = rng.integers(0, 2, (8, L)).astype(np.float32)
    which = rng.integers(0, 8, 300)
    z = protos[which].copy()
    flip = rng.random((300, L)) < 0.004
    z[flip] = 1 - z[flip]

    class DS:
        frames = [torch.full((1, 1, 1), float(i)) for i in range(300)]
        test_indices_per_state = [list(range(3, 50, 2)), list(range(50, 120)), [], list(range(201, 300, 3))]

    class Model:
        def eval(self):
            pass

        def encode(self, x, temperature, hard, noise_ratio):
            return torch.from_numpy(z[int(x.flatten()[0])])[None, None]

    weighted, pct = ns["calculate_state_consistency"](Model(), DS(), "cpu")
    idx = np.array([i for st in DS.test_indices_per_state for i in st])
    labels = np.array([ns["assign_label"](int(i), flags) for i in idx])
    np.savez_compressed(os.path.join(OUT, "evaluation.npz"), imgs=imgs, gauss=gauss, noise=noise, gauss_mean=0.05,
                        gauss_std=0.2, occ=occ, occ_xy=xy, occ_size=size, z=z, idx=idx, labels=labels,
                        flags=np.array(flags), weighted=float(weighted), percentages=np.array(pct, dtype=np.float64))
    print("evaluation: weighted", weighted, "pct", pct, "occ size", size, xy.tolist())


DATASET_SEGMENTS = [(0, 37), (37, 90), (90, 101), (101, 160)]


def dataset_embeddings(n=160, hw=(2, 3), seed=9):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(n, 4, *hw, generator=g).numpy()
    # mixed key styles, as _load_embedding accepts both (:343-349)
    return {(f"{i:010d}.jpg" if i % 3 else f"{i:010d}"): lat[i:i + 1] for i in range(n)}


def dataset_case():
    """Reference ShuffledStatePairDataset (percep_RBVAE_train.py:181-360) on a seeded toy embedding dict."""
    import random
    cls = ref_shim.train_dataset_class()
    emb = dataset_embeddings()
    out = {}
    for mode in ("train", "val", "test"):
        random.seed(31)
        ds = cls(emb, DATASET_SEGMENTS, test_pct=0.15, val_pct=0.1, mode=mode)
        out[mode + "_pairs"] = np.array([[list(p[i % len(p)]) for p in ds.pairs_per_state] for i in range(len(ds))])
        out[mode + "_items"] = torch.stack([ds[i] for i in range(len(ds))]).numpy()
        out[mode + "_rand_after"] = random.random()
    np.savez_compressed(os.path.join(OUT, "dataset.npz"), segments=np.array(DATASET_SEGMENTS), **out)
    print("dataset:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def precompute_case(n_frames=96, target=(128, 72)):
    """The reference's embedding precompute (get_percep_embeddings.py:76-114) on the first frames of its sample
    video: per frame, in frame order, ``load_img``'s arithmetic on the decoded frame (RGB, LANCZOS to `target`,
    LANCZOS again to multiples of 32, /255, NCHW, 2x-1; :54-71 -- executed from the reference's own source text
    with ``Image.open(path)`` replaced by the in-memory frame), ``encode_first_stage`` (= first_stage_model.encode,
    ddpm.py:861-863) and ``get_first_stage_encoding`` (= scale_factor * posterior.sample(), ddpm.py:542-549) with
    the global RNG seeded ``torch.manual_seed(0)`` once before the loop; keys are the extractor's ``%010d.jpg``.
    ``LatentDiffusion`` itself cannot be imported here (omegaconf / CLIP), its two methods are the one-liners above.
    Stored: keys, the embedding dict's dtype/shape, per-frame latent sums, a few whole latents, and the codes / h a
    percep RBVAE (chinchess weights) gives on those sampled latents."""
    import ast
    import cv2
    import PIL
    from PIL import Image
    path = os.path.join(ref_shim.LIVE_ROOT, "src/stable-diffusion/get_percep_embeddings.py")
    fn = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "load_img"][0]
    src_lines = open(path).read().splitlines()[fn.lineno - 1:fn.end_lineno]
    body = "\n".join(src_lines).replace('Image.open(path).convert("RGB")', 'path.convert("RGB")')
    body = body.replace("target_size = (1280, 720)", f"target_size = {tuple(target)!r}")
    ns = {"Image": Image, "PIL": PIL, "np": np, "torch": torch, "print": lambda *a, **k: None}
    exec(body, ns)
    load_img = ns["load_img"]
    cap = cv2.VideoCapture(ref_shim.video_path())
    sd = kl_f8.init_state_dict(0)
    m = ref_shim.autoencoder_kl(sd)
    rsd, (fh, fw) = chinchess.rbvae_weights()
    L = chinchess.L
    rb = ref_shim.rbvae("percep", 4, L, rsd, feat_hw=(fh, fw))
    emb = {}
    torch.manual_seed(0)
    with torch.no_grad():
        for i in range(n_frames):
            ok, f = cap.read()
            assert ok
            img = load_img(Image.fromarray(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)))
            encoded = m.encode(img)                                  # encode_first_stage
            latent = 0.18215 * encoded.sample()                      # get_first_stage_encoding
            emb["{:010d}.jpg".format(i)] = latent.cpu().numpy()
    keys = list(emb.keys())
    lat = np.concatenate([emb[k] for k in keys])
    with torch.no_grad():
        z = torch.from_numpy(lat)
        zs = rb.encode(z.unsqueeze(1), temperature=0.5, hard=True, noise_ratio=0.0)[:, 0].numpy()
        hs = rb.encoder_rnn(rb.encoder_cnn(z).reshape(-1, 1, L))[0][:, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "precompute_chinchess.npz"), keys=np.array(keys), target=np.array(target),
                        item_shape=np.array(emb[keys[0]].shape), item_dtype=str(emb[keys[0]].dtype),
                        latent_sum=lat.astype(np.float64).sum(axis=(1, 2, 3)), latents_head=lat[:4], latents_tail=lat[-2:],
                        h=hs, z_hard=zs, noise_seed=0)
    print("precompute:", len(keys), emb[keys[0]].shape, emb[keys[0]].dtype, "distinct codes", len(np.unique(zs, axis=0)))


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-precompute" in sys.argv:
        return precompute_case()
    if "--only-chinchess" in sys.argv:
        return chinchess_case()
    if "--only-evaluation" in sys.argv:
        return evaluation_case()
    if "--only-dataset" in sys.argv:
        return dataset_case()
    if "--only-forward" in sys.argv:
        rbvae_forward_case("rbvae_forward_percep_L25_88x160", "percep", 4, 256, 4, 25, (88, 160), 11, 1, 2)
        rbvae_forward_case("rbvae_forward_contrastive_L25_256x256", "contrastive", 3, 64, 2, 25, (256, 256), 12, 1, 1)
        return losses_case()
    encoder_case("kl_f8_seed0_2x64x96_white", 0, (2, 64, 96), 1234, False)
    encoder_case("kl_f8_seed1_1x128x128_smooth", 1, (1, 128, 128), 1234, True)
    encoder_case("kl_f8_seed0_2x256x256_white", 0, (2, 256, 256), 1234, False)      # BASELINE config 1 shape
    rbvae_case("rbvae_percep_L25_32x32_T1", "percep", 4, 256, 4, 25, (32, 32), 1, 1)
    rbvae_case("rbvae_percep_L25_64x64_T1", "percep", 4, 256, 4, 25, (64, 64), 2, 1)
    rbvae_case("rbvae_percep_L100_88x160_T1", "percep", 4, 256, 4, 100, (88, 160), 3, 1)   # reference-native shape
    rbvae_case("rbvae_percep_L50_32x32_T4", "percep", 4, 256, 4, 50, (32, 32), 4, 4)
    rbvae_case("rbvae_contrastive_L25_256x256_T1", "contrastive", 3, 64, 2, 25, (256, 256), 5, 1)
    rbvae_forward_case("rbvae_forward_percep_L25_88x160", "percep", 4, 256, 4, 25, (88, 160), 11, 1, 2)
    rbvae_forward_case("rbvae_forward_contrastive_L25_256x256", "contrastive", 3, 64, 2, 25, (256, 256), 12, 1, 1)
    losses_case()
    resize_case()
    chinchess_case()
    evaluation_case()
    dataset_case()
    precompute_case()


if __name__ == "__main__":
    main()
