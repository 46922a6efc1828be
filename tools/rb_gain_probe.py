"""Probe (GPU box): for candidate RBVAE gain recipes (weights.make_rbvae_responsive), how many distinct codes 16
synthetic 512x512 frames get, how far the mixed-mode h is from the fp32 oracle's (CUDA fp32, TF32 off) and how many
bits flip outside the |h| < 1e-3 band.  Used to pick bench.py's RB_GAINS; writes gpurun_out/rb_gain_probe.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sfv_b200  # noqa: E402
import gpu_checks as c  # noqa: E402
from oracle import frames, kl_f8, rbvae as orb  # noqa: E402

R, L, N = 512, 25, 16
sd = sfv_b200.init_encoder_state_dict(0)
u8 = sfv_b200.synthetic_frames(64, R, R, 1234, smooth=True)[:N]
sd_dev = {k: v.to("cuda") for k, v in sd.items()}
x = frames.normalise_u8(u8.numpy())
ref = torch.cat([c.cuda_oracle(kl_f8.encode_moments, x[i:i + 1].cuda(), sd_dev) for i in range(N)])
lat_ref = kl_f8.SCALE_FACTOR * ref[:, :4]
out = {}
for prec in ("mixed", "fp32"):
    vae = sfv_b200.AutoencoderKL(precision=prec)
    vae.load_state_dict(sd)
    post = vae.encode_uint8(u8.cuda())
    lat = sfv_b200.FirstStage(vae).get_first_stage_mode(post)
    vae.check_async_error()
    out[prec] = dict(latent_rel_l2=float((lat - lat_ref).norm() / lat_ref.norm()))
    for fc, bg, ih in [(100, .002, 8), (40, .02, 4), (400, .02, 4), (400, .002, 4), (200, .002, 6), (400, .002, 6), (1000, .002, 4),
                       (3000, .002, 2), (150, .002, 5)]:
        rsd = sfv_b200.make_rbvae_responsive(sfv_b200.init_rbvae_state_dict(4, L, (R // 64, R // 64), seed=1), fc, bg, ih)
        rsd_dev = {k: v.cuda() for k, v in rsd.items()}
        z_ref, h_ref = c.cuda_oracle(lambda a, b: orb.encode(a, b, hard=True, noise_ratio=0.0, return_h=True), lat_ref[:, None], rsd_dev)
        rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, L, L, input_hw=(R // 8, R // 8), precision=prec)
        rb.load_state_dict(rsd)
        codes, h = rb.encode_codes(lat[:, None])
        z = sfv_b200.unpack_codes(codes, L).cpu().numpy()
        zr, hr = z_ref[:, 0].cpu().numpy(), h_ref[:, 0].cpu().numpy()
        o, i, n = c.code_flips(z, zr, hr)
        out[prec][f"{fc},{bg},{ih}"] = dict(distinct=int(len(np.unique(zr, axis=0))), band=n, flips_outside=o, flips_inside=i,
                                            h_maxabs=float(np.abs(h[:, 0].cpu().numpy() - hr).max()),
                                            h_std_over_frames=float(hr.std(0).mean()),
                                            bits_below_3e3=int((np.abs(hr) < 3e-3).sum()))
        print(prec, fc, bg, ih, out[prec][f"{fc},{bg},{ih}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rb_gain_probe.json"), "w"), indent=1)
