"""Throughput of the BASELINE.json configurations that are not the bench.py headline (one GPU, device-resident
uint8 frames, CUDA events, 3 warm-up + 5 timed passes each):
  C1 shape 256x256 (batch 64), C2 512x512 (batch 64), C5 1024x1024 (batch 16), the reference's native 1280x704
  (batch 16), C4 contrastive RBVAE on 512x512 frames (batch 512 = 256 pairs).
Writes gpurun_out/config_sweep.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import sfv_b200

PREC = sys.argv[1] if len(sys.argv) > 1 else "bf16"


def down3(n):
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


def timed(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def flops_per_frame(H, W):
    """2*MAC of encoder + quant_conv (SURVEY 8d general formula)."""
    f = 0.0
    f += 2 * 128 * H * W * 3 * 9
    ch = [(128, 128), (128, 256), (256, 512), (512, 512)]
    h, w = H, W
    for lvl, (ci, co) in enumerate(ch):
        f += 2 * co * h * w * ci * 9 + 2 * co * h * w * co * 9            # block 0
        if ci != co:
            f += 2 * co * h * w * ci
        f += 2 * 2 * co * h * w * co * 9                                  # block 1
        if lvl != 3:
            h //= 2; w //= 2
            f += 2 * co * h * w * co * 9
    L = h * w
    f += 4 * 2 * 512 * L * 512 * 9                                        # mid block_1, block_2
    f += 4 * 2 * 512 * L * 512 + 4 * L * L * 512                          # q,k,v,proj + QK^T + PV
    f += 2 * 8 * L * 512 * 9 + 2 * 8 * L * 8
    return f


def main():
    out = {"precision": PREC, "gpu": torch.cuda.get_device_name(0)}
    sd = sfv_b200.init_encoder_state_dict(0)
    vae = sfv_b200.AutoencoderKL(precision=PREC)
    vae.load_state_dict(sd)
    for name, B, H, W in (("C1_256x256", 64, 256, 256), ("C2_512x512", 64, 512, 512), ("C5_1024x1024", 16, 1024, 1024),
                          ("native_1280x704", 16, 704, 1280)):
        rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, 25, 25, input_hw=(H // 8, W // 8), precision=PREC)
        rb.load_state_dict(sfv_b200.init_rbvae_state_dict(4, 25, (down3(H // 8), down3(W // 8)), seed=1))
        pipe = sfv_b200.FramePipeline(vae, rb, batch=B)
        u8 = sfv_b200.synthetic_frames(B, H, W, 1234, smooth=True).cuda()
        ms = timed(lambda: pipe.encode_device(u8))
        vae.check_async_error()
        fl = flops_per_frame(H, W)
        out[name] = dict(batch=B, ms=ms, frames_per_s=B / ms * 1e3, gflop_per_frame=fl / 1e9,
                         pipeline_tflops=fl * B / ms / 1e9)
        del rb, pipe, u8
        torch.cuda.empty_cache()
    # C4: contrastive RBVAE on pixel frames in [0,1]
    B, R, L = 512, 512, 25
    rb = sfv_b200.Seq2SeqBinaryVAE(3, 3, L, L, kind="contrastive", input_hw=(R, R), precision=PREC)
    rb.load_state_dict(sfv_b200.init_rbvae_state_dict(3, L, (down3(R), down3(R)), channels=64, num_layers=2, seed=2))
    x = torch.rand(B // 8, 1, 3, R, R, device="cuda")        # fp32 frames; 64 per call keeps the input at 201 MB
    ms = timed(lambda: rb.encode_codes(x, noise_ratio=0.0))
    n = x.shape[0]
    # algorithmic bytes per frame: fp32 input 3 MB + conv0 out (16-bit) 8.4 MB w+r + conv1 out 2.1 MB w+r + conv2 out fp32 1 MB w+r
    bytes_frame = R * R * 3 * 4 + 2 * (R // 2) ** 2 * 64 * 2 + 2 * (R // 4) ** 2 * 64 * (2 if PREC != "fp32" else 4) + 2 * (R // 8) ** 2 * 64 * 4
    out["C4_contrastive_512x512"] = dict(frames_per_call=n, ms=ms, frames_per_s=n / ms * 1e3,
                                         algorithmic_GB_per_s=bytes_frame * n / ms / 1e6)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
