"""The "library bar" of SURVEY 8d: PyTorch eager + cuDNN/cuBLAS in bf16 (channels_last) on the same B200,
running the same KL-f8 encoder graph with the same weights and the same frames as the product.
It is a yard-stick, not a product path: nothing in symbols-from-video_b200/ imports it.

    python tools/library_bar.py [--batch 8] [--size 512] [--iters 10]

Writes gpurun_out/library_bar.json: frames/s of (a) torch eager bf16, (b) libsfv bf16, (c) libsfv fp16,
per-stage torch timings, and the rel-L2 of each against torch fp32 eager on the GPU (TF32 off).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sfv_b200  # noqa: E402
from sfv_b200.weights import init_encoder_state_dict, synthetic_frames  # noqa: E402


class TorchEncoder:
    """model.py:368-459 + autoencoder.py:324-328 written with torch.nn.functional only."""

    def __init__(self, sd, dtype, channels_last=True):
        self.dt = dtype
        self.cl = channels_last
        self.sd = {}
        for k, v in sd.items():
            v = v.cuda().to(dtype)
            if v.dim() == 4 and channels_last:
                v = v.contiguous(memory_format=torch.channels_last)
            self.sd[k] = v

    def conv(self, x, name, stride=1, padding=0):
        return F.conv2d(x, self.sd[name + ".weight"], self.sd[name + ".bias"], stride=stride, padding=padding)

    def gn(self, x, name, silu=True):
        y = F.group_norm(x, 32, self.sd[name + ".weight"], self.sd[name + ".bias"], eps=1e-6)
        return F.silu(y) if silu else y

    def res(self, x, name):
        h = self.conv(self.gn(x, name + ".norm1"), name + ".conv1", padding=1)
        h = self.conv(self.gn(h, name + ".norm2"), name + ".conv2", padding=1)
        if name + ".nin_shortcut.weight" in self.sd:
            x = self.conv(x, name + ".nin_shortcut")
        return x + h

    def attn(self, x, name):
        h = self.gn(x, name + ".norm", silu=False)
        q, k, v = (self.conv(h, f"{name}.{n}") for n in "qkv")
        B, C, H, W = q.shape
        q, k, v = (t.reshape(B, C, H * W).transpose(1, 2).unsqueeze(1) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)                  # library flash/cuDNN attention
        o = o.squeeze(1).transpose(1, 2).reshape(B, C, H, W)
        return x + self.conv(o, name + ".proj_out")

    def __call__(self, x, marks=None):
        def mark(tag):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((tag, e))
        x = x.to(self.dt)
        if self.cl:
            x = x.contiguous(memory_format=torch.channels_last)
        mark("start")
        h = self.conv(x, "encoder.conv_in", padding=1); mark("conv_in")
        for lvl in range(4):
            for blk in range(2):
                h = self.res(h, f"encoder.down.{lvl}.block.{blk}")
            if lvl != 3:
                h = self.conv(F.pad(h, (0, 1, 0, 1)), f"encoder.down.{lvl}.downsample.conv", stride=2)
            mark(f"level{lvl}")
        h = self.res(h, "encoder.mid.block_1"); mark("mid1")
        h = self.attn(h, "encoder.mid.attn_1"); mark("attn")
        h = self.res(h, "encoder.mid.block_2"); mark("mid2")
        h = self.conv(self.gn(h, "encoder.norm_out"), "encoder.conv_out", padding=1)
        h = self.conv(h, "quant_conv"); mark("tail")
        return h.float()


def time_fn(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    B, R = args.batch, args.size
    sd = init_encoder_state_dict(0)
    u8 = synthetic_frames(B, R, R, 1234, smooth=True).cuda()
    x = (u8.float() / 127.5 - 1.0).permute(0, 3, 1, 2).contiguous()
    out = {"batch": B, "size": R, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}

    with torch.no_grad():
        ref = TorchEncoder(sd, torch.float32, channels_last=False)(x[:2])[:, :4]      # fp32 yard-stick, 2 frames
        eager = TorchEncoder(sd, torch.bfloat16)
        ms = time_fn(lambda: eager(x), args.iters)
        out["torch_eager_bf16"] = {"ms": ms, "fps": B / ms * 1e3,
                                   "rel_l2_vs_fp32": float((eager(x[:2])[:, :4] - ref).norm() / ref.norm())}
        marks = []
        eager(x, marks); torch.cuda.synchronize()
        out["torch_eager_bf16"]["stages_ms"] = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1])
                                                for i in range(1, len(marks))}
        eager16 = TorchEncoder(sd, torch.float16)
        ms = time_fn(lambda: eager16(x), args.iters)
        out["torch_eager_fp16"] = {"ms": ms, "fps": B / ms * 1e3,
                                   "rel_l2_vs_fp32": float((eager16(x[:2])[:, :4] - ref).norm() / ref.norm())}
        del eager, eager16
        torch.cuda.empty_cache()
        for prec in ("bf16", "fp16"):
            vae = sfv_b200.AutoencoderKL(precision=prec)
            vae.load_state_dict(sd)
            vae = vae.cuda()
            ms = time_fn(lambda: vae.encode_uint8(u8), args.iters)
            got = vae.encode_uint8(u8[:2]).mean
            out[f"libsfv_{prec}"] = {"ms": ms, "fps": B / ms * 1e3,
                                     "rel_l2_vs_fp32": float((got - ref).norm() / ref.norm())}
    out["speedup_bf16"] = out["libsfv_bf16"]["fps"] / out["torch_eager_bf16"]["fps"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "library_bar.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
