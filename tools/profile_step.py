"""One forward of one chunk (8 frames 512x512 -> codes) inside a cudaProfiler range,
for ncu (--profile-from-start off).  Exits 0 without ncu too."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import sfv_b200

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
R = int(sys.argv[3]) if len(sys.argv) > 3 else 512
vae, rb, sd, rsd = bench.build_models(prec, R)
pipe = sfv_b200.FramePipeline(vae, rb, batch=B)
u8 = sfv_b200.synthetic_frames(B, R, R, 1234, smooth=True).cuda()
for _ in range(2):
    pipe.encode_device(u8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
r = pipe.encode_device(u8)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
vae.check_async_error()
print("ok", float(r.latents.abs().mean()), sfv_b200.lib().sfv_launch_count())
