"""Throughput of the RBVAE decoder half (training-side forward) and of the training losses on one GPU,
with the CPU restatement timed beside it.  percep decoder at the reference's native 88x160 output.
Algorithmic FLOPs of ConvTranspose2d(k3,s2): 2 * Hin*Win * Cin*Cout * 9 per frame (what the four sub-pixel phase
convolutions execute)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import sfv_b200
from oracle import rbvae as orb          # checker / CPU baseline only

L, hw, feat, ch = 25, (88, 160), (11, 20), 256
sd = orb.init_state_dict(4, L, feat, seed=0)
sd.update(orb.init_decoder_state_dict(4, L, feat, seed=0))
m = sfv_b200.Seq2SeqBinaryVAE(4, 4, L, L, kind="percep", input_hw=hw)
m.load_state_dict(sd)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
z = torch.rand(N, 1, L, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3):
    x = m.decode(z, hw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    x = m.decode(z, hw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flops = sum(2.0 * (feat[0] << i) * (feat[1] << i) * ch * co * 9 for i, co in enumerate((ch, ch, 4))) + 2.0 * L * ch * feat[0] * feat[1]
xo, _ = orb.decode(z[:8].cpu(), sd, feat)
err = float((x[:8].cpu() - xo).abs().max())
t0 = time.perf_counter()
orb.decode(z[:32].cpu(), sd, feat)
cpu_s = time.perf_counter() - t0
from sfv_b200 import losses
xr = torch.rand(N, 1, 4, *hw).cuda()
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    losses.recon_loss(x, xr)
e1.record()
torch.cuda.synchronize()
ms_mse = e0.elapsed_time(e1) / iters
print(json.dumps(dict(frames=N, decode_ms=ms, decode_frames_per_s=N / ms * 1e3, algorithmic_gflop_per_frame=flops / 1e9,
                      algorithmic_tflops=flops * N / ms / 1e9,
                      x_recon_maxabs_vs_oracle=err, cpu_oracle_frames_per_s=32 / cpu_s, cpu_threads=torch.get_num_threads(),
                      mse_ms=ms_mse, mse_gb_per_s=2 * x.numel() * 4 / ms_mse / 1e6)))
