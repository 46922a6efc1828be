"""Role-level cycle accounting of the tcgen05 kernel (SFV_TC_DEBUG=1): one chunk forward,
prints per launch where the producer / MMA / epilogue roles spent their cycles."""
import os, sys
os.environ["SFV_TC_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, sfv_b200

vae, rb, sd, rsd = bench.build_models(os.environ.get("SFV_PRECISION", "mixed"))
pipe = sfv_b200.FramePipeline(vae, rb, batch=8)
u8 = sfv_b200.synthetic_frames(8, 512, 512, 1234, smooth=True).cuda()
for i in range(3):
    sys.stderr.write(f"==== pass {i}\n")
    pipe.encode_device(u8)
    torch.cuda.synchronize()
vae.check_async_error()
