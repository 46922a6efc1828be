"""A/B of the fused GroupNorm transform (SFV_GN_FUSE=1, opt-in) against the default stand-alone apply pass (=0): same frames,
same weights, mixed mode.  The transform warps use the apply pass's arithmetic, so the latents must agree to the
noise of the fp64 statistics atomics (~1e-6); also reports the step time of both."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import sfv_b200
    out = {}
    vae = sfv_b200.AutoencoderKL(precision="mixed")
    vae.load_state_dict(sfv_b200.init_encoder_state_dict(0))
    for R, B in ((256, 2), (512, 8), (1024, 1)):
        u8 = sfv_b200.synthetic_frames(B, R, R, 77, smooth=True).cuda()
        for _ in range(2):
            p = vae.encode_uint8(u8).parameters
        vae.check_async_error()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p = vae.encode_uint8(u8).parameters
        e1.record()
        torch.cuda.synchronize()
        vae.check_async_error()
        out[str(R)] = dict(ms=e0.elapsed_time(e1), params=p.double().cpu().flatten()[::97].tolist(), norm=float(p.double().norm()))
    print("JSON::" + json.dumps(out))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
        sys.exit(0)
    res = {}
    for v in ("1", "0"):
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], capture_output=True, text=True,
                           env=dict(os.environ, SFV_GN_FUSE=v), timeout=600)
        got = [l for l in p.stdout.splitlines() if l.startswith("JSON::")]
        if not got:
            print("SFV_GN_FUSE=" + v, "FAILED", p.stderr[-1500:])
            sys.exit(1)
        res[v] = json.loads(got[0][6:])
    import numpy as np
    for R in res["1"]:
        a, b = np.array(res["1"][R]["params"]), np.array(res["0"][R]["params"])
        print(f"{R}x{R}: fused {res['1'][R]['ms']:.2f} ms, unfused {res['0'][R]['ms']:.2f} ms, "
              f"rel diff of the latents {np.linalg.norm(a - b) / np.linalg.norm(b):.3e}")
