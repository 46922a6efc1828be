"""Per-launch device times of one chunk forward (CUDA events bracketing every launch),
for a matrix of kernel variants selected by environment toggles.
    python tools/layer_profile.py            # runs the matrix, one subprocess per variant
    python tools/layer_profile.py --one      # run once with the current environment
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(B=8, R=512, prec=None):
    import torch
    import bench
    import sfv_b200

    prec = prec or os.environ.get("SFV_PRECISION", "mixed")
    B = int(os.environ.get("SFV_PROFILE_FRAMES", B))
    vae, rb, sd, rsd = bench.build_models(prec, R)
    pipe = sfv_b200.FramePipeline(vae, rb, batch=B)
    u8 = sfv_b200.synthetic_frames(B, R, R, 1234, smooth=True).cuda()
    for _ in range(2):
        r = pipe.encode_device(u8)
    torch.cuda.synchronize()
    lib = sfv_b200.lib()
    lib.sfv_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = pipe.encode_device(u8)
    e1.record()
    torch.cuda.synchronize()
    log = lib.sfv_profile_log().decode()
    lib.sfv_profile_enable(0)
    vae.check_async_error()
    recs = []
    for line in log.strip().splitlines():
        cat, ms, work, tag = line.split(",", 3)
        recs.append(dict(cat=int(cat), ms=float(ms), work=float(work), tag=tag))
    out = dict(total_ms=e0.elapsed_time(e1), recs=recs, lat_mean=float(r.latents.double().abs().mean()),
               lat_sum=float(r.latents.double().sum()), codes=r.codes.cpu().flatten().tolist()[:8])
    print("JSON::" + json.dumps(out))


def main():
    if "--one" in sys.argv:
        return one()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if "--current" in sys.argv:       # the product configuration (+ the fp32-stream A/B), every launch of every class
        for name, env in (("mixed_stream16", {}), ("mixed_fp32stream", dict(SFV_STREAM16="0")), ("bf16", dict(SFV_PRECISION="bf16"))):
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], capture_output=True, text=True,
                               env=dict(os.environ, **env), timeout=240)
            got = [l for l in p.stdout.splitlines() if l.startswith("JSON::")]
            if not got:
                print(name, "FAILED", p.stderr[-800:]); continue
            d = json.loads(got[0][6:])
            names = ["tc_gemm", "igemm", "gn_stats", "gn_apply", "softmax", "other"]
            agg = {}
            for r in d["recs"]:
                agg[names[r["cat"]]] = agg.get(names[r["cat"]], 0) + r["ms"]
            print(f"== {name}: total {d['total_ms']:.2f} ms  " + " ".join(f"{k}={v:.2f}" for k, v in agg.items()), flush=True)
            for i, r in enumerate(d["recs"]):
                rate = r["work"] / r["ms"] / 1e9 if r["ms"] > 0 else 0
                print(f"{i:3d} {names[r['cat']]:8s} {r['ms'] * 1e3:8.1f} us  {rate:8.1f} {'TF/s' if r['cat'] < 2 else 'GB/s'}  {r['tag']}")
        return
    variants = [("ncta1_epi0_st0", dict(SFV_NCTA="1", SFV_EPI="0", SFV_FUSED_STATS="0")),
                ("ncta1_epi0_st1", dict(SFV_NCTA="1", SFV_EPI="0", SFV_FUSED_STATS="1")),
                ("ncta1_epi1_st0", dict(SFV_NCTA="1", SFV_EPI="1", SFV_FUSED_STATS="0")),
                ("ncta1_epi1_st1", dict(SFV_NCTA="1", SFV_EPI="1", SFV_FUSED_STATS="1")),
                ("ncta2_epi0_st1", dict(SFV_NCTA="2", SFV_EPI="0", SFV_FUSED_STATS="1")),
                ("ncta2_epi1_st1", dict(SFV_NCTA="2", SFV_EPI="1", SFV_FUSED_STATS="1")),
                ("x_tmemonly", dict(SFV_NCTA="1", SFV_EPI="2", SFV_FUSED_STATS="1")),
                ("x_nostore", dict(SFV_NCTA="1", SFV_EPI="3", SFV_FUSED_STATS="1")),
                ("x_nostore_nostats", dict(SFV_NCTA="1", SFV_EPI="3", SFV_FUSED_STATS="0"))]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    res = {}
    for name, env in variants:
        if only and name not in only:
            continue
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], capture_output=True, text=True,
                               env=dict(os.environ, **env), timeout=240)
        except subprocess.TimeoutExpired:
            print(name, "TIMEOUT"); res[name] = "timeout"; continue
        got = [l for l in p.stdout.splitlines() if l.startswith("JSON::")]
        if not got:
            print(name, "FAILED rc", p.returncode, p.stderr[-600:]); res[name] = dict(err=p.stderr[-600:]); continue
        d = json.loads(got[0][6:])
        res[name] = d
        names = ["tc_gemm", "igemm", "gn_stats", "gn_apply", "softmax", "other"]
        agg = {}
        for r in d["recs"]:
            agg[names[r["cat"]]] = agg.get(names[r["cat"]], 0) + r["ms"]
        print(f"{name}: total {d['total_ms']:.2f} ms  " + " ".join(f"{k}={v:.2f}" for k, v in agg.items()) +
              f"  lat_mean={d['lat_mean']:.6f} lat_sum={d['lat_sum']:.4f} codes={d['codes']}", flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "layer_profile.json"), "w"))
    # per-launch table for the tc_gemm launches
    keys = [k for k in res if isinstance(res[k], dict) and "recs" in res[k]]
    if keys:
        rows = [[r for r in res[k]["recs"] if r["cat"] == 0] for k in keys]
        n = min(len(r) for r in rows)
        print("tc_gemm launches (us): " + " | ".join(keys))
        for i in range(n):
            tag = rows[-1][i]["tag"]
            print(f"{i:2d} " + " ".join(f"{rows[j][i]['ms'] * 1e3:8.1f}" for j in range(len(keys))) +
                  f"   TF/s(last)={rows[-1][i]['work'] / rows[-1][i]['ms'] / 1e9:7.1f}  {tag}")


if __name__ == "__main__":
    main()
