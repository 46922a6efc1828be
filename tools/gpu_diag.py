"""Run every GPU parity check in its own subprocess (a faulting kernel cannot take
the following checks down) and write gpurun_out/diag.json.  Usage on the GPU box:
    python tools/gpu_diag.py [--only substr] [--timeout 180]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan():
    items = []
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpu_checks as c
    for prec in ("fp32", "bf16", "fp16", "mixed"):
        for i, s in enumerate(c.CONV_SHAPES):
            N, H, W, Cin, Cout, ks, st, pad, res = s
            items.append((f"conv/{prec}/geom{i}", f"check_conv({prec!r},{N},{H},{W},{Cin},{Cout},{ks},{st},{pad},residual={res},seed={i})"))
    for prec in ("bf16", "fp32"):
        for i, s in enumerate(c.LAYER_SHAPES_64):
            N, H, W, Cin, Cout, ks, st, pad, res = s
            items.append((f"conv/{prec}/layer{i}", f"check_conv({prec!r},{N},{H},{W},{Cin},{Cout},{ks},{st},{pad},residual={res},seed={100 + i})"))
    items += [("gn/128", "check_group_norm(128,4096)"), ("gn/512", "check_group_norm(512,77,silu=False)"),
              ("gn/64", "check_group_norm(64,333)"),
              ("attn/fp32", "check_attention('fp32',2,256)"), ("attn/bf16", "check_attention('bf16',2,256)"),
              ("attn/fp16", "check_attention('fp16',2,1024)"), ("attn/bf16/L144", "check_attention('bf16',2,144)"),
              ("attn/bf16/L125", "check_attention('bf16',3,125)"), ("attn/fp16/L77", "check_attention('fp16',1,77)"),
              ("enc/odd_tokens", "check_odd_token_count()"), ("enc/odd_tokens/mixed", "check_odd_token_count('mixed')"),
              ("shapes/fp16", "check_shape_sweep('fp16')"), ("shapes/mixed", "check_shape_sweep('mixed')"),
              ("shapes/fp32", "check_shape_sweep('fp32')"), ("shapes/bf16", "check_shape_sweep('bf16')"),
              ("range_safety", "check_range_safety()"),
              ("resize", "check_resize()"), ("hamming", "check_hamming()")]
    for prec in ("fp32", "mixed", "bf16", "fp16"):
        items.append((f"taps/{prec}", f"check_encoder_taps({prec!r})"))
    for name in ("kl_f8_seed0_2x64x96_white", "kl_f8_seed1_1x128x128_smooth", "kl_f8_seed0_2x256x256_white"):
        for prec in ("fp32", "mixed", "fp16", "bf16"):
            items.append((f"enc/{prec}/{name}", f"check_encoder_golden({prec!r},{name!r})"))
    items.append(("chunking/fp32", "check_chunking_and_batch_independence('fp32')"))
    items.append(("chunking/mixed", "check_chunking_and_batch_independence('mixed')"))
    items.append(("chunking/bf16", "check_chunking_and_batch_independence('bf16')"))
    for name in ("rbvae_percep_L25_32x32_T1", "rbvae_percep_L25_64x64_T1", "rbvae_percep_L100_88x160_T1",
                 "rbvae_percep_L50_32x32_T4", "rbvae_contrastive_L25_256x256_T1"):
        items.append((f"rbvae/{name}", f"check_rbvae_golden({name!r})"))
    for name in ("rbvae_percep_L25_64x64_T1", "rbvae_percep_L100_88x160_T1", "rbvae_contrastive_L25_256x256_T1"):
        for prec in ("bf16", "fp16", "mixed"):
            items.append((f"rbvae_tc/{prec}/{name}", f"check_rbvae_tensor_core({name!r},{prec!r})"))
    for prec in ("fp32", "mixed", "fp16", "bf16"):
        items.append((f"pipeline/{prec}", f"check_pipeline({prec!r})"))
    items.append(("fullsize/mixed", "check_full_size_properties('mixed',4,512)"))
    items += [("native/mixed", "check_native_frame_size('mixed')"), ("native/fp16", "check_native_frame_size('fp16')"),
              ("native/bf16", "check_native_frame_size('bf16')"),
              ("large1024/mixed", "check_large_frame_properties('mixed',1024,2)"),
              ("contrastive512/fp32", "check_contrastive_512('fp32')"), ("contrastive512/bf16", "check_contrastive_512('bf16')"),
              ("contrastive512/mixed", "check_contrastive_512('mixed')"),
              ("chinchess/fp32", "check_chinchess_video('fp32')"), ("chinchess/mixed", "check_chinchess_video('mixed')"),
              ("chinchess/fp16", "check_chinchess_video('fp16')"), ("chinchess/bf16", "check_chinchess_video('bf16')"),
              ("conv_in_tc/fp16", "check_conv_in_tensor_core('fp16')"), ("conv_in_tc/bf16", "check_conv_in_tensor_core('bf16')"),
              ("conv_in_tc/mixed", "check_conv_in_tensor_core('mixed')"),
              ("eval/kernels", "check_evaluation_kernels()"), ("eval/pipeline/fp32", "check_state_consistency_pipeline('fp32')"),
              ("eval/pipeline/mixed", "check_state_consistency_pipeline('mixed')"),
              ("eval/pipeline/bf16", "check_state_consistency_pipeline('bf16')"),
              ("edge", "check_edge_cases()"),
              ("fullsize_oracle/mixed/512", "check_full_size_oracle('mixed',512,4)"),
              ("fullsize_oracle/fp32/512", "check_full_size_oracle('fp32',512,1)"),
              ("fullsize_oracle/mixed/1024", "check_full_size_oracle('mixed',1024,2)"),
              ("fullsize_oracle/fp32/1024", "check_full_size_oracle('fp32',1024,1)"),
              ("fullsize_oracle/bf16/512", "check_full_size_oracle('bf16',512,4)"),
              ("precompute/mixed", "check_precompute_driver('mixed')"), ("precompute/fp32", "check_precompute_driver('fp32')"),
              ("dataset/device", "check_resident_dataset()"),
              ("gn_fused", "check_gn_fused_transform()"),
              ("sampling/fp32", "check_encode_host_sampling('fp32')"), ("sampling/mixed", "check_encode_host_sampling('mixed')"),
              ("decoder/percep", "check_decoder_golden('rbvae_forward_percep_L25_88x160')"),
              ("decoder/contrastive", "check_decoder_golden('rbvae_forward_contrastive_L25_256x256')"),
              ("decoder/shapes", "check_decoder_shapes()"), ("losses", "check_losses()")]
    return items


# pure bf16 operands miss the 1e-2 latent gate on random-init weights (operand-rounding floor, oracle/numerics_model.py):
# these checks assert the north-star gate like every other mode and are EXPECTED to fail (pytest marks them xfail)
KNOWN_MISS = ("enc/bf16/", "taps/bf16", "pipeline/bf16", "native/bf16", "chinchess/bf16", "shapes/bf16",
              "fullsize_oracle/bf16")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--group", type=int, default=12, help="checks per subprocess")
    a = ap.parse_args()
    only = [t for t in a.only.split(",") if t] or [""]
    items = [it for it in plan() if any(t in it[0] for t in only)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = {}
    t_all = time.time()
    for g0 in range(0, len(items), a.group):
        grp = items[g0:g0 + a.group]
        code = ["import sys, json, traceback", f"sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})",
                "import gpu_checks as c", "from gpu_checks import *", "res = {}"]
        for name, expr in grp:
            code += [f"try:\n    res[{name!r}] = dict(ok=True, out=c.{expr})\nexcept BaseException as e:\n"
                     f"    res[{name!r}] = dict(ok=False, err=repr(e)[:1500], tb=traceback.format_exc()[-1500:])",
                     f"print('DIAG', {name!r}, res[{name!r}]['ok'], flush=True)"]
        code.append("print('JSON::' + json.dumps(res, default=str))")
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-c", "\n".join(code)], capture_output=True, text=True,
                               timeout=a.timeout * len(grp), cwd=ROOT)
            got = None
            for line in p.stdout.splitlines():
                if line.startswith("JSON::"):
                    got = json.loads(line[6:])
            if got is None:
                done = [l.split()[1].strip("'") for l in p.stdout.splitlines() if l.startswith("DIAG")]
                for name, _ in grp:
                    results[name] = dict(ok=False, err="subprocess died", rc=p.returncode, finished=name in done,
                                         stderr=p.stderr[-1500:])
            else:
                results.update(got)
        except subprocess.TimeoutExpired as e:
            for name, _ in grp:
                results[name] = dict(ok=False, err="timeout", stdout=(e.stdout or b"")[-800:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-800:])
        print(f"[{time.time() - t_all:6.1f}s] group {g0 // a.group}: " +
              " ".join(f"{n}={'ok' if results[n]['ok'] else 'FAIL'}" for n, _ in grp), flush=True)
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1, default=str)
    miss = [n for n, r in results.items() if not r["ok"] and n.startswith(KNOWN_MISS)]
    bad = [n for n, r in results.items() if not r["ok"] and not n.startswith(KNOWN_MISS)]
    print(f"{len(results) - len(bad) - len(miss)}/{len(results)} checks ok; known bf16 misses of the 1e-2 gate: {miss}; failed: {bad}")
    for n in bad[:40]:
        print("----", n, results[n].get("err", "")[:600])


if __name__ == "__main__":
    main()
