#!/usr/bin/env python
"""bench.py -- frames/sec 512x512 -> binary code (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic frames:
uint8 [64,512,512,3] -> KL-f8 encoder -> posterior mode * 0.18215 -> percep RBVAE
encoder -> bit-packed code (latent_dim 25, noise_ratio 0).
  value : frames/s with the uint8 frames already resident in HBM (CUDA events)
  e2e   : same through the public Python API / C ABI from pinned HOST buffers,
          H2D of the frames and D2H of latents + codes inside the timed region
  roofline     : tcgen05 implicit-GEMM kernel, algorithmic FLOPs / its summed
                 CUDA-event launch durations, against MEASURED_PEAKS.json
  cpu_baseline : the oracle port of the reference's CPU path on a bounded sample
N > 1 (torchrun, one rank per GPU): each rank encodes its own contiguous range of
64 frames (weak scaling, no data-path collective) and the packed codes + latents
are all-gathered with NCCL inside the step; time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R, BATCH, LATENT_DIM = 512, 64, 25
N_INPUT_BUFFERS = 4          # 4 x 50 MB of distinct frames > 126 MB L2: inputs never L2-resident across steps


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"],
                    source="measured")
    return dict(tensor_burst=1590.0, tensor_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed regions: NVML every 20 ms (nvidia-smi subprocess fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, uuid=None):
        super().__init__(daemon=True)
        self.idx, self.stop_flag = gpu_index, False
        self.sm, self.mx, self.reasons, self.power, self.source = [], [], set(), [], "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid is not None:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            self.h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) >= 9:
            self.sm.append(float(r[1])); self.mx.append(float(r[2]))
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        sm = sorted(self.sm)
        out = dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                   reasons=sorted(self.reasons), samples=len(sm), source=self.source)
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


def build_models(precision):
    import torch
    import sfv_b200
    sd = sfv_b200.init_encoder_state_dict(0)       # seeded random init with the reference key names
    rsd = sfv_b200.init_rbvae_state_dict(4, LATENT_DIM, (R // 64, R // 64), seed=1)
    vae = sfv_b200.AutoencoderKL(precision=precision)
    vae.load_state_dict(sd)
    # the two 256->256 RBVAE convs follow the encoder's operand format (conv.0, fc, LSTM stay fp32)
    rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, LATENT_DIM, LATENT_DIM, input_hw=(R // 8, R // 8),
                                   precision=os.environ.get("SFV_RBVAE_PRECISION", precision))
    rb.load_state_dict(rsd)
    return vae, rb, sd, rsd


def workload_name(batch):
    """The one workload both arms are quoted on (BASELINE.json configs[1])."""
    return (f"BASELINE configs[1]: percep pipeline {R}x{R} uint8 frames -> KL-f8 encoder -> "
            f"4x{R // 8}x{R // 8} latent -> RBVAE binary code (latent_dim {LATENT_DIM}), batch {batch} per GPU")


def cpu_port_fps(sd, rsd, n_frames, frames_u8):
    """Oracle port of the reference CPU path (fp32, all host threads) on n_frames frames."""
    import torch
    from oracle import frames, kl_f8, rbvae as orb
    torch.set_num_threads(os.cpu_count())
    x = frames.normalise_u8(frames_u8[:n_frames])
    t0 = time.time()
    post = kl_f8.encode(x, sd)
    lat = kl_f8.first_stage_encoding(post, use_mode=True)
    z, h = orb.encode(lat[:, None], rsd, hard=True, noise_ratio=0.0, return_h=True)
    dt = time.time() - t0
    return n_frames / dt, dt, dict(z=z[:, 0].numpy(), h=h[:, 0].numpy(), lat=lat)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference
    is Python on PyTorch and cannot travel to the GPU box, so the timed code is the oracle
    port (oracle/kl_f8.py, oracle/rbvae.py: the reference's call sequence on the same ATen
    CPU kernels), fp32, all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    os.environ["MKL_NUM_THREADS"] = str(os.cpu_count())
    import torch
    import sfv_b200                     # weights / frames generators only; nothing of ours is on the timed path
    sd = sfv_b200.init_encoder_state_dict(0)
    rsd = sfv_b200.init_rbvae_state_dict(4, LATENT_DIM, (R // 64, R // 64), seed=1)
    sample = 2
    u8 = sfv_b200.synthetic_frames(sample, R, R, 1234, smooth=True).numpy()
    for _ in range(min(args.warmup, 1)):
        cpu_port_fps(sd, rsd, 1, u8)
    times = []
    for _ in range(args.steps):
        fps, dt, _ = cpu_port_fps(sd, rsd, sample, u8)
        times.append(dt)
    tot = sum(times)
    val = sample * len(times) / tot
    line = dict(impl="reference", metric="frames_per_sec_512x512_to_binary_code", value=val, unit="frames/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1000 * tot / len(times),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args.batch), frames_per_step_per_gpu=args.batch,
                            weights="seeded random init (no checkpoint offline)",
                            sample_frames_per_step=sample, operand_format="f32",
                            note="CPU arm: each step is a bounded sample of the batch (same frames, same weights)"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=os.cpu_count(), kind="port",
                                  sample=f"{sample} frames of {R}x{R} per step, fp32, torch CPU"),
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default=os.environ.get("SFV_PRECISION", "bf16"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-precision", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import sfv_b200

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    B = args.batch
    lib = sfv_b200.lib()
    vae, rb, sd, rsd = build_models(args.precision)
    pipe = sfv_b200.FramePipeline(vae, rb, batch=B, device=dev)

    # distinct synthetic frames per rank (contiguous ranges of one long synthetic video), smooth noise
    host = [sfv_b200.synthetic_frames(B, R, R, 1234 + 17 * (rank * N_INPUT_BUFFERS + i), smooth=True).pin_memory()
            for i in range(N_INPUT_BUFFERS)]
    devbuf = [h.to(dev) for h in host]
    words = (LATENT_DIM + 31) // 32
    gather_codes = torch.empty(world * B, words, dtype=torch.int32, device=dev) if world > 1 else None
    gather_lat = torch.empty(world * B, 4, R // 8, R // 8, dtype=torch.float32, device=dev) if world > 1 else None

    def step_device(i):
        r = pipe.encode_device(devbuf[i % N_INPUT_BUFFERS])
        if world > 1:
            dist.all_gather_into_tensor(gather_codes, r.codes)
            dist.all_gather_into_tensor(gather_lat, r.latents)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up -------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    vae.check_async_error()

    # ---- timed: device-resident inputs ---------------------------------------
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None))
    sampler.start()
    lib.sfv_profile_enable(1)
    launches0 = lib.sfv_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.sfv_launch_count() - launches0
    prof = {}
    for cat, name in enumerate(["tc_gemm", "conv_in", "gn_stats", "gn_apply", "softmax", "other"]):
        ms, work, n = C.c_double(), C.c_double(), C.c_int64()
        lib.sfv_profile_read(cat, C.byref(ms), C.byref(work), C.byref(n))
        prof[name] = dict(ms=ms.value, work=work.value, launches=n.value)
    lib.sfv_profile_enable(0)
    vae.check_async_error()

    # ---- timed: end to end from pinned host buffers ---------------------------
    for i in range(2):
        pipe.encode_host(host[i % N_INPUT_BUFFERS])
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        res = pipe.encode_host(host[i % N_INPUT_BUFFERS], reuse_output=True)   # H2D frames, kernels, D2H latents+codes+h (pinned)
        if world > 1:
            dist.all_gather_into_tensor(gather_codes, res.codes.to(dev))
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = B * R * R * 3
    d2h = B * (4 * (R // 8) ** 2 * 4 + words * 4 + LATENT_DIM * 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    total_frames = B * world * args.steps
    value = total_frames / (ms_dev / 1e3)
    e2e_val = total_frames / (ms_e2e / 1e3)
    tc = prof["tc_gemm"]
    achieved = tc["work"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    peak = pk["tensor_sustained"]     # the kernel is timed inside a long step -> sustained figure
    flops_frame = 1116.658466816e9 if R == 512 else None      # 2*MAC of encoder + quant_conv per 512^2 frame
    traffic = None
    tp = os.path.join(ROOT, "profiles", "tc_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # ---- parity of this very configuration vs the oracle on a bounded sample + CPU baseline ----
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        n_cpu = 2
        fps, dt, ref = cpu_port_fps(sd, rsd, n_cpu, host[0][:n_cpu].numpy())
        if dt < 8:                                # bounded sample: ~10-30 s of CPU work
            n2 = min(B, max(n_cpu, int(n_cpu * 15 / dt)))
            if n2 > n_cpu:
                fps, dt, ref = cpu_port_fps(sd, rsd, n2, host[0][:n2].numpy())
                n_cpu = n2
        cpu = dict(value=fps, unit="frames/s", cores=os.cpu_count(), kind="port",
                   sample=f"{n_cpu} frames of {R}x{R} (oracle port of the reference path, fp32, torch CPU threads={os.cpu_count()}), {dt:.1f} s")
        # parity of this very configuration on the same frames the CPU port just encoded (checker only)
        r = pipe.encode_device(devbuf[0][:n_cpu].contiguous())
        z = sfv_b200.unpack_codes(r.codes.cpu(), LATENT_DIM).numpy()
        diff = z != ref["z"]
        band = np.abs(ref["h"]) < 1e-3
        parity = dict(latent_rel_l2=float((r.latents.cpu() - ref["lat"]).norm() / ref["lat"].norm()),
                      code_bits=int(z.size), flips_outside_band=int((diff & ~band).sum()),
                      flips_inside_band=int((diff & band).sum()), sample_frames=n_cpu,
                      h_maxabs=float(np.abs(r.h.cpu().numpy() - ref["h"]).max()))

    # ---- the other 16-bit operand format on the same workload (device-resident, same timing rules) ----
    alt = None
    if world == 1 and not args.no_alt_precision and args.precision in ("bf16", "fp16"):
        ap_ = "fp16" if args.precision == "bf16" else "bf16"
        del pipe, vae, rb
        torch.cuda.empty_cache()
        vae2, rb2, _, _ = build_models(ap_)
        pipe2 = sfv_b200.FramePipeline(vae2, rb2, batch=B, device=dev)
        for i in range(3):
            pipe2.encode_device(devbuf[i % N_INPUT_BUFFERS])
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(args.steps):
            pipe2.encode_device(devbuf[i % N_INPUT_BUFFERS])
        e1.record()
        torch.cuda.synchronize(dev)
        alt = dict(operand_format=ap_, value=B * args.steps / (e0.elapsed_time(e1) / 1e3), unit="frames/s")
        if parity is not None:
            r2 = pipe2.encode_device(devbuf[0][:parity["sample_frames"]].contiguous())
            z2 = sfv_b200.unpack_codes(r2.codes.cpu(), LATENT_DIM).numpy()
            d2 = z2 != ref["z"]
            alt.update(latent_rel_l2=float((r2.latents.cpu() - ref["lat"]).norm() / ref["lat"].norm()),
                       flips_outside_band=int((d2 & ~band).sum()), flips_inside_band=int((d2 & band).sum()))
        vae2.check_async_error()

    line = dict(
        metric="frames_per_sec_512x512_to_binary_code", value=value, unit="frames/s", n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype=args.precision, data="synthetic",
        config=dict(workload=workload_name(B), frames_per_step_per_gpu=B, weights="seeded random init (no checkpoint offline)",
                    l2_policy=f"{N_INPUT_BUFFERS} rotating input batches (> L2) and a multi-GB activation working set",
                    operand_format=args.precision, parallelism=f"frames sharded x{world}, all_gather(codes, latents)"),
        e2e=dict(value=e2e_val, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=ms_e2e / args.steps),
        gpu_launches=int(launches),
        clocks=sampler.summary(),
        roofline=dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                      frac=achieved / peak if peak else None, traffic=traffic,
                      kernel="tc_gemm_kernel (tcgen05 implicit GEMM: all 3x3/1x1 convs + attention GEMMs)",
                      kernel_ms_per_step=tc["ms"] / args.steps, kernel_launches_per_step=tc["launches"] / args.steps,
                      kernel_share_of_step=tc["ms"] / ms_dev if ms_dev else None,
                      peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']}); burst {pk['tensor_burst']}",
                      pipeline_tflops=flops_frame * value / world / 1e12 if flops_frame else None,
                      pipeline_frac_of_peak=flops_frame * value / world / 1e12 / peak if flops_frame else None),
        kernel_classes={k: dict(ms_per_step=v["ms"] / args.steps, launches_per_step=v["launches"] / args.steps,
                                rate=(v["work"] / (v["ms"] * 1e-3) / (1e12 if k in ("tc_gemm", "conv_in") else 1e9))
                                if v["ms"] > 0 else None,
                                rate_unit="TFLOP/s" if k in ("tc_gemm", "conv_in") else "GB/s")
                        for k, v in prof.items()},
        cpu_baseline=cpu, parity=parity, alt_precision=alt)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
