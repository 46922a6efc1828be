#!/usr/bin/env python
"""bench.py -- frames/sec 512x512 -> binary code (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c4|c5]

A "step" is one pass of the hot path over one batch of synthetic frames:
uint8 [64,512,512,3] -> KL-f8 encoder -> posterior mode * 0.18215 -> percep RBVAE
encoder -> bit-packed code (latent_dim 25, noise_ratio 0).
  value : frames/s with the uint8 frames already resident in HBM (CUDA events)
  e2e   : same through the public Python API / C ABI from pinned HOST buffers,
          H2D of the frames and D2H of latents + codes inside the timed region
  roofline     : tcgen05 implicit-GEMM kernel, algorithmic FLOPs / its summed
                 CUDA-event launch durations (a second pass of the same K steps
                 with an event pair around every launch), against MEASURED_PEAKS.json
  cpu_baseline : the UNMODIFIED reference modules (oracle/_ref, placed there by
                 oracle/build_ref.py) on the host cores, on a bounded sample; the
                 oracle port is checked against them in the same run (N = 1 only;
                 at N > 1 cpu_baseline and parity are null)
N > 1 (torchrun, one rank per GPU): each rank encodes its own contiguous range of
64 frames (weak scaling, no data-path collective); its kernels write latents, codes
and h straight into its block of double-buffered gather buffers, and one in-place
NCCL all-gather per tensor runs asynchronously under the next step; time = max over
ranks.  Other BASELINE configs: --config c5 (1024x1024, batch 16), --config c4
(contrastive RBVAE on 512x512 frame pairs, batch 256).
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
# codes follow the frame (weights.make_rbvae_responsive).  Picked with tools/rb_gain_probe.py on the GPU box: 9 distinct codes
# over the 16 parity frames, 3 bits inside the |h| < 1e-3 band, and the mixed-mode h within 1.0e-3 of the oracle's -- a
# steeper recipe (100, 0.002, 8: 14 distinct codes) amplifies the 2.4e-3 latent error to 4e-3 in h, which makes flips
# OUTSIDE the band a coin toss and says nothing about the kernels
RB_GAINS = dict(fc_gain=1000.0, bias_gain=0.002, ih_gain=4.0)
PARITY_FRAMES = 16
DTYPE_NAMES = {"mixed": "mixed fp16/bf16 operands, fp32 accumulate", "bf16": "bf16", "fp16": "fp16", "fp32": "f32"}
OPERAND_FORMATS = {
    "mixed": "fp16: weights (per-layer 2^k scale), GroupNorm+SiLU outputs, conv1 outputs, x copies (x 2^-6); "
             "bf16: q, k, V^T, P, attention output; fp32: accumulators, residual stream, GroupNorm statistics",
    "bf16": "bf16 operands, fp32 accumulate / residual stream", "fp16": "fp16 operands, fp32 accumulate / residual stream",
    "fp32": "fp32 CUDA-core check mode"}


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


def make_weights(res):
    """Seeded random-init weights with the reference key names (no checkpoint exists offline).  The RBVAE's fc /
    LSTM-input weights are boosted and its biases damped so the binary code depends on the frame: on the default
    init every frame gets the same code and a code comparison would prove nothing."""
    import sfv_b200
    sd = sfv_b200.init_encoder_state_dict(0)
    gains = dict(RB_GAINS)
    if res >= 1024:
        gains["fc_gain"] = 400.0      # fc sums 4x more features at 1024x1024: same recipe, flatter, so |h| errors stay < 1e-3
    rsd = sfv_b200.make_rbvae_responsive(sfv_b200.init_rbvae_state_dict(4, LATENT_DIM, (res // 64, res // 64), seed=1),
                                         **gains)
    return sd, rsd


def build_models(precision, res=R):
    import sfv_b200
    sd, rsd = make_weights(res)
    vae = sfv_b200.AutoencoderKL(precision=precision)
    vae.load_state_dict(sd)
    # the two 256->256 RBVAE convs follow the encoder's operand mode (conv.0, fc, LSTM stay fp32)
    rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, LATENT_DIM, LATENT_DIM, input_hw=(res // 8, res // 8),
                                   precision=os.environ.get("SFV_RBVAE_PRECISION", precision))
    rb.load_state_dict(rsd)
    return vae, rb, sd, rsd


def workload_name(batch, res=R):
    """The one workload both arms are quoted on (BASELINE.json configs[1]; configs[4] at 1024)."""
    cfg = {512: "configs[1]", 1024: "configs[4]", 256: "configs[0]"}.get(res, "custom")
    return (f"BASELINE {cfg}: percep pipeline {res}x{res} uint8 frames -> KL-f8 encoder -> "
            f"4x{res // 8}x{res // 8} latent -> RBVAE binary code (latent_dim {LATENT_DIM}), batch {batch} per GPU")


# ---- the CPU side: the unmodified reference (oracle/_ref) and the oracle port ----------------------------------
def reference_models(sd, rsd, res):
    """The reference's own AutoencoderKL and percep Seq2SeqBinaryVAE (unmodified files, oracle/build_ref.py), with
    fc resized for the latent of a res x res frame (the reference hard-wires 88x160).  None if absent."""
    from oracle import ref_shim
    if not ref_shim.available():
        return None
    return ref_shim.autoencoder_kl(sd), ref_shim.rbvae("percep", 4, LATENT_DIM, rsd, feat_hw=(res // 64, res // 64))


def cpu_reference(models, frames_u8):
    """frames -> latents -> hard codes through the reference's public API, exactly the calls of
    get_percep_embeddings.py:94-101 (mode() instead of sample()) and percep_RBVAE_model.py:172-191."""
    import torch
    from oracle import frames
    vae, rb = models
    x = frames.normalise_u8(frames_u8)            # load_img: /255, HWC->NCHW, 2x-1 (get_percep_embeddings.py:67-71)
    t0 = time.time()
    with torch.no_grad():
        post = vae.encode(x)
        lat = 0.18215 * post.mode()
        z = rb.encode(lat[:, None], temperature=0.5, hard=True, noise_ratio=0.0)
    dt = time.time() - t0
    return len(frames_u8) / dt, dt, dict(z=z[:, 0].numpy(), lat=lat)


def cpu_port(sd, rsd, frames_u8):
    """The oracle port of the same path (also yields h, the thresholded LSTM state, for the |h| < 1e-3 band)."""
    from oracle import frames, kl_f8, rbvae as orb
    x = frames.normalise_u8(frames_u8)
    t0 = time.time()
    post = kl_f8.encode(x, sd)
    lat = kl_f8.first_stage_encoding(post, use_mode=True)
    z, h = orb.encode(lat[:, None], rsd, hard=True, noise_ratio=0.0, return_h=True)
    dt = time.time() - t0
    return len(frames_u8) / dt, dt, dict(z=z[:, 0].numpy(), h=h[:, 0].numpy(), lat=lat)


def port_vs_reference(ref, port):
    import numpy as np
    return dict(latent_maxabs=float((ref["lat"] - port["lat"]).abs().max()),
                latent_rel_l2=float((ref["lat"] - port["lat"]).norm() / ref["lat"].norm()),
                code_bits_differing=int((ref["z"] != port["z"]).sum()), bits=int(np.asarray(ref["z"]).size))


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core."""
    import torch
    n = os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["MKL_NUM_THREADS"] = str(n)
    torch.set_num_threads(n)
    return n


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path -- its unmodified modules from
    oracle/_ref (oracle port only if they are absent), fp32, all host threads, each step a bounded sample
    (8 frames) of the workload, same frames and weights as the GPU arm's first batch."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    os.environ["MKL_NUM_THREADS"] = str(os.cpu_count())
    import torch
    import sfv_b200                     # weights / frames generators only; nothing of ours is on the timed path
    cores = host_threads()
    res = args.res
    sd, rsd = make_weights(res)
    sample = 8 if res <= 512 else 2
    u8 = sfv_b200.synthetic_frames(args.batch, res, res, 1234, smooth=True).numpy()[:sample]
    models = reference_models(sd, rsd, res)
    kind = "reference" if models is not None else "port"
    run = (lambda f: cpu_reference(models, f)) if models is not None else (lambda f: cpu_port(sd, rsd, f))
    for _ in range(min(args.warmup, 1)):
        run(u8[:1])
    times, last = [], None
    for _ in range(args.steps):
        fps, dt, last = run(u8)
        times.append(dt)
    tot = sum(times)
    val = sample * len(times) / tot
    agree = None
    if models is not None:
        _, _, port = cpu_port(sd, rsd, u8)
        agree = port_vs_reference(last, port)
    # BASELINE configs[0] exactly as BASELINE.md section 4 defines the CPU baseline: 8 frames of 256x256, fp32,
    # 1 warm-up + best of 3
    c0 = None
    if res == 512 and not args.no_config0:
        sd0, rsd0 = make_weights(256)
        m0 = reference_models(sd0, rsd0, 256)
        u0 = sfv_b200.synthetic_frames(8, 256, 256, 1234, smooth=True).numpy()
        run0 = (lambda f: cpu_reference(m0, f)) if m0 is not None else (lambda f: cpu_port(sd0, rsd0, f))
        run0(u0)
        best = min(run0(u0)[1] for _ in range(3))
        c0 = dict(workload=workload_name(8, 256), value=8 / best, unit="frames/s", seconds_per_batch=best,
                  timing="1 warm-up + best of 3, wall clock", kind=kind, cores=cores)
    line = dict(impl="reference", metric=f"frames_per_sec_{res}x{res}_to_binary_code", value=val, unit="frames/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1000 * tot / len(times),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args.batch, res), frames_per_step_per_gpu=args.batch,
                            weights="seeded random init (no checkpoint offline), RBVAE fc / LSTM-input gains so codes follow the frame",
                            sample_frames_per_step=sample, operand_format="f32",
                            note="CPU arm: each step is a bounded sample of the batch (the first frames of the GPU arm's "
                                 "first batch, same weights)"),
                cpu_baseline=dict(value=val, unit="frames/s", cores=cores, kind=kind,
                                  sample=f"{sample} frames of {res}x{res} per step, fp32, "
                                         + ("unmodified reference modules (oracle/_ref)" if kind == "reference"
                                            else "oracle port (oracle/_ref absent)")),
                port_vs_reference=agree, baseline_config0=c0,
                e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---- BASELINE configs[3]: contrastive RBVAE on 512x512 frame pairs ----------------------------------------------
def run_contrastive(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import sfv_b200
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    L, Rc, pairs = 32, 512, args.batch
    rsd = sfv_b200.init_rbvae_state_dict(3, L, (Rc // 8, Rc // 8), channels=64, num_layers=2, seed=8)
    rb = sfv_b200.Seq2SeqBinaryVAE(3, 3, L, L, kind="contrastive", input_hw=(Rc, Rc), precision=args.precision)
    rb.load_state_dict(rsd)
    g = torch.Generator().manual_seed(99 + rank)
    n_buf = 2
    # frames in [0,1] fp32 NCHW (what the reference's dataset yields), pairs = T = 2; sub-batches of 64 pairs
    sub = 64
    bufs = [[torch.rand(sub, 2, 3, Rc, Rc, generator=g).to(dev) for _ in range(max(1, pairs // sub))] for _ in range(n_buf)]
    lib = sfv_b200.lib()

    def step(i):
        out = None
        for x in bufs[i % n_buf]:
            out = rb.encode_codes(x)
        return out

    for i in range(args.warmup):
        step(i)
    rb.check_async_error(dev)
    launches0 = lib.sfv_launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rb.check_async_error(dev)
    if rank == 0:
        n_pairs = (pairs // sub if pairs >= sub else 1) * sub
        frames_total = 2 * n_pairs * world * args.steps
        fps = frames_total / (ms / 1e3)
        # algorithmic HBM bytes per frame: fp32 NCHW frame in, 16-bit conv0/conv1 maps out+in, conv2 fp32 out + fc in
        ch = 64
        px = [(Rc // 2) ** 2, (Rc // 4) ** 2, (Rc // 8) ** 2]
        bpf = 3 * Rc * Rc * 4 + 2 * px[0] * ch * 2 + 2 * px[1] * ch * 2 + 2 * px[2] * ch * 4
        pk = peaks()
        print(json.dumps(dict(
            metric="frames_per_sec_contrastive_512x512_to_binary_code", value=fps, unit="frames/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype=DTYPE_NAMES.get(args.precision, args.precision), data="synthetic",
            config=dict(workload=f"BASELINE configs[3]: contrastive_RBVAE encoder on 512x512 frame pairs, {n_pairs} pairs per GPU per step",
                        l2_policy=f"{n_buf} rotating input sets of {n_pairs * 2 * 3 * Rc * Rc * 4 >> 20} MB (> L2)"),
            gpu_launches=int(lib.sfv_launch_count() - launches0),
            roofline=dict(bound="hbm", achieved=fps / world * bpf / 1e9, peak=pk["hbm"], unit="GB/s",
                          frac=fps / world * bpf / 1e9 / pk["hbm"], traffic=None, bytes_per_frame=bpf,
                          note="whole-path algorithmic bytes / step time (all kernels)"))), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default=os.environ.get("SFV_PRECISION", "mixed"))
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"],
                    help="c2 = BASELINE configs[1] (default, 512x512 batch 64); c5 = configs[4] (1024x1024 batch 16); "
                         "c4 = configs[3] (contrastive RBVAE on 512x512 pairs, batch 256)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-precision", action="store_true")
    ap.add_argument("--no-config0", action="store_true")
    args = ap.parse_args()
    args.res = 1024 if args.config == "c5" else R
    if args.batch is None:
        args.batch = {"c2": BATCH, "c5": 16, "c4": 256}[args.config]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.config == "c4":
        return run_contrastive(args, rank, world, local)

    import numpy as np
    import torch
    import torch.distributed as dist
    import sfv_b200

    res = args.res
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    B = args.batch
    lib = sfv_b200.lib()
    vae, rb, sd, rsd = build_models(args.precision, res)
    pipe = sfv_b200.FramePipeline(vae, rb, batch=B, device=dev)
    n_in = N_INPUT_BUFFERS if res <= 512 else 3

    # distinct synthetic frames per rank (contiguous ranges of one long synthetic video), smooth noise
    host = [sfv_b200.synthetic_frames(B, res, res, 1234 + 17 * (rank * n_in + i), smooth=True).pin_memory()
            for i in range(n_in)]
    devbuf = [h.to(dev) for h in host]
    words = (LATENT_DIM + 31) // 32
    lh = res // 8
    # multi-GPU: two sets of gather buffers; this rank's kernels write its block of set i % 2 directly and the in-place
    # all-gathers of step i run on NCCL's stream under the kernels of step i + 1
    gath = None
    if world > 1:
        gath = [dict(lat=torch.empty(world * B, 4, lh, lh, dtype=torch.float32, device=dev),
                     codes=torch.empty(world * B, words, dtype=torch.int32, device=dev),
                     h=torch.empty(world * B, LATENT_DIM, dtype=torch.float32, device=dev), works=[]) for _ in range(2)]

    def my_block(gs):
        sl = slice(rank * B, (rank + 1) * B)
        return sfv_b200.EncodeResult(gs["lat"][sl], gs["codes"][sl], gs["h"][sl])

    def gather_async(gs):
        gs["works"] = [sfv_b200.all_gather_slices(gs[k], rank, world, async_op=True) for k in ("lat", "codes", "h")]

    def gather_wait(gs):
        for w in gs["works"]:
            w.wait()
        gs["works"] = []

    def step_device(i):
        if world == 1:
            return pipe.encode_device(devbuf[i % n_in])
        gs = gath[i % 2]
        gather_wait(gs)                      # the gather that last read this set (step i - 2) is done
        r = pipe.encode_device(devbuf[i % n_in], out=my_block(gs))
        gather_async(gs)
        return r

    def drain():
        if world > 1:
            for gs in gath:
                gather_wait(gs)

    def barrier():
        drain()
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
    drain()
    vae.check_async_error()

    # ---- timed: device-resident inputs ---------------------------------------
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None))
    sampler.start()
    launches0 = lib.sfv_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    drain()                                  # the last gathers are part of the job: current stream waits for them
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.sfv_launch_count() - launches0

    # ---- the same K steps again with a CUDA event pair around every launch (on the launching stream): per-kernel-class
    #      durations for the roofline.  The events cost 0.5-2 % of the step (they keep consecutive kernels from overlapping
    #      their launch latencies), so `value` is timed without them and this pass reports its own step time.
    lib.sfv_profile_enable(1)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    drain()
    e1.record()
    barrier()
    ms_prof = max_over_ranks(e0.elapsed_time(e1))
    prof = {}
    for cat, name in enumerate(["tc_gemm", "conv_in", "gn_stats", "gn_apply", "softmax", "other"]):
        ms, work, n = C.c_double(), C.c_double(), C.c_int64()
        lib.sfv_profile_read(cat, C.byref(ms), C.byref(work), C.byref(n))
        prof[name] = dict(ms=ms.value, work=work.value, launches=n.value)
    lib.sfv_profile_enable(0)
    vae.check_async_error()

    # ---- timed: end to end from pinned host buffers ---------------------------
    def step_e2e(i):
        if world == 1:
            return pipe.encode_host(host[i % n_in], reuse_output=True)    # H2D frames, kernels, D2H latents+codes+h (pinned)
        gs = gath[i % 2]
        gather_wait(gs)
        r = pipe.encode_host(host[i % n_in], reuse_output=True, device_out=my_block(gs))
        gather_async(gs)                     # latents, codes and h gathered device to device, no host round trip
        return r

    for i in range(2):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    drain()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = B * res * res * 3
    d2h = B * (4 * lh * lh * 4 + words * 4 + LATENT_DIM * 4)

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
    # 2*MAC of encoder + quant_conv per frame
    flops_frame = {512: 1116.658466816e9, 1024: 4878.95e9}.get(res)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "tc_gemm_traffic.json")
    if os.path.exists(tp) and res == 512:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # ---- CPU baseline (the unmodified reference on a bounded sample) + parity of this very configuration ----
    cpu = None
    parity = None
    ref = None
    if not args.no_cpu_baseline and world == 1:     # the CPU leg (and the parity it anchors) belongs to the N=1 run only
        cores = host_threads()
        n_cpu = PARITY_FRAMES if res <= 512 else 2
        u8 = host[0][:n_cpu].numpy()
        models = reference_models(sd, rsd, res)
        _, dt_p, port = cpu_port(sd, rsd, u8)                 # checker: h for the band, and a cross-check of the reference
        if models is not None:
            fps, dt, ref = cpu_reference(models, u8)
            kind, agree = "reference", port_vs_reference(ref, port)
        else:
            fps, dt, ref, kind, agree = n_cpu / dt_p, dt_p, port, "port", None
        cpu = dict(value=fps, unit="frames/s", cores=cores, kind=kind,
                   sample=f"{n_cpu} frames of {res}x{res}, fp32, torch CPU threads={cores}, {dt:.1f} s: "
                          + ("the unmodified reference modules (oracle/_ref): AutoencoderKL.encode -> 0.18215*mode() -> "
                             "Seq2SeqBinaryVAE.encode(hard)" if kind == "reference" else "oracle port (oracle/_ref absent)"),
                   port_vs_reference=agree)
        # parity of this very configuration on the same frames the CPU side just encoded (checker only)
        r = pipe.encode_device(devbuf[0][:n_cpu].contiguous())
        vae.check_async_error()
        z = sfv_b200.unpack_codes(r.codes.cpu(), LATENT_DIM).numpy()
        diff = z != ref["z"]
        band = np.abs(port["h"]) < 1e-3
        parity = dict(latent_rel_l2=float((r.latents.cpu() - ref["lat"]).norm() / ref["lat"].norm()),
                      latent_gate=1e-2, against=kind,
                      code_bits=int(z.size), flips_outside_band=int((diff & ~band).sum()),
                      flips_inside_band=int((diff & band).sum()), bits_in_band=int(band.sum()),
                      distinct_codes=int(len(np.unique(ref["z"], axis=0))), sample_frames=n_cpu,
                      h_maxabs=float(np.abs(r.h.cpu().numpy() - port["h"]).max()))

    # ---- pure bf16 operands on the same workload (device-resident, same timing rules): the known miss, recorded ----
    alt = None
    if world == 1 and not args.no_alt_precision and args.precision in ("mixed", "bf16", "fp16"):
        ap_ = "bf16" if args.precision != "bf16" else "mixed"
        del pipe, vae, rb
        torch.cuda.empty_cache()
        vae2, rb2, _, _ = build_models(ap_, res)
        pipe2 = sfv_b200.FramePipeline(vae2, rb2, batch=B, device=dev)
        for i in range(3):
            pipe2.encode_device(devbuf[i % n_in])
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(args.steps):
            pipe2.encode_device(devbuf[i % n_in])
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
        metric=f"frames_per_sec_{res}x{res}_to_binary_code", value=value, unit="frames/s", n_gpus=world,
        steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype=DTYPE_NAMES.get(args.precision, args.precision), data="synthetic",
        config=dict(workload=workload_name(B, res), frames_per_step_per_gpu=B,
                    weights="seeded random init (no checkpoint offline), RBVAE fc / LSTM-input gains so codes follow the frame",
                    l2_policy=f"{n_in} rotating input batches (> L2) and a multi-GB activation working set",
                    precision=args.precision, operand_format=OPERAND_FORMATS.get(args.precision, args.precision),
                    parallelism=f"frames sharded x{world}" + (", async in-place all_gather(latents, codes, h) per step, double buffered" if world > 1 else "")),
        e2e=dict(value=e2e_val, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=ms_e2e / args.steps),
        gpu_launches=int(launches),
        clocks=sampler.summary(),
        roofline=dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                      frac=achieved / peak if peak else None, traffic=traffic,
                      kernel="tc_gemm_kernel (tcgen05 implicit GEMM: all 3x3/1x1 convs + attention GEMMs)",
                      kernel_ms_per_step=tc["ms"] / args.steps, kernel_launches_per_step=tc["launches"] / args.steps,
                      kernel_share_of_step=tc["ms"] / ms_prof if ms_prof else None,
                      measured="per-launch CUDA events on the launching stream over a second pass of the same K steps "
                               "(the events cost 0.5-2 % of a step, so `value` is timed without them)",
                      step_ms_with_events=ms_prof / args.steps,
                      peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']}); burst {pk['tensor_burst']}; "
                                  "fp16 and bf16 operands run at the same tcgen05 kind::f16 rate",
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
