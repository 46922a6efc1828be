"""Drop-in host mirror of the reference's KL-f8 first stage.

Same names, arguments and error behaviour as the reference classes, with the
arithmetic done by libsfv (hand-written sm_100a CUDA behind a C ABI):

* ``AutoencoderKL``                -- src/stable-diffusion/ldm/models/autoencoder.py:285-328
* ``DiagonalGaussianDistribution`` -- src/stable-diffusion/ldm/modules/distributions/distributions.py:24-62
* ``FirstStage``                   -- the two ``LatentDiffusion`` methods the repo calls,
                                      src/stable-diffusion/ldm/models/diffusion/ddpm.py:542-549,825-863

Only ``encode`` is implemented (the decoder is outside the hot path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

KL_F8_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                      ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
SCALE_FACTOR = 0.18215  # configs/stable-diffusion/v1-inference.yaml:17


def encoder_param_shapes():
    """(name -> shape) of the 108 tensors the C side ingests."""
    shapes = {}

    def conv(n, ci, co, k):
        shapes[n + ".weight"] = (co, ci, k, k)
        shapes[n + ".bias"] = (co,)

    def norm(n, c):
        shapes[n + ".weight"] = (c,)
        shapes[n + ".bias"] = (c,)

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co, 3)
        norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".nin_shortcut", ci, co, 1)

    conv("encoder.conv_in", 3, 128, 3)
    cin = 128
    for lvl, m in enumerate((1, 2, 4, 4)):
        for blk in range(2):
            res(f"encoder.down.{lvl}.block.{blk}", cin, 128 * m)
            cin = 128 * m
        if lvl != 3:
            conv(f"encoder.down.{lvl}.downsample.conv", cin, cin, 3)
    res("encoder.mid.block_1", 512, 512)
    norm("encoder.mid.attn_1.norm", 512)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", 512, 512, 1)
    res("encoder.mid.block_2", 512, 512)
    norm("encoder.norm_out", 512)
    conv("encoder.conv_out", 512, 8, 3)
    conv("quant_conv", 8, 8, 1)
    return shapes


class DiagonalGaussianDistribution(object):
    """distributions.py:24-62.  Built either from a moments tensor (reference
    constructor) or directly from the tensors the native head kernel wrote."""

    def __init__(self, parameters, deterministic=False, _native=None):
        self.parameters = parameters
        self.deterministic = deterministic
        self.mean, raw_logvar = torch.chunk(parameters, 2, dim=1)
        if _native is not None:
            self.logvar, self.std, self.var = _native
        else:
            self.logvar = torch.clamp(raw_logvar, -30.0, 20.0)
            self.std = torch.exp(0.5 * self.logvar)
            self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self, noise=None):
        # the reference draws the noise on the CPU from the global RNG and moves it (distributions.py:36);
        # keeping that call keeps seeded runs reproducible against it
        if noise is None:
            noise = torch.randn(self.mean.shape)
        noise = noise.to(device=self.parameters.device, dtype=torch.float32).contiguous()
        if not self.parameters.is_cuda or self.deterministic:
            return self.mean + self.std * noise
        return _scaled_sample(self, noise, 1.0)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var
                               - 1.0 - self.logvar + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.])
        logtwopi = np.log(2.0 * np.pi)
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)


def _scaled_sample(post, noise, scale, out=None):
    """scale * (mean + std * noise) (noise None -> scale * mean) on the device; `out`: optional destination
    (e.g. this rank's slice of an all-gather buffer, so no staging copy precedes the collective)."""
    _lib.require_cuda(post.parameters, "posterior")
    mean = post.mean.contiguous()
    if post.deterministic:            # distributions.py:29-30: std = var = 0 -> the sample is the mean
        noise = None
    if out is None:
        out = torch.empty_like(mean)
    elif (tuple(out.shape) != tuple(mean.shape) or out.dtype != torch.float32 or out.device != mean.device
          or not out.is_contiguous()):
        raise ValueError("out must be a contiguous fp32 tensor shaped like the posterior mean, on its device")
    lv = post.logvar.contiguous()
    with _lib.on_device(mean):
        _lib.check(_lib.lib().sfv_posterior_sample(_lib.ptr(mean), _lib.ptr(lv), _lib.ptr(noise), float(scale),
                                                   _lib.ptr(out), mean.numel(), _lib.stream_ptr(mean.device)))
    return out


class _Leaf(nn.Module):
    def __init__(self, wshape, bshape):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(wshape), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(bshape), requires_grad=False)


def _build_tree(root: nn.Module, shapes: dict):
    """Create a module tree whose state_dict keys are exactly `shapes`' keys."""
    groups = {}
    for k in shapes:
        groups.setdefault(k.rsplit(".", 1)[0], {})[k.rsplit(".", 1)[1]] = shapes[k]
    for path, d in groups.items():
        parts = path.split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, nn.Module())
            m = getattr(m, p)
        m.add_module(parts[-1], _Leaf(d["weight"], d["bias"]))


class AutoencoderKL(nn.Module):
    """``AutoencoderKL(ddconfig, lossconfig, embed_dim, ckpt_path=None, ignore_keys=[], ...)``
    with the reference signature (autoencoder.py:285-311).  ``precision`` is the
    one extra knob: "mixed" (default: fp16 operands where they are bounded by construction, bf16 for the
    attention operands, every fp16 store range-checked), "bf16", "fp16" or "fp32" (check mode)."""

    def __init__(self, ddconfig=None, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=[],
                 image_key="image", colorize_nlabels=None, monitor=None, precision=None, chunk=None):
        super().__init__()
        ddconfig = dict(KL_F8_DDCONFIG if ddconfig is None else ddconfig)
        assert ddconfig["double_z"]
        for key in ("ch", "ch_mult", "num_res_blocks", "z_channels", "in_channels", "attn_resolutions"):
            if list(np.atleast_1d(ddconfig[key])) != list(np.atleast_1d(KL_F8_DDCONFIG[key])):
                raise ValueError(f"only the kl-f8 ddconfig is supported ({key}={ddconfig[key]!r})")
        if embed_dim != 4:
            raise ValueError("only embed_dim=4 (kl-f8) is supported")
        self.image_key = image_key
        self.embed_dim = embed_dim
        self.precision = (precision or os.environ.get("SFV_PRECISION", _lib.DEFAULT_PRECISION)).lower()
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.chunk = chunk or int(os.environ.get("SFV_CHUNK", "0")) or None
        _build_tree(self, encoder_param_shapes())
        self._handle = None
        self._handle_device = None
        self._ws = _lib.Workspace()
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    # -- weights -------------------------------------------------------------
    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        keys = list(sd.keys())
        for k in keys:
            for ik in ignore_keys:
                if k.startswith(ik):
                    print("Deleting key {} from state_dict.".format(k))
                    del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    def load_state_dict(self, state_dict, strict=True, **kw):
        # accept SD checkpoints that carry the autoencoder under "first_stage_model."
        if any(k.startswith("first_stage_model.") for k in state_dict):
            state_dict = {k[len("first_stage_model."):]: v for k, v in state_dict.items()
                          if k.startswith("first_stage_model.")}
        if not strict:   # decoder / post_quant_conv / loss keys are not part of this path
            own = set(self.state_dict().keys())
            state_dict = {k: v for k, v in state_dict.items() if k in own}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._release()
        return out

    def _release(self):
        if self._handle is not None:
            _lib.lib().sfv_encoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _native(self, device):
        """The native handle for `device` (weights are uploaded to the device that is current at creation;
        one handle per device)."""
        device = torch.device(device)
        if self._handle is not None and self._handle_device != device:
            self._release()
        if self._handle is None:
            table, n, keep = _lib.make_tensor_table(self.state_dict())
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.lib().sfv_encoder_create(table, n, _lib.PRECISIONS[self.precision], C.byref(h)))
            self._handle = h
            self._handle_device = device
            if self.chunk:
                _lib.check(_lib.lib().sfv_encoder_set_chunk(h, int(self.chunk)))
        return self._handle

    # -- the hot path --------------------------------------------------------
    def _run(self, x, u8: bool, taps=None):
        L = _lib.lib()
        if u8:
            B, H, W, _ = x.shape
        else:
            B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError(f"H and W must be multiples of 8, got {H}x{W}")
        dev = x.device
        h = self._native(dev)
        with torch.cuda.device(dev):
            return self._run_on(L, h, x, u8, taps, B, H, W, dev)

    def _run_on(self, L, h, x, u8, taps, B, H, W, dev):
        params = torch.empty(B, 8, H // 8, W // 8, dtype=torch.float32, device=dev)
        logvar = torch.empty(B, 4, H // 8, W // 8, dtype=torch.float32, device=dev)
        std = torch.empty_like(logvar)
        var = torch.empty_like(logvar)
        nbytes = C.c_size_t()
        _lib.check(L.sfv_encoder_workspace_bytes(h, B, H, W, C.byref(nbytes)))
        ws = self._ws.get(nbytes.value, dev)
        if u8:
            _lib.check(L.sfv_encoder_forward_u8(h, _lib.ptr(x), B, H, W, _lib.ptr(params), _lib.ptr(logvar),
                                                _lib.ptr(std), _lib.ptr(var), _lib.ptr(ws), nbytes.value,
                                                _lib.stream_ptr(dev)))
        else:
            tap_arr = None
            if taps is not None:
                tap_arr = (C.c_void_p * _lib.NUM_TAPS)()
                for i in range(_lib.NUM_TAPS):
                    tap_arr[i] = taps[i].data_ptr() if taps[i] is not None else None
            _lib.check(L.sfv_encoder_forward_nchw(h, _lib.ptr(x), B, H, W, _lib.ptr(params), _lib.ptr(logvar),
                                                  _lib.ptr(std), _lib.ptr(var), _lib.ptr(ws), nbytes.value,
                                                  tap_arr, _lib.stream_ptr(dev)))
        return DiagonalGaussianDistribution(params, _native=(logvar, std, var))

    @torch.no_grad()
    def encode(self, x):
        """autoencoder.py:324-328: x f32 [B,3,H,W] in [-1,1] -> posterior."""
        _lib.require_cuda(x, "AutoencoderKL.encode input")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected [B,3,H,W], got {tuple(x.shape)}")
        return self._run(x.to(torch.float32).contiguous(), u8=False)

    @torch.no_grad()
    def encode_uint8(self, frames):
        """uint8 RGB frames [B,H,W,3] (already at the target size) -> posterior;
        load_img's /255 and 2x-1 (get_percep_embeddings.py:67-71) are fused into conv_in."""
        _lib.require_cuda(frames, "AutoencoderKL.encode_uint8 input")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [B,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
        return self._run(frames.contiguous(), u8=True)

    @torch.no_grad()
    def encode_with_taps(self, x):
        """encode() that also returns every block output (fp32 NHWC) for layer-wise parity."""
        _lib.require_cuda(x, "input")
        B, _, H, W = x.shape
        taps = [torch.empty(B, H >> d, W >> d, c, dtype=torch.float32, device=x.device)
                for c, d in zip(_lib.TAP_CHANNELS, _lib.TAP_DOWN)]
        post = self._run(x.to(torch.float32).contiguous(), u8=False, taps=taps)
        return post, dict(zip(_lib.TAP_NAMES, taps))

    def check_async_error(self):
        """Synchronise and raise if a device-side watchdog / fp16 range check fired (see _lib.check_async_error)."""
        _lib.check_async_error(self._handle_device)

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("decoder is outside the accelerated path; use encode()")

    def decode(self, z):
        raise NotImplementedError("decoder is outside the accelerated path")


def _latent_dist(self):
    return self


# `posterior.latent_dist` aliases the posterior itself so both call styles work
DiagonalGaussianDistribution.latent_dist = property(_latent_dist)


class FirstStage:
    """The slice of ``LatentDiffusion`` this repo uses (ddpm.py:542-549, 825-863):
    ``encode_first_stage`` and ``get_first_stage_encoding`` with ``scale_factor``."""

    def __init__(self, first_stage_model: AutoencoderKL, scale_factor: float = SCALE_FACTOR):
        self.first_stage_model = first_stage_model
        self.scale_factor = scale_factor

    @torch.no_grad()
    def encode_first_stage(self, x):
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            if noise is None:
                noise = torch.randn(encoder_posterior.mean.shape)   # same global-RNG draw as sample()
            noise = noise.to(device=encoder_posterior.parameters.device, dtype=torch.float32).contiguous()
            return _scaled_sample(encoder_posterior, noise, self.scale_factor)
        elif isinstance(encoder_posterior, torch.Tensor):
            return self.scale_factor * encoder_posterior
        raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")

    def get_first_stage_mode(self, encoder_posterior):
        """scale_factor * posterior.mode(): the deterministic embedding used for parity runs."""
        return _scaled_sample(encoder_posterior, None, self.scale_factor)
