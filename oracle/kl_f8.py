"""Oracle: functional fp32 restatement of the SD KL-f8 ``AutoencoderKL.encode``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function takes the
reference's own state-dict (key names of SURVEY.md section 8b) and mirrors the
reference forward line by line; citations are ``path:line`` under
/root/reference/src/stable-diffusion/.

Encoder hyper-parameters are the kl-f8 ddconfig
(configs/stable-diffusion/v1-inference.yaml:46-67): ch=128, ch_mult=(1,2,4,4),
num_res_blocks=2, attn_resolutions=[], z_channels=4, double_z, embed_dim=4.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CH = 128
CH_MULT = (1, 2, 4, 4)
NUM_RES_BLOCKS = 2
Z_CHANNELS = 4
EMBED_DIM = 4
GN_GROUPS = 32
GN_EPS = 1e-6
SCALE_FACTOR = 0.18215  # v1-inference.yaml:17

DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
                ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[],
                dropout=0.0)


def swish(x):
    """ldm/modules/diffusionmodules/model.py:33-35."""
    return x * torch.sigmoid(x)


def group_norm(x, sd, prefix):
    """model.py:38-39 -- GroupNorm(32, C, eps=1e-6, affine)."""
    return F.group_norm(x, GN_GROUPS, sd[prefix + ".weight"], sd[prefix + ".bias"], GN_EPS)


def conv(x, sd, prefix, stride=1, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], stride=stride, padding=padding)


def resnet_block(x, sd, prefix, taps=None):
    """model.py:121-141 (temb is None, dropout p=0)."""
    h = group_norm(x, sd, prefix + ".norm1")
    h = swish(h)
    h = conv(h, sd, prefix + ".conv1", 1, 1)
    if taps is not None:
        taps[prefix + ".conv1"] = h
    h = group_norm(h, sd, prefix + ".norm2")
    h = swish(h)
    h = conv(h, sd, prefix + ".conv2", 1, 1)
    if (prefix + ".nin_shortcut.weight") in sd:
        x = conv(x, sd, prefix + ".nin_shortcut", 1, 0)
    return x + h


def downsample(x, sd, prefix):
    """model.py:72-79 -- zero pad right/bottom only, then 3x3 stride-2 pad-0 conv."""
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return conv(x, sd, prefix + ".conv", 2, 0)


def attn_block(x, sd, prefix):
    """model.py:178-202 -- single-head spatial self-attention, head dim = C."""
    h_ = group_norm(x, sd, prefix + ".norm")
    q = conv(h_, sd, prefix + ".q")
    k = conv(h_, sd, prefix + ".k")
    v = conv(h_, sd, prefix + ".v")
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = conv(h_, sd, prefix + ".proj_out")
    return x + h_


def encoder_forward(x, sd, prefix="encoder.", taps=None):
    """model.py:434-459 ``Encoder.forward``.  ``taps`` (dict) collects the
    output of every block for layer-wise bisection in the parity tests."""
    p = prefix
    h = conv(x, sd, p + "conv_in", 1, 1)
    if taps is not None:
        taps["conv_in"] = h
    for lvl in range(len(CH_MULT)):
        for blk in range(NUM_RES_BLOCKS):
            name = f"down.{lvl}.block.{blk}"
            h = resnet_block(h, sd, p + name, taps)
            if taps is not None:
                taps[name] = h
        if lvl != len(CH_MULT) - 1:
            name = f"down.{lvl}.downsample"
            h = downsample(h, sd, p + name)
            if taps is not None:
                taps[name] = h
    h = resnet_block(h, sd, p + "mid.block_1", taps)
    if taps is not None:
        taps["mid.block_1"] = h
    h = attn_block(h, sd, p + "mid.attn_1")
    if taps is not None:
        taps["mid.attn_1"] = h
    h = resnet_block(h, sd, p + "mid.block_2", taps)
    if taps is not None:
        taps["mid.block_2"] = h
    h = group_norm(h, sd, p + "norm_out")
    h = swish(h)
    h = conv(h, sd, p + "conv_out", 1, 1)
    return h


def encode_moments(x, sd, taps=None):
    """ldm/models/autoencoder.py:324-326 -- encoder then quant_conv (1x1, 8->8)."""
    h = encoder_forward(x, sd, "encoder.", taps)
    return conv(h, sd, "quant_conv")


class Posterior:
    """ldm/modules/distributions/distributions.py:24-62."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean

    def kl(self):
        return 0.5 * torch.sum(self.mean ** 2 + self.var - 1.0 - self.logvar, dim=[1, 2, 3])


def encode(x, sd, taps=None):
    """autoencoder.py:324-328 ``AutoencoderKL.encode`` -> posterior."""
    with torch.no_grad():
        return Posterior(encode_moments(x, sd, taps))


def first_stage_encoding(posterior, noise=None, use_mode=False):
    """ldm/models/diffusion/ddpm.py:542-549: scale_factor * posterior.sample()."""
    z = posterior.mode() if use_mode else posterior.sample(noise)
    return SCALE_FACTOR * z


def strip_prefix(sd, prefix="first_stage_model."):
    """SD checkpoints carry the autoencoder under ``first_stage_model.``
    (get_percep_embeddings.py:34-39)."""
    if any(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return sd


def init_state_dict(seed=0, gain=1.0):
    """Seeded default-PyTorch-init weights with the reference key names
    (no pretrained checkpoint exists offline).  Builds plain
    nn.Conv2d / nn.GroupNorm modules so the init distribution is the one the
    reference's constructors would produce."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv_(name, ci, co, k):
        bound = 1.0 / (ci * k * k) ** 0.5  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        sd[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound * gain
        sd[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def norm_(name, c):
        # perturbed affine so that gamma/beta handling is actually exercised
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def res_(name, ci, co):
        norm_(name + ".norm1", ci)
        conv_(name + ".conv1", ci, co, 3)
        norm_(name + ".norm2", co)
        conv_(name + ".conv2", co, co, 3)
        if ci != co:
            conv_(name + ".nin_shortcut", ci, co, 1)

    conv_("encoder.conv_in", 3, CH, 3)
    cin = CH
    for lvl, m in enumerate(CH_MULT):
        cout = CH * m
        for blk in range(NUM_RES_BLOCKS):
            res_(f"encoder.down.{lvl}.block.{blk}", cin, cout)
            cin = cout
        if lvl != len(CH_MULT) - 1:
            conv_(f"encoder.down.{lvl}.downsample.conv", cin, cin, 3)
    res_("encoder.mid.block_1", cin, cin)
    norm_("encoder.mid.attn_1.norm", cin)
    for n in ("q", "k", "v", "proj_out"):
        conv_(f"encoder.mid.attn_1.{n}", cin, cin, 1)
    res_("encoder.mid.block_2", cin, cin)
    norm_("encoder.norm_out", cin)
    conv_("encoder.conv_out", cin, 2 * Z_CHANNELS, 3)
    conv_("quant_conv", 2 * Z_CHANNELS, 2 * EMBED_DIM, 1)
    return sd


def flops_per_frame(H, W):
    """2*MAC of encoder + quant_conv: conv = 2*Co*Ho*Wo*Ci*k^2,
    attention = 4*L^2*512."""
    f = 2 * 128 * H * W * 3 * 9
    cin = 128
    h, w = H, W
    for lvl, m in enumerate(CH_MULT):
        cout = 128 * m
        for blk in range(2):
            f += 2 * cout * h * w * cin * 9 + 2 * cout * h * w * cout * 9
            if cin != cout:
                f += 2 * cout * h * w * cin
            cin = cout
        if lvl != 3:
            h, w = h // 2, w // 2
            f += 2 * cin * h * w * cin * 9
    f += 2 * (2 * 512 * h * w * 512 * 9 * 2)           # mid.block_1, mid.block_2
    L = h * w
    f += 4 * 2 * 512 * L * 512 + 4 * L * L * 512       # q,k,v,proj + QK^T + PV
    f += 2 * 8 * L * 512 * 9 + 2 * 8 * L * 8           # conv_out + quant_conv
    return f
