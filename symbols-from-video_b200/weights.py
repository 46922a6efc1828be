"""Seeded random-init state-dicts and synthetic frames with the reference's key names and shapes.

No pretrained checkpoint exists offline, so benchmarks and smoke runs use default-PyTorch-init
weights drawn from a seeded generator: Conv2d/Linear U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (what
kaiming_uniform_(a=sqrt(5)) gives), LSTM U(-1/sqrt(hidden), 1/sqrt(hidden)), GroupNorm affine slightly
perturbed so gamma/beta handling is exercised.  Real checkpoints load through the same
``load_state_dict`` / ``init_from_ckpt`` paths.
"""
from __future__ import annotations

import torch

from .autoencoder import encoder_param_shapes


def init_encoder_state_dict(seed: int = 0) -> dict:
    """KL-f8 encoder + quant_conv (108 tensors, reference key names)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv_(name, ci, co, k):
        bound = 1.0 / (ci * k * k) ** 0.5
        sd[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def norm_(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def res_(name, ci, co):
        norm_(name + ".norm1", ci); conv_(name + ".conv1", ci, co, 3)
        norm_(name + ".norm2", co); conv_(name + ".conv2", co, co, 3)
        if ci != co:
            conv_(name + ".nin_shortcut", ci, co, 1)

    conv_("encoder.conv_in", 3, 128, 3)
    cin = 128
    for lvl, m in enumerate((1, 2, 4, 4)):
        for blk in range(2):
            res_(f"encoder.down.{lvl}.block.{blk}", cin, 128 * m)
            cin = 128 * m
        if lvl != 3:
            conv_(f"encoder.down.{lvl}.downsample.conv", cin, cin, 3)
    res_("encoder.mid.block_1", cin, cin)
    norm_("encoder.mid.attn_1.norm", cin)
    for n in ("q", "k", "v", "proj_out"):
        conv_(f"encoder.mid.attn_1.{n}", cin, cin, 1)
    res_("encoder.mid.block_2", cin, cin)
    norm_("encoder.norm_out", cin)
    conv_("encoder.conv_out", cin, 8, 3)
    conv_("quant_conv", 8, 8, 1)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in encoder_param_shapes().items()}
    return sd


def init_rbvae_state_dict(in_channels, latent_dim, feat_hw, channels=256, num_layers=4, seed=0) -> dict:
    """RBVAE encoder half; ``fc`` sized for a (feat_h, feat_w) feature map."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def u(shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    ci = in_channels
    for idx in (0, 3, 6):
        b = 1.0 / (ci * 9) ** 0.5
        sd[f"encoder_cnn.conv.{idx}.weight"] = u((channels, ci, 3, 3), b)
        sd[f"encoder_cnn.conv.{idx}.bias"] = u((channels,), b)
        ci = channels
    fin = channels * feat_hw[0] * feat_hw[1]
    b = 1.0 / fin ** 0.5
    sd["encoder_cnn.fc.weight"] = u((latent_dim, fin), b)
    sd["encoder_cnn.fc.bias"] = u((latent_dim,), b)
    b = 1.0 / latent_dim ** 0.5
    for l in range(num_layers):
        sd[f"encoder_rnn.lstm.weight_ih_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"encoder_rnn.lstm.weight_hh_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"encoder_rnn.lstm.bias_ih_l{l}"] = u((4 * latent_dim,), b)
        sd[f"encoder_rnn.lstm.bias_hh_l{l}"] = u((4 * latent_dim,), b)
    return sd


def init_rbvae_decoder_state_dict(out_channels, latent_dim, feat_hw, channels=256, num_layers=4, seed=0) -> dict:
    """RBVAE decoder half (decoder_rnn + decoder_cnn, reference key names and layouts); ``fc`` sized for a
    (feat_h, feat_w) feature map = output / 8."""
    g = torch.Generator().manual_seed(seed + 7919)
    sd = {}

    def u(shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    fout = channels * feat_hw[0] * feat_hw[1]
    b = 1.0 / latent_dim ** 0.5
    sd["decoder_cnn.fc.weight"] = u((fout, latent_dim), b)
    sd["decoder_cnn.fc.bias"] = u((fout,), b)
    for idx, co in ((0, channels), (3, channels), (6, out_channels)):
        b = 1.0 / (co * 9) ** 0.5                      # ConvTranspose2d: fan_in is counted over weight.size(1) = Cout
        sd[f"decoder_cnn.deconv.{idx}.weight"] = u((channels, co, 3, 3), b)
        sd[f"decoder_cnn.deconv.{idx}.bias"] = u((co,), b)
    b = 1.0 / latent_dim ** 0.5
    for l in range(num_layers):
        sd[f"decoder_rnn.lstm.weight_ih_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"decoder_rnn.lstm.weight_hh_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"decoder_rnn.lstm.bias_ih_l{l}"] = u((4 * latent_dim,), b)
        sd[f"decoder_rnn.lstm.bias_hh_l{l}"] = u((4 * latent_dim,), b)
    return sd


def make_rbvae_responsive(sd: dict, fc_gain: float = 40.0, bias_gain: float = 0.02, ih_gain: float = 4.0) -> dict:
    """Default-init RBVAE weights give ONE constant code whatever the frame (the LSTM biases decide every sign;
    per-bit std of h across frames ~5e-6), so a code comparison on them proves nothing.  No trained checkpoint
    exists offline; this rescaling -- boost ``fc`` and the LSTM input weights, damp the fc / LSTM
    biases -- makes the code follow the frame (distinct codes per frame, a populated |h| < 1e-3 band), which is
    what the bench / smoke parity lines need.  Same recipe as the chinchess fixture (oracle/chinchess.py)."""
    out = dict(sd)
    out["encoder_cnn.fc.weight"] = sd["encoder_cnn.fc.weight"] * fc_gain
    for k in list(sd):
        if "bias" in k and ("lstm" in k or k.endswith("fc.bias")):
            out[k] = sd[k] * bias_gain
        if "weight_ih" in k:
            out[k] = sd[k] * ih_gain
    return out


def synthetic_frames(n, H, W, seed=1234, smooth=False) -> torch.Tensor:
    """SURVEY 8d synthetic inputs: i.i.d. uniform uint8 [n,H,W,3], or a low-pass-filtered variant
    (GroupNorm / softmax statistics on white noise are atypical)."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randint(0, 256, (n, H, W, 3), generator=g, dtype=torch.uint8)
    if smooth:
        import torch.nn.functional as F
        x = f.permute(0, 3, 1, 2).float()
        x = F.avg_pool2d(F.pad(x, (4, 4, 4, 4), mode="reflect"), 9, 1)
        x = (x - x.amin()) / (x.amax() - x.amin()) * 255.0
        f = x.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    return f
