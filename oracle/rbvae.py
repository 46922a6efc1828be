"""Oracle: functional fp32 restatement of the RBVAE encoder path and (training-side forward) of its
decoder half.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Citations are ``path:line``
under /root/reference/.  Two families share one code path:

* percep RBVAE   -- models/percep_RBVAE/percep_RBVAE_model.py
    conv(Cin->256,s2,p1)+ReLU, conv(256->256,s2,p1)+ReLU, conv(256->256,s2,p1),
    flatten(C,H,W), fc, 4-layer LSTM, binary-concrete threshold     (:46-68,94-107,172-191)
* contrastive RBVAE -- models/contrastive_RBVAE/contrastive_RBVAE_model.py
    same with 64 channels and a 2-layer LSTM                        (:45-67,93-106,171-190)

The number of conv channels and LSTM layers is read off the state-dict, so the
same functions serve both.  Dropout is identity (eval mode).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def conv_encoder(x, sd, prefix="encoder_cnn."):
    """ConvEncoder.forward, percep_RBVAE_model.py:63-68.  x: [N,C,H,W] -> [N,L]."""
    h = F.relu(F.conv2d(x, sd[prefix + "conv.0.weight"], sd[prefix + "conv.0.bias"], stride=2, padding=1))
    h = F.relu(F.conv2d(h, sd[prefix + "conv.3.weight"], sd[prefix + "conv.3.bias"], stride=2, padding=1))
    h = F.conv2d(h, sd[prefix + "conv.6.weight"], sd[prefix + "conv.6.bias"], stride=2, padding=1)
    h = h.flatten(1)  # nn.Flatten: (C,H,W) order
    return F.linear(h, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])


def lstm_num_layers(sd, prefix="encoder_rnn.lstm."):
    n = 0
    while (prefix + f"weight_ih_l{n}") in sd:
        n += 1
    return n


def lstm_forward(x_seq, sd, prefix="encoder_rnn.lstm."):
    """nn.LSTM(batch_first=True) from zero state, percep_RBVAE_model.py:100-105.
    PyTorch's published cell (gate order i,f,g,o):
        gates = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh
        c_t = sigmoid(f)*c_{t-1} + sigmoid(i)*tanh(g);  h_t = sigmoid(o)*tanh(c_t)
    x_seq: [B,T,L] -> h_seq [B,T,L] of the top layer."""
    B, T, _ = x_seq.shape
    layer_in = x_seq
    for l in range(lstm_num_layers(sd, prefix)):
        w_ih, w_hh = sd[prefix + f"weight_ih_l{l}"], sd[prefix + f"weight_hh_l{l}"]
        b_ih, b_hh = sd[prefix + f"bias_ih_l{l}"], sd[prefix + f"bias_hh_l{l}"]
        Hd = w_hh.shape[1]
        h = torch.zeros(B, Hd, dtype=x_seq.dtype, device=x_seq.device)
        c = torch.zeros(B, Hd, dtype=x_seq.dtype, device=x_seq.device)
        outs = []
        for t in range(T):
            gates = F.linear(layer_in[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        layer_in = torch.stack(outs, dim=1)
    return layer_in


def logistic_noise(U, noise_ratio, eps=1e-8):
    """percep_RBVAE_model.py:35-37."""
    return noise_ratio * (torch.log(U + eps) - torch.log(1.0 - U + eps))


def binary_concrete(logits, temperature=0.5, hard=False, noise_ratio=0.1, U=None, eps=1e-8):
    """percep_RBVAE_model.py:17-44.  ``U`` replaces the reference's global-RNG
    ``torch.rand`` draw so that both sides can consume the same uniform tensor;
    with noise_ratio == 0 the result does not depend on U."""
    if U is None:
        U = torch.rand(logits.shape)                 # the reference draws on the CPU (:33)
    U = U.to(logits.device)
    y = torch.sigmoid((logits + logistic_noise(U, noise_ratio, eps)) / temperature)
    if hard:
        y_hard = (y > 0.5).float()
        y = (y_hard - y) + y
    return y


def encode(x, sd, temperature=0.5, hard=False, noise_ratio=0.1, U=None, return_h=False):
    """Seq2SeqBinaryVAE.encode, percep_RBVAE_model.py:172-191.
    x: [B,T,C,H,W] -> z_seq [B,T,L] (and the LSTM hidden ``h_seq`` if asked)."""
    with torch.no_grad():
        B, T, C, H, W = x.shape
        logits = conv_encoder(x.reshape(B * T, C, H, W), sd)
        L = logits.shape[1]
        h_seq = lstm_forward(logits.reshape(B, T, L), sd)
        z = binary_concrete(h_seq.reshape(B * T, L), temperature, hard, noise_ratio,
                            None if U is None else U.reshape(B * T, L))
        z_seq = z.reshape(B, T, L)
    return (z_seq, h_seq) if return_h else z_seq


def decode(z_seq, sd, feat_hw):
    """Second half of Seq2SeqBinaryVAE.forward, percep_RBVAE_model.py:159-168:
    d_seq = decoder_rnn(z_seq); x_recon = decoder_cnn(d_seq) = sigmoid(deconv3(relu(deconv2(relu(deconv1(fc(d)))))));
    each deconv is ConvTranspose2d(3, stride 2, padding 1, output_padding 1) (:75-84), dropout = identity (eval).
    z_seq: [B,T,L] -> (x_recon [B,T,C,H,W], d_seq [B,T,L])."""
    with torch.no_grad():
        B, T, L = z_seq.shape
        d_seq = lstm_forward(z_seq, sd, prefix="decoder_rnn.lstm.")
        p = "decoder_cnn.deconv."
        ch = sd[p + "0.weight"].shape[0]
        h = F.linear(d_seq.reshape(B * T, L), sd["decoder_cnn.fc.weight"], sd["decoder_cnn.fc.bias"])
        h = h.reshape(B * T, ch, feat_hw[0], feat_hw[1])      # the reference hard-wires (256, 11, 20) / (64, 32, 32)
        h = F.relu(F.conv_transpose2d(h, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1, output_padding=1))
        h = F.relu(F.conv_transpose2d(h, sd[p + "3.weight"], sd[p + "3.bias"], stride=2, padding=1, output_padding=1))
        h = torch.sigmoid(F.conv_transpose2d(h, sd[p + "6.weight"], sd[p + "6.bias"], stride=2, padding=1, output_padding=1))
        x_recon = h.reshape(B, T, h.shape[1], h.shape[2], h.shape[3])
    return x_recon, d_seq


def forward(x, sd, temperature=1.0, hard=False, noise_ratio=0.1, U=None):
    """Seq2SeqBinaryVAE.forward, percep_RBVAE_model.py:143-170 -> (x_recon, h_seq, z_seq)."""
    z_seq, h_seq = encode(x, sd, temperature, hard, noise_ratio, U, return_h=True)
    fh, fw = x.shape[-2], x.shape[-1]
    for _ in range(3):
        fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
    x_recon, _ = decode(z_seq, sd, (fh, fw))
    return x_recon, h_seq, z_seq


def init_decoder_state_dict(out_channels, latent_dim, feat_hw, channels=256, num_layers=4, seed=0):
    """Seeded default-init decoder weights with the reference key names (ConvTranspose2d weights are [Cin,Cout,3,3])."""
    g = torch.Generator().manual_seed(seed + 7919)
    sd = {}

    def u(shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    fout = channels * feat_hw[0] * feat_hw[1]
    b = 1.0 / latent_dim ** 0.5
    sd["decoder_cnn.fc.weight"] = u((fout, latent_dim), b)
    sd["decoder_cnn.fc.bias"] = u((fout,), b)
    for idx, co in ((0, channels), (3, channels), (6, out_channels)):
        b = 1.0 / (co * 9) ** 0.5
        sd[f"decoder_cnn.deconv.{idx}.weight"] = u((channels, co, 3, 3), b)
        sd[f"decoder_cnn.deconv.{idx}.bias"] = u((co,), b)
    b = 1.0 / latent_dim ** 0.5
    for l in range(num_layers):
        sd[f"decoder_rnn.lstm.weight_ih_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"decoder_rnn.lstm.weight_hh_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"decoder_rnn.lstm.bias_ih_l{l}"] = u((4 * latent_dim,), b)
        sd[f"decoder_rnn.lstm.bias_hh_l{l}"] = u((4 * latent_dim,), b)
    return sd


def pack_codes(z):
    """Bit-pack a {0,1} code [..., L] into uint32 words [..., ceil(L/32)],
    bit j of word w = z[..., 32*w + j] (the layout the CUDA threshold kernel emits)."""
    import numpy as np
    z = z.detach().cpu().numpy() > 0.5
    L = z.shape[-1]
    nw = (L + 31) // 32
    pad = np.zeros(z.shape[:-1] + (nw * 32 - L,), dtype=bool)
    bits = np.concatenate([z, pad], axis=-1).reshape(z.shape[:-1] + (nw, 32))
    weights = (1 << np.arange(32, dtype=np.uint64))
    return (bits.astype(np.uint64) * weights).sum(-1).astype(np.uint32)


def init_state_dict(in_channels, latent_dim, feat_hw, channels=256, num_layers=4, seed=0,
                    lstm_gain=1.0):
    """Seeded default-init weights with the reference key names; ``fc`` is
    sized from the latent shape (the reference hard-wires 256*11*20).
    feat_hw = (ceil(h/8), ceil(w/8))."""
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
    b = lstm_gain / latent_dim ** 0.5
    for l in range(num_layers):
        sd[f"encoder_rnn.lstm.weight_ih_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"encoder_rnn.lstm.weight_hh_l{l}"] = u((4 * latent_dim, latent_dim), b)
        sd[f"encoder_rnn.lstm.bias_ih_l{l}"] = u((4 * latent_dim,), b)
        sd[f"encoder_rnn.lstm.bias_hh_l{l}"] = u((4 * latent_dim,), b)
    return sd
