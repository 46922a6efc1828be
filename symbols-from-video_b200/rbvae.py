"""Drop-in host mirror of the reference's RBVAE: the encoder half (the precompute hot path) and, when the state-dict
carries it, the decoder half of ``forward`` (training-side forward; fp32, forward values only).

* percep      -- models/percep_RBVAE/percep_RBVAE_model.py:125-191
                 (256-channel convs, 4-layer LSTM, fc hard-wired to 256*11*20)
* contrastive -- models/contrastive_RBVAE/contrastive_RBVAE_model.py:124-190
                 (64-channel convs, 2-layer LSTM, fc hard-wired to 64*32*32)

``Seq2SeqBinaryVAE(in_channels, out_channels, latent_dim, hidden_dim)`` keeps
the reference constructor; ``kind`` / ``input_hw`` are the extra knobs that let
the fc layer follow the latent shape (the reference fails on anything but its
native 88x160 / 256x256 input).  State-dict key names are the
reference's, so ``load_state_dict(torch.load(...)['model_state_dict'])`` works
unchanged.  The decoder modules (``decoder_rnn``, ``decoder_cnn``) are created by ``load_state_dict`` when the
state-dict has their keys; without them ``forward`` returns ``x_recon = None``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_KINDS = {"percep": dict(channels=256, layers=4, input_hw=(88, 160)),
          "contrastive": dict(channels=64, layers=2, input_hw=(256, 256))}


def _down3(n):
    for _ in range(3):
        n = (n - 1) // 2 + 1
    return n


class _ConvStack(nn.Module):
    def __init__(self, cin, ch, latent_dim, fin):
        super().__init__()
        self.conv = nn.Module()
        for idx, ci in ((0, cin), (3, ch), (6, ch)):     # Sequential indices of the three Conv2d
            leaf = nn.Module()
            leaf.weight = nn.Parameter(torch.zeros(ch, ci, 3, 3), requires_grad=False)
            leaf.bias = nn.Parameter(torch.zeros(ch), requires_grad=False)
            self.conv.add_module(str(idx), leaf)
        self.fc = nn.Module()
        self.fc.weight = nn.Parameter(torch.zeros(latent_dim, fin), requires_grad=False)
        self.fc.bias = nn.Parameter(torch.zeros(latent_dim), requires_grad=False)


class _Lstm(nn.Module):
    def __init__(self, L, layers):
        super().__init__()
        self.lstm = nn.Module()
        for l in range(layers):
            for n, shape in (("weight_ih", (4 * L, L)), ("weight_hh", (4 * L, L)),
                             ("bias_ih", (4 * L,)), ("bias_hh", (4 * L,))):
                self.lstm.register_parameter(f"{n}_l{l}", nn.Parameter(torch.zeros(shape), requires_grad=False))


class _DeconvStack(nn.Module):
    """ConvDecoder's parameters (percep_RBVAE_model.py:71-85): fc(L -> ch*fh*fw), three ConvTranspose2d."""

    def __init__(self, cout, ch, latent_dim, fout):
        super().__init__()
        self.fc = nn.Module()
        self.fc.weight = nn.Parameter(torch.zeros(fout, latent_dim), requires_grad=False)
        self.fc.bias = nn.Parameter(torch.zeros(fout), requires_grad=False)
        self.deconv = nn.Module()
        for idx, co in ((0, ch), (3, ch), (6, cout)):     # Sequential indices of the three ConvTranspose2d
            leaf = nn.Module()
            leaf.weight = nn.Parameter(torch.zeros(ch, co, 3, 3), requires_grad=False)   # [Cin, Cout, 3, 3]
            leaf.bias = nn.Parameter(torch.zeros(co), requires_grad=False)
            self.deconv.add_module(str(idx), leaf)


class Seq2SeqBinaryVAE(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_dim=32, hidden_dim=32, kind=None, input_hw=None,
                 precision="fp32"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.precision = precision
        if hidden_dim != latent_dim:
            # the reference ignores hidden_dim too: EncoderRNN(latent_dim, hidden_dim=latent_dim) (:139)
            hidden_dim = latent_dim
        if kind is None:
            kind = "percep" if in_channels == 4 else "contrastive"
        if kind not in _KINDS:
            raise ValueError(f"kind must be one of {sorted(_KINDS)}")
        cfg = _KINDS[kind]
        self.kind = kind
        self.latent_dim = latent_dim
        self.in_channels = in_channels
        self.channels = cfg["channels"]
        self.input_hw = tuple(input_hw) if input_hw is not None else cfg["input_hw"]
        fin = self.channels * _down3(self.input_hw[0]) * _down3(self.input_hw[1])
        self.encoder_cnn = _ConvStack(in_channels, self.channels, latent_dim, fin)
        self.encoder_rnn = _Lstm(latent_dim, cfg["layers"])
        self.out_channels = out_channels
        self.decoder_cnn = None         # created by load_state_dict when the checkpoint carries the decoder
        self.decoder_rnn = None
        self._dec_handles = {}
        self._dec_ws = _lib.Workspace()
        self._handles = {}
        self._ws = _lib.Workspace()

    # -- weights -------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        if "decoder_cnn.fc.weight" in state_dict and "decoder_rnn.lstm.weight_ih_l0" in state_dict:
            fout = state_dict["decoder_cnn.fc.weight"].shape[0]
            layers = 0
            while f"decoder_rnn.lstm.weight_ih_l{layers}" in state_dict:
                layers += 1
            self.decoder_cnn = _DeconvStack(self.out_channels, self.channels, self.latent_dim, fout)
            self.decoder_rnn = _Lstm(self.latent_dim, layers)
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in state_dict.items() if k in own}
        missing = own - set(sd)
        if strict and missing:
            raise RuntimeError(f"missing keys: {sorted(missing)}")
        fcw = sd.get("encoder_cnn.fc.weight")
        if fcw is not None and tuple(fcw.shape) != tuple(self.encoder_cnn.fc.weight.shape):
            if fcw.shape[0] != self.latent_dim:
                raise RuntimeError(f"fc.weight has latent_dim {fcw.shape[0]}, model has {self.latent_dim}")
            self.encoder_cnn.fc.weight = nn.Parameter(torch.zeros_like(fcw), requires_grad=False)
        out = super().load_state_dict(sd, strict=False, **kw)
        self._release()
        return out

    def _release(self):
        for h in self._handles.values():
            _lib.lib().sfv_rbvae_destroy(h)
        self._handles = {}
        for h in self._dec_handles.values():
            _lib.lib().sfv_rbvae_decoder_destroy(h)
        self._dec_handles = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _native(self, H, W, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        key = (H, W, device)
        if key not in self._handles:
            fin = self.channels * _down3(H) * _down3(W)
            have = self.encoder_cnn.fc.weight.shape[1]
            if fin != have:
                # same failure the reference raises from nn.Linear on a mismatching flatten
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: conv features {fin} for a "
                                   f"{H}x{W} input vs fc.in_features {have}")
            table, n, keep = _lib.make_tensor_table(self.state_dict())
            h = C.c_void_p()
            with torch.cuda.device(device):       # weights land on `device`: one handle per (shape, device)
                _lib.check(_lib.lib().sfv_rbvae_create_ex(table, n, self.in_channels, H, W,
                                                          _lib.PRECISIONS[self.precision], C.byref(h)))
            self._handles[key] = h
        return self._handles[key]

    # -- the hot path --------------------------------------------------------
    @torch.no_grad()
    def _encode(self, x, temperature, hard, noise_ratio, U=None, in_scale=1.0, want_codes=False, out_codes=None,
                out_h=None):
        _lib.require_cuda(x, "Seq2SeqBinaryVAE input")
        if x.dim() != 5:
            raise ValueError(f"expected [B,T,C,H,W], got {tuple(x.shape)}")
        B, T, Cc, H, W = x.shape
        if Cc != self.in_channels:
            raise ValueError(f"expected {self.in_channels} channels, got {Cc}")
        L = self.latent_dim
        h = self._native(H, W, x.device)
        x = x.to(torch.float32).contiguous()
        dev = x.device
        if U is None:
            # same global-RNG CPU draw as binary_concrete_logits (percep_RBVAE_model.py:33); the reference
            # draws it even when noise_ratio == 0, so the generator state after the call matches too
            U = torch.rand(B * T, L)
            if noise_ratio == 0:
                U = None
        if U is not None:
            U = U.to(device=dev, dtype=torch.float32).reshape(B * T, L).contiguous()
        words = (L + 31) // 32
        for t, shape, dt, what in ((out_h, (B, T, L), torch.float32, "out_h"), (out_codes, (B * T, words), torch.int32, "out_codes")):
            if t is not None and (tuple(t.shape) != shape or t.dtype != dt or t.device != dev or not t.is_contiguous()):
                raise ValueError(f"{what} must be a contiguous {dt} tensor of shape {shape} on {dev}")
        h_seq = out_h if out_h is not None else torch.empty(B, T, L, dtype=torch.float32, device=dev)
        z_seq = torch.empty(B, T, L, dtype=torch.float32, device=dev)
        codes = out_codes if out_codes is not None else (
            torch.empty(B * T, words, dtype=torch.int32, device=dev) if want_codes else None)
        nbytes = C.c_size_t()
        lib = _lib.lib()
        _lib.check(lib.sfv_rbvae_workspace_bytes(h, B * T, C.byref(nbytes)))
        ws = self._ws.get(nbytes.value, dev)
        _lib.run(lib.sfv_rbvae_encode, x, h, _lib.ptr(x), B, T, float(in_scale), _lib.ptr(U), float(noise_ratio),
                 float(temperature), int(bool(hard)), _lib.ptr(h_seq), _lib.ptr(z_seq),
                 _lib.ptr(codes), _lib.ptr(ws), nbytes.value)
        return z_seq, h_seq, codes

    def check_async_error(self, device=None):
        """Synchronise and raise if a device-side watchdog / range check fired (the 16-bit modes run the two
        C->C convs on the tcgen05 kernel, whose pipeline waits are bounded)."""
        _lib.check_async_error(device)

    def encode(self, x, temperature=0.5, hard=False, noise_ratio=0.1, U=None):
        """percep_RBVAE_model.py:172-191: x [B,T,C,H,W] -> z_seq [B,T,L]."""
        return self._encode(x, temperature, hard, noise_ratio, U)[0]

    def encode_codes(self, x, temperature=0.5, noise_ratio=0.0, U=None, in_scale=1.0, out_codes=None, out_h=None):
        """hard code, bit-packed: (codes uint32-as-int32 [B*T, ceil(L/32)], h_seq [B,T,L]).  out_codes / out_h:
        optional destinations the LSTM+threshold+pack kernel writes directly (a rank's slice of a gather buffer)."""
        _, h_seq, codes = self._encode(x, temperature, True, noise_ratio, U, in_scale, want_codes=True,
                                       out_codes=out_codes, out_h=out_h)
        return codes, h_seq

    # -- decoder half (training-side forward) ---------------------------------
    def _native_decoder(self, H, W, device):
        key = (H, W, torch.device(device))
        if key not in self._dec_handles:
            if H % 8 or W % 8:
                raise ValueError(f"decoder output {H}x{W} must be a multiple of 8")
            fout = self.channels * (H // 8) * (W // 8)
            have = self.decoder_cnn.fc.weight.shape[0]
            if fout != have:
                # the reference reshapes fc's output to (ch, 11, 20) / (64, 32, 32) and fails likewise on other sizes
                raise RuntimeError(f"shape '[-1, {self.channels}, {H // 8}, {W // 8}]' is invalid for fc.out_features {have}")
            sd = {k: v for k, v in self.state_dict().items() if k.startswith("decoder_")}
            table, n, keep = _lib.make_tensor_table(sd)
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.lib().sfv_rbvae_decoder_create(table, n, self.out_channels, H, W, C.byref(h)))
            self._dec_handles[key] = h
        return self._dec_handles[key]

    @torch.no_grad()
    def decode(self, z_seq, hw, return_d=False):
        """decoder_rnn + decoder_cnn (percep_RBVAE_model.py:159-168): z_seq [B,T,L] -> x_recon [B,T,C,H,W], hw = (H, W)."""
        if self.decoder_cnn is None:
            raise RuntimeError("this model was loaded without decoder weights (decoder_cnn.* / decoder_rnn.*)")
        _lib.require_cuda(z_seq, "z_seq")
        B, T, L = z_seq.shape
        if L != self.latent_dim:
            raise ValueError(f"expected latent_dim {self.latent_dim}, got {L}")
        H, W = hw
        dev = z_seq.device
        h = self._native_decoder(H, W, dev)
        z = z_seq.to(torch.float32).contiguous()
        d_seq = torch.empty(B, T, L, dtype=torch.float32, device=dev)
        x_recon = torch.empty(B, T, self.out_channels, H, W, dtype=torch.float32, device=dev)
        nbytes = C.c_size_t()
        lib = _lib.lib()
        _lib.check(lib.sfv_rbvae_decoder_workspace_bytes(h, B * T, C.byref(nbytes)))
        ws = self._dec_ws.get(nbytes.value, dev)
        _lib.run(lib.sfv_rbvae_decode, z, h, _lib.ptr(z), B, T, _lib.ptr(d_seq), _lib.ptr(x_recon), _lib.ptr(ws), nbytes.value)
        return (x_recon, d_seq) if return_d else x_recon

    def forward(self, x, temperature=1.0, hard=False, noise_ratio=0.1, U=None):
        """percep_RBVAE_model.py:143-170: returns (x_recon, h_seq, z_seq).  x_recon is None when the model was loaded
        without decoder weights (the precompute path needs the encoder half only)."""
        z_seq, h_seq, _ = self._encode(x, temperature, hard, noise_ratio, U)
        x_recon = self.decode(z_seq, x.shape[-2:]) if self.decoder_cnn is not None else None
        return x_recon, h_seq, z_seq


def unpack_codes(codes: torch.Tensor, L: int) -> torch.Tensor:
    """uint32 words [N, ceil(L/32)] -> float {0,1} [N, L] (bit j of word w = latent 32w+j)."""
    c = codes.to(torch.int64) & 0xFFFFFFFF
    bits = (c.unsqueeze(-1) >> torch.arange(32, device=codes.device)) & 1
    return bits.reshape(codes.shape[0], -1)[:, :L].to(torch.float32)


def hamming_matrix(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Pairwise Hamming distance of packed codes (embedding_hamming_distance.py:53-87) on the device."""
    _lib.require_cuda(a, "codes")
    a = a.contiguous(); b = b.contiguous()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.int32, device=a.device)
    _lib.run(_lib.lib().sfv_hamming, a, _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], a.shape[1],
                                      _lib.ptr(out))
    return out
