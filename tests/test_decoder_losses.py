"""CPU: the oracle's decoder half and training losses (oracle/rbvae.py decode / forward, oracle/losses.py) against the
golden vectors minted from the UNMODIFIED reference (oracle/make_golden.py --only-forward) and, when /root/reference is
present, against the live reference functions.  SURVEY 8 f4 (training-side forward)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses, rbvae, ref_shim

from conftest import GOLDEN

FWD_CASES = ["rbvae_forward_percep_L25_88x160", "rbvae_forward_contrastive_L25_256x256"]


def forward_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    hw = [int(v) for v in g["hw"]]
    fh, fw = hw
    for _ in range(3):
        fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
    kw = dict(channels=int(g["ch"]), num_layers=int(g["layers"]), seed=int(g["seed"]))
    sd = rbvae.init_state_dict(int(g["cin"]), int(g["L"]), (fh, fw), **kw)
    sd.update(rbvae.init_decoder_state_dict(int(g["cin"]), int(g["L"]), (fh, fw), **kw))
    return g, sd, (fh, fw)


@pytest.mark.parametrize("name", FWD_CASES)
def test_oracle_forward_matches_reference_golden(name):
    g, sd, feat = forward_case(name)
    x = torch.from_numpy(g["x"])
    for tag, hard, temp, nr in (("soft", False, 1.0, 0.1), ("hard", True, 0.5, 0.0)):
        x_recon, h_seq, z_seq = rbvae.forward(x, sd, temperature=temp, hard=hard, noise_ratio=nr, U=torch.from_numpy(g[f"U_{tag}"]))
        assert np.abs(h_seq.numpy() - g[f"h_{tag}"]).max() < 1e-6
        assert np.abs(z_seq.numpy() - g[f"z_{tag}"]).max() < 1e-6
        assert np.abs(x_recon.numpy() - g[f"x_recon_{tag}"]).max() < 1e-6
        # the decoder on the reference's own z_seq: d_seq too
        xr, d = rbvae.decode(torch.from_numpy(g[f"z_{tag}"]), sd, feat)
        assert np.abs(d.numpy() - g[f"d_{tag}"]).max() < 1e-6
        assert np.abs(xr.numpy() - g[f"x_recon_{tag}"]).max() < 1e-6
        assert xr.shape == x.shape


def loss_inputs():
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    t = {k: torch.from_numpy(g[k]) for k in ("q", "a", "p", "n", "label", "xr", "x")}
    return g, t


def oracle_losses(t):
    return dict(l1=olosses.l1_loss(t["q"], 0.01), mse=olosses.recon_loss(t["xr"], t["x"]),
                triplet_swap=olosses.triplet_loss(t["a"], t["p"], t["n"]),
                triplet_noswap=olosses.triplet_loss(t["a"], t["p"], t["n"], margin=0.5, swap=False),
                kl_half=olosses.kl_binary_concrete(t["q"]), kl_p03=olosses.kl_binary_concrete(t["q"], p=0.3),
                contrast_euclid=olosses.contrast_loss(t["a"], t["p"], t["label"]),
                contrast_cos=olosses.contrast_loss(t["a"], t["p"], t["label"], margin=0.7, dist="cosine"))


def test_oracle_losses_match_reference_golden():
    g, t = loss_inputs()
    for k, v in oracle_losses(t).items():
        assert abs(float(v) - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))


@pytest.mark.skipif(not ref_shim.live(), reason="needs the live reference tree (build container only)")
def test_oracle_losses_match_live_reference_functions():
    fn = ref_shim.train_loss_functions()
    g = torch.Generator().manual_seed(7)
    for N, D in ((1, 1), (5, 25), (64, 100)):
        q = torch.randn(N, D, generator=g) * 5
        a, p, n = (torch.randn(N, D, generator=g) for _ in range(3))
        a[0] = p[0]                                   # zero distance: the eps inside pairwise_distance
        label = (torch.rand(N, generator=g) > 0.5).float()
        pairs = [(fn["l1_loss"](q, 0.3), olosses.l1_loss(q, 0.3)),
                 (fn["recon_loss"](a, p), olosses.recon_loss(a, p)),
                 (fn["triplet_loss"](a, p, n), olosses.triplet_loss(a, p, n)),
                 (fn["triplet_loss"](a, p, n, margin=2.0, swap=False), olosses.triplet_loss(a, p, n, margin=2.0, swap=False)),
                 (fn["kl_binary_concrete"](q, p=0.2), olosses.kl_binary_concrete(q, p=0.2)),
                 (fn["contrast_loss"](a, p, label), olosses.contrast_loss(a, p, label)),
                 (fn["contrast_loss"](a, p, label, dist="cosine"), olosses.contrast_loss(a, p, label, dist="cosine"))]
        for i, (r, o) in enumerate(pairs):
            assert abs(float(r) - float(o)) <= 2e-6 * max(1.0, abs(float(r))), (N, D, i, float(r), float(o))


@pytest.mark.skipif(not ref_shim.available(), reason="needs the reference model files")
def test_oracle_decoder_matches_live_reference_module():
    """Fresh weights and inputs (not the golden's): the unmodified Seq2SeqBinaryVAE.forward at its native shape."""
    L, hw, feat = 16, (88, 160), (11, 20)
    sd = rbvae.init_state_dict(4, L, feat, seed=31)
    sd.update(rbvae.init_decoder_state_dict(4, L, feat, seed=31))
    m = ref_shim.rbvae("percep", 4, L, sd)
    x = torch.randn(2, 3, 4, *hw, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(99)
    with torch.no_grad():
        xr, h, z = m(x, temperature=0.8, hard=False, noise_ratio=0.2)
    torch.manual_seed(99)
    U = torch.rand(6, L)
    xo, ho, zo = rbvae.forward(x, sd, temperature=0.8, hard=False, noise_ratio=0.2, U=U)
    assert np.abs(xo.numpy() - xr.numpy()).max() < 1e-6
    assert np.abs(ho.numpy() - h.numpy()).max() < 1e-6 and np.abs(zo.numpy() - z.numpy()).max() < 1e-6
