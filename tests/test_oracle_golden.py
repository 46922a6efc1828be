"""CPU: the oracle (oracle/*.py) against the golden vectors minted from the
unmodified reference modules (oracle/make_golden.py), against independent numpy
restatements of the primitives, and -- when /root/reference is present -- against
the live reference classes."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import chinchess, evaluation as oev, frames, kl_f8, primitives_np, rbvae, ref_shim

from conftest import GOLDEN


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


ENC_CASES = ["kl_f8_seed0_2x64x96_white", "kl_f8_seed1_1x128x128_smooth", "kl_f8_seed0_2x256x256_white"]


@pytest.mark.parametrize("name", ENC_CASES[:2])
def test_encoder_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, H, W = [int(v) for v in g["shape"]]
    sd = kl_f8.init_state_dict(int(g["weight_seed"]))
    x = frames.normalise_u8(frames.synthetic_frames(B, H, W, int(g["frame_seed"]), bool(g["smooth"])))
    post = kl_f8.encode(x, sd)
    # same ATen kernels as the reference ran: agreement to fp32 round-off (bit-exact with equal thread counts)
    assert rel_l2(post.mean, g["mean"]) < 2e-6
    assert rel_l2(post.logvar, g["logvar"]) < 2e-6
    assert rel_l2(post.std, g["std"]) < 2e-6
    assert rel_l2(post.var, g["var"]) < 2e-6
    assert rel_l2(post.parameters, g["parameters"]) < 2e-6


RB_CASES = ["rbvae_percep_L25_32x32_T1", "rbvae_percep_L25_64x64_T1", "rbvae_percep_L100_88x160_T1",
            "rbvae_percep_L50_32x32_T4", "rbvae_contrastive_L25_256x256_T1"]


def rb_sd(g):
    hw = [int(v) for v in g["hw"]]
    fh, fw = hw
    for _ in range(3):
        fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
    return rbvae.init_state_dict(int(g["cin"]), int(g["L"]), (fh, fw), channels=int(g["ch"]),
                                 num_layers=int(g["layers"]), seed=int(g["seed"]))


@pytest.mark.parametrize("name", RB_CASES)
def test_rbvae_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = rb_sd(g)
    x = torch.from_numpy(g["x"])
    z, h = rbvae.encode(x, sd, temperature=0.5, hard=True, noise_ratio=0.0, return_h=True)
    assert np.abs(h.numpy() - g["h"]).max() < 2e-7          # explicit LSTM cell vs cuDNN-style fused nn.LSTM
    far = np.abs(g["h"]) > 1e-6
    assert np.array_equal(z.numpy()[far], g["z_hard"][far])
    # stochastic variants reproduce when the oracle is given the reference's uniform draw
    zn = rbvae.encode(x, sd, temperature=0.5, hard=True, noise_ratio=0.3, U=torch.from_numpy(g["U"]))
    noise = rbvae.logistic_noise(torch.from_numpy(g["U"]), 0.3).reshape(g["h"].shape).numpy()
    far = np.abs(g["h"] + noise) > 1e-5
    assert np.array_equal(zn.numpy()[far], g["z_noise"][far])
    zs = rbvae.encode(x, sd, temperature=0.7, hard=False, noise_ratio=0.1, U=torch.from_numpy(g["U_soft"]))
    assert np.abs(zs.numpy() - g["z_soft"]).max() < 1e-6


def test_t1_lstm_closed_form_is_single_step():
    """at T=1 from zero state the forget gate and W_hh drop out."""
    sd = rbvae.init_state_dict(4, 25, (4, 4), seed=9)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(5, 1, 25, generator=g)
    h = rbvae.lstm_forward(logits, sd)
    x = logits[:, 0]
    for l in range(4):
        gates = F.linear(x, sd[f"encoder_rnn.lstm.weight_ih_l{l}"], sd[f"encoder_rnn.lstm.bias_ih_l{l}"]) \
            + sd[f"encoder_rnn.lstm.bias_hh_l{l}"]
        i, f, gg, o = gates.chunk(4, 1)
        x = torch.sigmoid(o) * torch.tanh(torch.sigmoid(i) * torch.tanh(gg))
    assert torch.allclose(h[:, 0], x, atol=1e-7)


def test_pack_codes_layout():
    z = torch.zeros(2, 40); z[0, 0] = 1; z[0, 33] = 1; z[1, 31] = 1; z[1, 39] = 1
    p = rbvae.pack_codes(z)
    assert p.shape == (2, 2) and p.dtype == np.uint32
    assert p[0, 0] == 1 and p[0, 1] == 2 and p[1, 0] == 2 ** 31 and p[1, 1] == 2 ** 7


def test_resize_restatement_matches_pil_golden():
    g = np.load(os.path.join(GOLDEN, "resize_pil.npz"))
    fr = g["frame0"]
    out = frames.lanczos_resize_u8(frames.lanczos_resize_u8(fr, 720, 1280), 704, 1280)
    assert np.array_equal(out[::16], g["frame0_1280x704_rows"])
    assert int(out.astype(np.int64).sum()) == int(g["frame0_checksum"])
    assert np.array_equal(frames.lanczos_resize_u8(g["small"], 72, 128), g["small_up"])
    assert np.array_equal(frames.lanczos_resize_u8(g["small"], 24, 32), g["small_dn"])


def test_resize_restatement_matches_live_pil():
    from PIL import Image
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for (h, w) in [(64, 96), (16, 24), (37, 80), (50, 53)]:
        ref = np.array(Image.fromarray(img).resize((w, h), resample=Image.LANCZOS))
        assert np.array_equal(frames.lanczos_resize_u8(img, h, w), ref)


# ---- torch.nn.functional primitives vs explicit numpy restatements ----------
def test_conv2d_primitive():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 9, 11, generator=g); w = torch.randn(7, 5, 3, 3, generator=g); b = torch.randn(7, generator=g)
    for stride, pad in [(1, (1, 1, 1, 1)), (2, (0, 1, 0, 1)), (2, (1, 1, 1, 1))]:
        ref = primitives_np.conv2d(x.numpy(), w.numpy(), b.numpy(), stride, pad)
        out = F.conv2d(F.pad(x, pad), w, b, stride=stride)
        assert np.abs(out.numpy() - ref).max() < 1e-4
    # Downsample.forward == pad (0,1,0,1) + stride 2
    sd = {"d.conv.weight": w, "d.conv.bias": b}
    assert np.abs(kl_f8.downsample(x, sd, "d").numpy() - primitives_np.conv2d(x.numpy(), w.numpy(), b.numpy(), 2, (0, 1, 0, 1))).max() < 1e-4


def test_group_norm_and_attention_primitives():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 6, 5, generator=g) * 3 + 1
    ga = torch.randn(64, generator=g); be = torch.randn(64, generator=g)
    ref = primitives_np.group_norm(x.numpy(), ga.numpy(), be.numpy(), 32, 1e-6)
    out = kl_f8.group_norm(x, {"n.weight": ga, "n.bias": be}, "n")
    assert np.abs(out.numpy() - ref).max() < 1e-4
    assert np.abs(kl_f8.swish(x).numpy() - primitives_np.silu(x.numpy().astype(np.float64))).max() < 1e-5
    C = 64
    sd = {}
    for n in ("q", "k", "v", "proj_out"):
        sd[f"a.{n}.weight"] = torch.randn(C, C, 1, 1, generator=g) / 8
        sd[f"a.{n}.bias"] = torch.randn(C, generator=g) / 8
    sd["a.norm.weight"] = torch.ones(C); sd["a.norm.bias"] = torch.zeros(C)
    out = kl_f8.attn_block(x, sd, "a")
    hn = F.group_norm(x, 32, eps=1e-6)
    q, k, v = [F.conv2d(hn, sd[f"a.{n}.weight"], sd[f"a.{n}.bias"]).reshape(2, C, -1).numpy() for n in "qkv"]
    o = primitives_np.attention(q, k, v).reshape(2, C, 6, 5)
    ref = x.numpy() + F.conv2d(torch.from_numpy(o).float(), sd["a.proj_out.weight"], sd["a.proj_out.bias"]).numpy()
    assert np.abs(out.numpy() - ref).max() < 1e-4


def test_lstm_primitive():
    sd = rbvae.init_state_dict(4, 8, (2, 2), channels=16, num_layers=2, seed=2)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 5, 8, generator=g)
    out = rbvae.lstm_forward(x, sd).numpy()
    inp = x.numpy().astype(np.float64)
    for l in range(2):
        h = np.zeros((3, 8)); c = np.zeros((3, 8)); outs = []
        for t in range(5):
            h, c = primitives_np.lstm_cell(inp[:, t], h, c, *[sd[f"encoder_rnn.lstm.{n}_l{l}"].numpy().astype(np.float64)
                                                               for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")])
            outs.append(h)
        inp = np.stack(outs, 1)
    assert np.abs(out - inp).max() < 1e-6


def test_flop_count_matches_survey():
    assert abs(kl_f8.flops_per_frame(512, 512) / 1e9 - 1116.66) < 0.01
    assert abs(kl_f8.flops_per_frame(256, 256) / 1e9 - 272.72) < 0.01
    assert abs(kl_f8.flops_per_frame(1024, 1024) / 1e9 - 4878.95) < 0.01


# ---- live reference (build container only) -----------------------------------
@pytest.mark.skipif(not ref_shim.available(), reason="neither /root/reference nor oracle/_ref present")
def test_oracle_matches_live_reference_encoder():
    sd = kl_f8.init_state_dict(3)
    m = ref_shim.autoencoder_kl(sd)
    x = frames.normalise_u8(frames.synthetic_frames(1, 64, 64, seed=7))
    with torch.no_grad():
        p = m.encode(x)
    po = kl_f8.encode(x, sd)
    assert rel_l2(po.mean, p.mean) < 1e-6 and rel_l2(po.logvar, p.logvar) < 1e-6
    z = kl_f8.first_stage_encoding(po, noise=torch.zeros_like(po.mean))
    assert torch.allclose(z, 0.18215 * p.mode())


@pytest.mark.skipif(not ref_shim.available(), reason="neither /root/reference nor oracle/_ref present")
def test_oracle_matches_live_reference_rbvae():
    sd = rbvae.init_state_dict(4, 32, (11, 20), seed=11)
    m = ref_shim.rbvae("percep", 4, 32, sd)          # native hard-wired 88x160 shape, unmodified module
    x = torch.randn(2, 2, 4, 88, 160, generator=torch.Generator().manual_seed(0)) * 0.7
    with torch.no_grad():
        _, h_ref, z_ref = m(x, temperature=1.0, hard=True, noise_ratio=0.0)
    z, h = rbvae.encode(x, sd, temperature=1.0, hard=True, noise_ratio=0.0, return_h=True)
    assert (h - h_ref).abs().max() < 2e-7
    assert torch.equal(z, z_ref)


def test_chinchess_oracle_matches_reference_golden():
    """Sample-video fixture (reference classes on the 480 chinchess frames): the oracle on a slice of it."""
    g = np.load(os.path.join(GOLDEN, "chinchess_480x64x128.npz"))
    u8 = chinchess.frames_from_delta(g["frame_delta"])
    assert u8.shape == (480,) + chinchess.HW + (3,)
    sd = kl_f8.init_state_dict(int(g["weight_seed"]))
    rsd, _ = chinchess.rbvae_weights()
    idx = np.array([0, 1, 75, 207, 283, 390, 479])          # first two, one after each transition, last
    post = kl_f8.encode(frames.normalise_u8(u8[idx]), sd)
    lat = kl_f8.first_stage_encoding(post, use_mode=True)
    assert rel_l2(lat[:2], g["latent_first"]) < 2e-6
    assert rel_l2(lat[-1:], g["latent_last"]) < 2e-6
    np.testing.assert_allclose(lat.double().sum(dim=(1, 2, 3)).numpy(), g["latent_sum"][idx], atol=2e-4)
    z, h = rbvae.encode(lat[:, None], rsd, hard=True, noise_ratio=0.0, return_h=True)
    h = h[:, 0].numpy(); z = z[:, 0].numpy()
    np.testing.assert_allclose(h, g["h"][idx], atol=2e-6)
    band = np.abs(g["h"][idx]) < 1e-5
    assert ((z != g["z_hard"][idx]) & ~band).sum() == 0
    # the fixture itself: codes follow the frames, and some |h| sit inside the exemption band
    assert len(np.unique(g["z_hard"], axis=0)) > 8
    assert 0 < (np.abs(g["h"]) < 1e-3).sum() < g["h"].size // 4


def test_evaluation_oracle_matches_reference_golden():
    """State consistency + perturbations: oracle restatement vs the outputs of the reference's own
    functions (tests/golden/evaluation.npz, embedding_matching.py:141-297)."""
    g = np.load(os.path.join(GOLDEN, "evaluation.npz"))
    for i in range(g["imgs"].shape[0]):
        got = oev.gaussian_noise_u8(g["imgs"][i], torch.from_numpy(g["noise"][i]), float(g["gauss_mean"]), float(g["gauss_std"]))
        assert np.array_equal(got, g["gauss"][i])
        x, y = (int(v) for v in g["occ_xy"][i])
        assert np.array_equal(oev.occlusion_u8(g["imgs"][i], x, y, int(g["occ_size"])), g["occ"][i])
    w, pct = oev.state_consistency(g["z"][g["idx"]], g["labels"], len(g["flags"]) + 1)
    assert w == pytest.approx(float(g["weighted"]), abs=1e-12)
    np.testing.assert_allclose(pct, g["percentages"], atol=1e-12)


@pytest.mark.skipif(not ref_shim.live(), reason="needs the live /root/reference tree (scripts / training code)")
def test_evaluation_oracle_matches_live_reference():
    import random
    import torchvision.transforms as T
    from PIL import Image
    ns = ref_shim.embedding_matching_functions(flags=[4, 9])
    rng = np.random.default_rng(3)
    im = rng.integers(0, 256, (16, 24, 3), dtype=np.uint8)
    torch.manual_seed(11)
    ref = np.array(T.ToPILImage()(ns["add_gaussian_noise"](T.ToTensor()(Image.fromarray(im)), std=0.3)))
    torch.manual_seed(11)
    assert np.array_equal(oev.gaussian_noise_u8(im, torch.randn(1, 3, 16, 24), 0.0, 0.3), ref)
    random.seed(2)
    ref = np.array(T.ToPILImage()(ns["add_occlusion"](T.ToTensor()(Image.fromarray(im)), coverage=0.25)))
    random.seed(2)
    s = int(np.sqrt(0.25 * 16 * 24))
    x = random.randint(0, 24 - s); y = random.randint(0, 16 - s)
    assert np.array_equal(oev.occlusion_u8(im, x, y, s), ref)
    assert [ns["assign_label"](i, [4, 9]) for i in (0, 3, 4, 8, 9, 100)] == [0, 0, 1, 1, 2, 2]


@pytest.mark.skipif(not ref_shim.live(), reason="needs the live /root/reference tree (scripts / training code)")
def test_resident_pair_dataset_matches_live_reference():
    """sfv_b200.ShuffledStatePairDataset (HBM-resident mirror) vs the reference class, other seed/segments."""
    import random
    import sfv_b200
    from oracle.make_golden import dataset_embeddings
    cls = ref_shim.train_dataset_class()
    emb = dataset_embeddings(90, (3, 2), seed=4)
    segs = [(0, 20), (20, 21), (21, 64), (64, 90)]
    for mode in ("train", "test"):
        random.seed(77)
        ref = cls(emb, segs, test_pct=0.2, val_pct=0.2, mode=mode)
        ref_items = torch.stack([ref[i] for i in range(len(ref))]) if mode == "train" else None
        after = random.random()
        random.seed(77)
        if mode == "test":
            # the one-frame state has no test frames: the reference raises from __getitem__, so does the mirror
            ours = sfv_b200.ShuffledStatePairDataset(emb, segs, test_pct=0.2, val_pct=0.2, mode=mode, device="cpu")
            assert ours.pairs_per_state == ref.pairs_per_state
            with pytest.raises(ValueError):
                ref[0]
            with pytest.raises(ValueError):
                ours[0]
            continue
        ours = sfv_b200.ShuffledStatePairDataset(emb, segs, test_pct=0.2, val_pct=0.2, mode=mode, device="cpu")
        assert random.random() == after
        assert ours.pairs_per_state == ref.pairs_per_state
        assert torch.equal(torch.stack([ours[i] for i in range(len(ours))]), ref_items)
