"""CPU: the C-ABI library loads and exports every symbol include/sfv.h declares,
fails loudly without a GPU (no fallback), and the host-side mirror keeps the
reference's call surface (state-dict keys, signatures, npy format, sharding)."""
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import sfv_b200
from oracle import kl_f8, rbvae as orb

from conftest import GOLDEN, ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sfv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = sfv_b200.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libsfv.so does not export {s}"
    # and the ctypes table binds exactly the declared set
    assert sorted(sfv_b200._lib.SIGNATURES) == syms
    assert b"sm_100a" in lib.sfv_version()


def test_library_has_tcgen05_and_tma_sass():
    out = subprocess.run(["cuobjdump", "-sass", sfv_b200._lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in out.stdout and "UTMALDG" in out.stdout and "LDTM" in out.stdout
    assert "HMMA.16816" not in out.stdout        # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    assert sfv_b200.lib().sfv_device_ok() == 0
    vae = sfv_b200.AutoencoderKL()
    with pytest.raises(sfv_b200.SfvError):
        vae.encode(torch.zeros(1, 3, 64, 64))           # CPU tensor -> refuse
    table, n, keep = sfv_b200._lib.make_tensor_table(vae.state_dict())
    import ctypes as C
    h = C.c_void_p()
    st = sfv_b200.lib().sfv_encoder_create(table, n, 1, C.byref(h))
    assert st == -2 and b"no CPU fallback" in sfv_b200.lib().sfv_last_error()
    rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, 25, 25)
    with pytest.raises(sfv_b200.SfvError):
        rb.encode(torch.zeros(1, 1, 4, 88, 160))
    # decoder half and losses: same rule -- no CPU path, loud errors
    with pytest.raises(RuntimeError):
        rb.decode(torch.zeros(1, 1, 25), (88, 160))                      # loaded without decoder weights
    sd = orb.init_state_dict(4, 25, (11, 20), seed=0)
    sd.update(orb.init_decoder_state_dict(4, 25, (11, 20), seed=0))
    rb.load_state_dict(sd)
    with pytest.raises(sfv_b200.SfvError):
        rb.decode(torch.zeros(1, 1, 25), (88, 160))                      # CPU tensor -> refuse
    from sfv_b200 import losses
    a = torch.zeros(4, 8)
    for call in (lambda: losses.l1_loss(a, 0.1), lambda: losses.recon_loss(a, a), lambda: losses.triplet_loss(a, a, a),
                 lambda: losses.kl_binary_concrete(a), lambda: losses.contrast_loss(a, a, torch.zeros(4))):
        with pytest.raises(sfv_b200.SfvError):
            call()
    with pytest.raises(NotImplementedError):
        losses.triplet_loss(a, a, a, p=1.0)
    h = C.c_void_p()
    table, n, keep = sfv_b200._lib.make_tensor_table({k: v for k, v in sd.items() if k.startswith("decoder_")})
    assert sfv_b200.lib().sfv_rbvae_decoder_create(table, n, 4, 88, 160, C.byref(h)) == -2


def test_product_never_imports_oracle():
    code = "import sys, sfv_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    pkg = os.path.join(ROOT, "symbols-from-video_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#include\s+.*oracle)", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), f


def test_autoencoder_state_dict_keys_match_reference_names():
    vae = sfv_b200.AutoencoderKL(ddconfig=dict(kl_f8.DDCONFIG), lossconfig={"target": "torch.nn.Identity"}, embed_dim=4)
    ref_sd = kl_f8.init_state_dict(0)
    own = vae.state_dict()
    assert set(own) == set(ref_sd)
    for k in own:
        assert tuple(own[k].shape) == tuple(ref_sd[k].shape), k
    # SD checkpoints: prefixed keys + decoder keys, strict=False (get_percep_embeddings.py:34-39)
    ck = {"first_stage_model." + k: v for k, v in ref_sd.items()}
    ck["first_stage_model.decoder.conv_in.weight"] = torch.zeros(1)
    ck["model.diffusion_model.foo"] = torch.zeros(1)
    vae.load_state_dict(ck, strict=False)
    assert torch.equal(vae.state_dict()["encoder.conv_in.weight"], ref_sd["encoder.conv_in.weight"])
    with pytest.raises(ValueError):
        sfv_b200.AutoencoderKL(ddconfig=dict(kl_f8.DDCONFIG, ch=64))


def test_posterior_class_matches_reference_semantics_on_cpu():
    g = torch.Generator().manual_seed(0)
    params = torch.randn(2, 8, 4, 4, generator=g) * 20
    p = sfv_b200.DiagonalGaussianDistribution(params)
    o = kl_f8.Posterior(params)
    assert torch.equal(p.mean, o.mean) and torch.equal(p.logvar, o.logvar)
    assert torch.equal(p.std, o.std) and torch.equal(p.var, o.var)
    assert p.logvar.max() <= 20 and p.logvar.min() >= -30
    assert torch.allclose(p.kl(), o.kl())
    assert p.mode() is p.mean and p.latent_dist is p
    noise = torch.randn(p.mean.shape, generator=g)
    assert torch.allclose(p.sample(noise), o.sample(noise))
    torch.manual_seed(5); a = p.sample()
    torch.manual_seed(5); b = o.sample()
    assert torch.equal(a, b)                               # same global-RNG consumption as the reference


def test_rbvae_state_dict_keys_and_fc_resize():
    rb = sfv_b200.Seq2SeqBinaryVAE(in_channels=4, out_channels=4, latent_dim=25, hidden_dim=25)
    assert rb.encoder_cnn.fc.weight.shape == (25, 256 * 11 * 20)       # reference default, percep_RBVAE_model.py:61
    ref = orb.init_state_dict(4, 25, (11, 20), seed=0)
    assert set(ref) == set(rb.state_dict())
    assert rb.decoder_cnn is None and rb.decoder_rnn is None            # encoder half only until a checkpoint brings the rest
    # a checkpoint also carries the decoder (percep_RBVAE_train.py:697-702) -> the decoder modules appear, reference keys
    ck = dict(ref); ck.update(orb.init_decoder_state_dict(4, 25, (11, 20), seed=0))
    rb.load_state_dict(ck)
    assert torch.equal(rb.state_dict()["encoder_rnn.lstm.bias_hh_l3"], ref["encoder_rnn.lstm.bias_hh_l3"])
    assert set(rb.state_dict()) == set(ck)
    assert rb.decoder_cnn.fc.weight.shape == (256 * 11 * 20, 25)        # percep_RBVAE_model.py:74
    assert rb.decoder_cnn.deconv._modules["6"].weight.shape == (256, 4, 3, 3)     # ConvTranspose2d: [Cin, Cout, 3, 3]
    assert torch.equal(rb.state_dict()["decoder_rnn.lstm.weight_hh_l2"], ck["decoder_rnn.lstm.weight_hh_l2"])
    # the product's own seeded decoder init uses the same keys and shapes
    own = sfv_b200.init_rbvae_decoder_state_dict(4, 25, (11, 20), seed=0)
    assert {k: tuple(v.shape) for k, v in own.items()} == {k: tuple(v.shape) for k, v in ck.items() if k.startswith("decoder_")}
    # a truncated decoder is a strict-mode error, as with nn.Module.load_state_dict
    bad = dict(ref); bad["decoder_cnn.fc.weight"] = ck["decoder_cnn.fc.weight"]; bad["decoder_rnn.lstm.weight_ih_l0"] = ck["decoder_rnn.lstm.weight_ih_l0"]
    with pytest.raises(RuntimeError):
        sfv_b200.Seq2SeqBinaryVAE(4, 4, 25, 25).load_state_dict(bad)
    # square BASELINE shapes: fc follows the latent shape
    rb2 = sfv_b200.Seq2SeqBinaryVAE(4, 4, 25, 25, input_hw=(64, 64))
    assert rb2.encoder_cnn.fc.weight.shape == (25, 256 * 8 * 8)
    rb2.load_state_dict(orb.init_state_dict(4, 25, (4, 4), seed=0))    # adopts the checkpoint's fc shape
    assert rb2.encoder_cnn.fc.weight.shape == (25, 256 * 4 * 4)
    c = sfv_b200.Seq2SeqBinaryVAE(in_channels=3, out_channels=3, latent_dim=32, hidden_dim=32)
    assert c.kind == "contrastive" and c.encoder_cnn.fc.weight.shape == (32, 64 * 32 * 32)
    assert "encoder_rnn.lstm.weight_ih_l1" in c.state_dict() and "encoder_rnn.lstm.weight_ih_l2" not in c.state_dict()


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 64, 480, 1000):
        for w in (1, 2, 3, 4, 8):
            rs = [sfv_b200.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sfv_b200.shard_range(10, 2, 2)


def test_embedding_store_format(tmp_path):
    lat = np.random.default_rng(0).standard_normal((3, 4, 8, 8)).astype(np.float32)
    keys = [f"{i:010d}.jpg" for i in (0, 5, 9)]
    p = str(tmp_path / "chin_chess_perceps.npy")
    sfv_b200.save_embeddings_npy(p, keys, lat)
    d = np.load(p, allow_pickle=True).item()               # the reference's reader, percep_RBVAE_train.py:204
    assert sorted(d) == keys and d[keys[1]].shape == (1, 4, 8, 8) and d[keys[1]].dtype == np.float32
    assert np.array_equal(sfv_b200.lookup_embedding(d, 5)[0], lat[1])
    d2 = {k[:-4]: v for k, v in d.items()}                   # keys without ".jpg" also resolve (:337-350)
    assert np.array_equal(sfv_b200.lookup_embedding(d2, 9)[0], lat[2])
    with pytest.raises(KeyError):
        sfv_b200.lookup_embedding(d, 4)


def test_unpack_codes_roundtrip():
    z = (torch.rand(5, 70, generator=torch.Generator().manual_seed(0)) > 0.5).float()
    packed = torch.from_numpy(orb.pack_codes(z).astype(np.int64).astype(np.int32))
    assert torch.equal(sfv_b200.unpack_codes(packed, 70), z)


def test_all_gather_ragged_gloo_world2(tmp_path):
    """N>1 host path: contiguous ranges + ragged all-gather, gloo on CPU, world_size 2."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        import sfv_b200
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        N = 7
        lo, hi = sfv_b200.shard_range(N, r, w)
        full_codes = torch.arange(N * 2, dtype=torch.int32).reshape(N, 2)
        full_lat = torch.arange(N * 4 * 2 * 2, dtype=torch.float32).reshape(N, 4, 2, 2)
        counts = [sfv_b200.shard_range(N, i, w)[1] - sfv_b200.shard_range(N, i, w)[0] for i in range(w)]
        codes = sfv_b200.all_gather_ragged(full_codes[lo:hi].clone(), counts)
        lat = sfv_b200.all_gather_ragged(full_lat[lo:hi].clone(), counts)
        assert torch.equal(codes, full_codes) and torch.equal(lat, full_lat), (r, codes)
        # in-place gather of pre-filled blocks (what encode_sharded / precompute / bench do: the kernels write this
        # rank's block of the gather buffer, one collective per tensor, async handle), ragged tail padded to mx rows
        mx = max(counts)
        buf = torch.full((w * mx, 4, 2, 2), -1.0)
        buf[r * mx:r * mx + (hi - lo)] = full_lat[lo:hi]
        work = sfv_b200.all_gather_slices(buf, r, w, async_op=True)
        work.wait()
        got = torch.cat([buf[i * mx:i * mx + counts[i]] for i in range(w)])
        assert torch.equal(got, full_lat), (r, got)
        dist.destroy_process_group()
        print("rank", r, "ok", lo, hi)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_product_weight_and_frame_generators_match_the_oracles():
    """bench.py builds its models from the product's own seeded generators; the CPU-baseline / parity leg
    feeds the same tensors to the oracle, so both generators must agree bit for bit."""
    a, b = sfv_b200.init_encoder_state_dict(3), kl_f8.init_state_dict(3)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in b)
    a, b = sfv_b200.init_rbvae_state_dict(4, 25, (8, 8), seed=1), orb.init_state_dict(4, 25, (8, 8), seed=1)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in b)
    from oracle import frames
    assert np.array_equal(sfv_b200.synthetic_frames(2, 32, 40, 9, smooth=True).numpy(),
                          frames.synthetic_frames(2, 32, 40, 9, smooth=True))


def _toy_embeddings(n=160, hw=(2, 3), seed=9):
    # same construction as oracle/make_golden.py:dataset_embeddings (kept literal here: host test, no oracle import needed)
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(n, 4, *hw, generator=g).numpy()
    return {(f"{i:010d}.jpg" if i % 3 else f"{i:010d}"): lat[i:i + 1] for i in range(n)}


def test_flat_embedding_store_round_trip(tmp_path):
    emb = _toy_embeddings(20)
    flat = sfv_b200.FlatEmbeddingStore.from_pickled(emb)
    flat.save(tmp_path / "emb_flat")
    back = sfv_b200.FlatEmbeddingStore.load(tmp_path / "emb_flat")
    assert back.keys == list(emb.keys()) and len(back) == 20
    assert isinstance(back.latents, np.memmap)
    for i in (0, 1, 7, 19):
        assert np.array_equal(back.get(i), sfv_b200.lookup_embedding(emb, i))
    again = back.to_pickled()
    assert all(np.array_equal(again[k], emb[k]) and again[k].dtype == np.float32 and again[k].shape == (1, 4, 2, 3)
               for k in emb)
    # and through the reference's pickled format on disk
    sfv_b200.save_embeddings_npy(str(tmp_path / "emb.npy"), list(emb.keys()), np.concatenate(list(emb.values())))
    flat2 = sfv_b200.FlatEmbeddingStore.from_pickled(str(tmp_path / "emb.npy"))
    assert np.array_equal(np.asarray(flat2.latents), np.asarray(flat.latents))
    with pytest.raises(KeyError):
        back.get(20)


@pytest.mark.parametrize("mode", ["train", "val", "test"])
def test_resident_pair_dataset_matches_reference_golden(mode):
    """tests/golden/dataset.npz: the reference's ShuffledStatePairDataset under random.seed(31)."""
    import random
    g = np.load(os.path.join(GOLDEN, "dataset.npz"))
    segs = [tuple(int(v) for v in s) for s in g["segments"]]
    random.seed(31)
    ds = sfv_b200.ShuffledStatePairDataset(_toy_embeddings(), segs, test_pct=0.15, val_pct=0.1, mode=mode, device="cpu")
    assert random.random() == float(g[mode + "_rand_after"])          # consumed the RNG exactly like the reference
    assert len(ds) == g[mode + "_items"].shape[0]
    pairs = np.array([[list(p[i % len(p)]) for p in ds.pairs_per_state] for i in range(len(ds))])
    assert np.array_equal(pairs, g[mode + "_pairs"])
    items = torch.stack([ds[i] for i in range(len(ds))]).numpy()
    assert np.array_equal(items, g[mode + "_items"])
    assert np.array_equal(ds.batch(list(range(len(ds)))).numpy(), g[mode + "_items"])
    assert np.array_equal(ds._load_embedding(5).numpy(), _toy_embeddings()["0000000005.jpg"][0])
    with pytest.raises(ValueError):
        sfv_b200.ShuffledStatePairDataset(_toy_embeddings(), segs, mode="bogus", device="cpu")


def test_evaluation_host_logic():
    """assign_label / labels_from_flags (embedding_matching.py:195-206) and the occlusion geometry (:180)."""
    flags = [74, 206, 282, 389]                                   # transition_flags.txt: chinese_chess
    idx = [0, 73, 74, 75, 205, 206, 281, 282, 388, 389, 479]
    want = [0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 4]
    assert [sfv_b200.assign_label(i, flags) for i in idx] == want
    assert sfv_b200.labels_from_flags(idx, flags).tolist() == want
    assert sfv_b200.labels_from_flags([], flags).tolist() == []
    assert sfv_b200.evaluation.occlusion_size(432, 768, 0.2) == int(np.sqrt(0.2 * 432 * 768))
    # CPU tensors are refused by the device entry points (no CPU fallback)
    with pytest.raises(Exception):
        sfv_b200.state_consistency(torch.zeros(4, 1, dtype=torch.int32), torch.zeros(4, dtype=torch.int32), 2)
    with pytest.raises(Exception):
        sfv_b200.perturb_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))


def test_ctypes_signatures_match_header_arity():
    """Every prototype in include/sfv.h has as many parameters as its ctypes argtypes entry (ABI drift guard)."""
    hdr = open(os.path.join(ROOT, "include", "sfv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)                 # drop comments
    protos = dict(re.findall(r"\b(sfv_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S))
    assert set(protos) == set(sfv_b200._lib.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len([q for q in params.split(",") if q.strip()])
        assert n == len(sfv_b200._lib.SIGNATURES[name][1]), (name, n, len(sfv_b200._lib.SIGNATURES[name][1]))
