"""Import the UNMODIFIED reference modules (TEST INFRASTRUCTURE): from the live /root/reference in the
build container, else from the byte-for-byte copies ``oracle/build_ref.py`` placed under ``oracle/_ref/``
(git-ignored, shipped to the GPU box by gpurun).  Used by ``oracle/make_golden.py``, by the CPU tests that pin
the oracle against the reference, and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg as the timed
CPU implementation (``kind: "reference"``).  The evaluation / dataset helpers further down need files that are
not part of the copied set and only work against the live tree.

``ldm.models.autoencoder`` imports ``pytorch_lightning`` and ``taming`` at module
import time (autoencoder.py:2,6); neither is installed and neither is touched by
``AutoencoderKL.encode``, so tiny stand-ins are registered in ``sys.modules``.
No reference source is copied or modified.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from . import build_ref

LIVE_ROOT = "/root/reference"
REF_ROOT = build_ref.root() or LIVE_ROOT


def available() -> bool:
    """The model classes (AutoencoderKL, both Seq2SeqBinaryVAE) can be imported."""
    return os.path.isdir(os.path.join(REF_ROOT, "src", "stable-diffusion", "ldm"))


def live() -> bool:
    """The whole reference tree is present (scripts, training code): build container only."""
    return os.path.isdir(os.path.join(LIVE_ROOT, "scripts"))


def video_path():
    p = os.path.join(REF_ROOT, "videos", "chinchess_gettyimages-148739276-640_adpp.mp4")
    return p if os.path.exists(p) else None


def _install_import_stubs():
    import torch.nn as nn
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = pl
    if "taming" not in sys.modules:
        names = ["taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"]
        mods = [types.ModuleType(n) for n in names]
        for n, m in zip(names, mods):
            sys.modules[n] = m
        mods[-1].VectorQuantizer2 = type("VectorQuantizer2", (nn.Module,), {})


def autoencoder_kl(sd=None):
    """The reference ``AutoencoderKL`` with the kl-f8 ddconfig, eval mode."""
    _install_import_stubs()
    p = os.path.join(REF_ROOT, "src", "stable-diffusion")
    if p not in sys.path:
        sys.path.insert(0, p)
    from ldm.models.autoencoder import AutoencoderKL  # noqa: E402
    from . import kl_f8
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = AutoencoderKL(ddconfig=dict(kl_f8.DDCONFIG), lossconfig={"target": "torch.nn.Identity"},
                          embed_dim=4)
    if sd is not None:
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(k.startswith(("decoder.", "post_quant_conv.")) for k in missing), missing
    return m.eval()


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def percep_module():
    return _load_file("ref_percep_RBVAE_model",
                      os.path.join(REF_ROOT, "models/percep_RBVAE/percep_RBVAE_model.py"))


def contrastive_module():
    return _load_file("ref_contrastive_RBVAE_model",
                      os.path.join(REF_ROOT, "models/contrastive_RBVAE/contrastive_RBVAE_model.py"))


def rbvae(kind, in_channels, latent_dim, sd=None, feat_hw=None):
    """Reference ``Seq2SeqBinaryVAE``; if ``feat_hw`` differs from the hard-wired
    fc size the fc *module attribute* is replaced on the instance
    (the reference file is untouched) so square BASELINE shapes can run."""
    import torch.nn as nn
    mod = percep_module() if kind == "percep" else contrastive_module()
    m = mod.Seq2SeqBinaryVAE(in_channels=in_channels, out_channels=in_channels,
                             latent_dim=latent_dim, hidden_dim=latent_dim)
    if feat_hw is not None:
        ch = m.encoder_cnn.conv[0].out_channels
        fin = ch * feat_hw[0] * feat_hw[1]
        if m.encoder_cnn.fc.in_features != fin:
            m.encoder_cnn.fc = nn.Linear(fin, latent_dim)
    if sd is not None:
        own = m.state_dict()
        own.update({k: v for k, v in sd.items()})
        m.load_state_dict(own)
    return m.eval()


def embedding_matching_functions(flags=()):
    """The reference's evaluation helpers, UNMODIFIED, executed from their own source text:
    scripts/evaluation/state_consistency_eval/embedding_matching.py imports omegaconf / ldm at module level
    (absent here), so the four function definitions are cut out of the file with ``ast`` and exec'd in a
    namespace holding exactly the globals they use (torch, np, random, torchvision.transforms as T, and the
    module-level ``flags`` that __main__ sets, :389)."""
    import ast
    import random
    import numpy as np
    import torch
    import torchvision.transforms as T
    path = os.path.join(LIVE_ROOT, "scripts/evaluation/state_consistency_eval/embedding_matching.py")
    src = open(path).read()
    want = {"add_gaussian_noise", "add_occlusion", "assign_label", "calculate_state_consistency"}
    ns = {"torch": torch, "np": np, "random": random, "T": T, "flags": list(flags)}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def train_dataset_class():
    """The reference's ``ShuffledStatePairDataset`` (models/percep_RBVAE/percep_RBVAE_train.py:181-360),
    UNMODIFIED, cut out of the training script with ``ast`` (the script runs a wandb sweep at import)."""
    import ast
    import random
    from pathlib import Path
    import numpy as np
    import torch
    from torch.utils.data import Dataset
    path = os.path.join(LIVE_ROOT, "models/percep_RBVAE/percep_RBVAE_train.py")
    ns = {"torch": torch, "np": np, "random": random, "Path": Path, "Dataset": Dataset}
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.ClassDef) and node.name == "ShuffledStatePairDataset":
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["ShuffledStatePairDataset"]


def train_loss_functions():
    """The reference's loss functions, UNMODIFIED, executed from their own source text:
    models/percep_RBVAE/percep_RBVAE_train.py imports wandb / torchvision datasets at module level, so the five
    function definitions (:27-107) are cut out with ``ast`` and exec'd with the globals they use (torch, F, np).
    Live tree only."""
    import ast
    import numpy as np
    import torch
    import torch.nn.functional as F
    path = os.path.join(LIVE_ROOT, "models/percep_RBVAE/percep_RBVAE_train.py")
    src = open(path).read()
    want = {"l1_loss", "recon_loss", "triplet_loss", "kl_binary_concrete", "contrast_loss"}
    ns = {"torch": torch, "F": F, "np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    assert want <= set(ns), sorted(want - set(ns))
    return ns
