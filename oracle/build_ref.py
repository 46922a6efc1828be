"""Recipe that places the UNMODIFIED reference files of the hot path under ``oracle/_ref/`` (TEST INFRASTRUCTURE).

``oracle/_ref/`` is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored, so the
copies travel to the GPU box together with the built ``libsfv.so``.  They are what ``bench.py --impl reference``
times there (``cpu_baseline.kind = "reference"``) and what the CPU tests pin the oracle port against; nothing in
the product package reads them.  Run from ``__graft_entry__.build()`` whenever /root/reference is present; on the
GPU box (no /root/reference) the prebuilt tree is used as it is.

The list is SURVEY.md section 8(c)'s: the three ``ldm`` files on the path and the three they import, the two RBVAE
model files, plus the sample video and its state labels used by the chinchess parity gate.  The reference is pure
Python: there is nothing to compile, "building" is a byte-for-byte copy with a SHA-256 manifest.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

FILES = [
    "src/stable-diffusion/ldm/models/autoencoder.py",
    "src/stable-diffusion/ldm/modules/diffusionmodules/__init__.py",
    "src/stable-diffusion/ldm/modules/diffusionmodules/model.py",
    "src/stable-diffusion/ldm/modules/diffusionmodules/util.py",
    "src/stable-diffusion/ldm/modules/distributions/__init__.py",
    "src/stable-diffusion/ldm/modules/distributions/distributions.py",
    "src/stable-diffusion/ldm/modules/attention.py",
    "src/stable-diffusion/ldm/util.py",
    "models/percep_RBVAE/percep_RBVAE_model.py",
    "models/contrastive_RBVAE/contrastive_RBVAE_model.py",
    "videos/chinchess_gettyimages-148739276-640_adpp.mp4",
    "videos/frames/transition_flags.txt",
]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def build(verbose=False) -> bool:
    """Copy the files; returns True if oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF_ROOT):
        return available()
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("oracle/_ref <-", rel)
        manifest[rel] = sha256(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=REF_ROOT, files=manifest), f, indent=1)
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def root() -> str | None:
    """Where the reference tree is readable from: the live /root/reference if present, else oracle/_ref."""
    if os.path.isdir(os.path.join(REF_ROOT, "src", "stable-diffusion", "ldm")):
        return REF_ROOT
    return DEST if available() else None


if __name__ == "__main__":
    print("built" if build(verbose=True) else "no reference available")
