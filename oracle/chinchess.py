"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Helpers for the chinchess fixture tests/golden/chinchess_480x64x128.npz: the reference's sample video
(videos/chinchess_*.mp4, 480 frames, transition_flags.txt: chinese_chess) resized as load_img does
(get_percep_embeddings.py:48-71, at a 128x72 target), run through the unmodified reference classes by
oracle/make_golden.py.  The video itself does not travel; the fixture holds the resized frames as
wrap-around deltas plus the reference's h / codes / latent checksums.
"""
from __future__ import annotations

import numpy as np

from . import rbvae

HW = (64, 128)     # 768x432 -> LANCZOS 128x72 -> top-left crop to a multiple of 32 (load_img :63-66)
L = 25             # best latent_dim for this video (best_models.txt:17-18)
FC_GAIN, BIAS_GAIN, IH_GAIN = 40.0, 0.02, 4.0
RBVAE_SEED = 11


def rbvae_weights():
    """Default-init RBVAE weights give one constant code (the LSTM biases decide every sign) and no trained
    checkpoint exists offline: boost fc and the LSTM input weights, damp the biases, so the
    code follows the frame (27 distinct codes over the 480 frames; 745 of 12000 |h| inside the 1e-3 band)."""
    fh, fw = HW[0] // 8, HW[1] // 8
    for _ in range(3):
        fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
    sd = rbvae.init_state_dict(4, L, (fh, fw), channels=256, num_layers=4, seed=RBVAE_SEED)
    sd["encoder_cnn.fc.weight"] = sd["encoder_cnn.fc.weight"] * FC_GAIN
    for k in list(sd):
        if "bias" in k and ("lstm" in k or k.endswith("fc.bias")):
            sd[k] = sd[k] * BIAS_GAIN
        if "weight_ih" in k:
            sd[k] = sd[k] * IH_GAIN
    return sd, (fh, fw)


def frames_from_delta(delta: np.ndarray) -> np.ndarray:
    """uint8 wrap-around prefix sum along the frame axis (inverse of make_golden's delta coding)."""
    return np.cumsum(delta.astype(np.uint64), axis=0).astype(np.uint8)
