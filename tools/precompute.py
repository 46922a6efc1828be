#!/usr/bin/env python
"""CLI of the full-video embedding precompute (BASELINE configs[2]); see symbols-from-video_b200/precompute.py.

    python tools/precompute.py VIDEO_OR_FRAME_FOLDER --out NAME_perceps.npy [--ckpt sd.ckpt] [--parts DIR]
    torchrun --nproc-per-node 8 tools/precompute.py VIDEO --out NAME_perceps.npy      # frame range sharded over GPUs
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfv_b200  # noqa: E402

if __name__ == "__main__":
    sfv_b200.precompute.main()
