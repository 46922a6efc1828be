"""The data formats either side of the hot path.

* ``save_embeddings_npy`` / ``load_embeddings_npy`` / ``lookup_embedding`` -- the reference's on-disk format,
  bit-compatible: ``np.save`` of a pickled dict {frame basename: float32 (1,4,h,w)}
  (src/stable-diffusion/get_percep_embeddings.py:106,113; read at models/percep_RBVAE/percep_RBVAE_train.py:204).
* ``FlatEmbeddingStore`` -- the faster sibling: one contiguous float32 ``[N,4,h,w]`` ``.npy`` (memory-mappable,
  no pickle) plus a ``.keys.json`` index; converts to and from the pickled dict.
* ``ShuffledStatePairDataset`` -- percep_RBVAE_train.py:181-360 with the embeddings resident in HBM: same
  constructor, same contiguous train/val/test split, same pair building (consumes ``random`` exactly like the
  reference, so a seeded run yields the same pairs), ``__getitem__`` is one device gather returning the
  reference's ``[2, T, C, H, W]``; ``batch(indices)`` serves a whole mini-batch ``[B, 2, T, C, H, W]`` in one
  gather, which is what the DataLoader + default collate produce from ``__getitem__``.
"""
from __future__ import annotations

import json
import os
import random
from pathlib import Path

import numpy as np
import torch


# ---- the reference's pickled-dict .npy -------------------------------------------------
def save_embeddings_npy(path: str, keys: list[str], latents: torch.Tensor | np.ndarray):
    """get_percep_embeddings.py:106,113: np.save of {basename: float32 ndarray (1,4,h,w)}."""
    lat = latents.detach().cpu().numpy() if isinstance(latents, torch.Tensor) else latents
    if len(keys) != lat.shape[0]:
        raise ValueError(f"{len(keys)} keys for {lat.shape[0]} latents")
    emb = {k: np.ascontiguousarray(lat[i:i + 1]).astype(np.float32) for i, k in enumerate(keys)}
    np.save(path, emb)


def load_embeddings_npy(path: str) -> dict:
    """percep_RBVAE_train.py:204: np.load(path, allow_pickle=True).item()."""
    return np.load(path, allow_pickle=True).item()


def lookup_embedding(emb: dict, index: int) -> np.ndarray:
    """percep_RBVAE_train.py:337-360 `_load_embedding`: key with and without '.jpg'."""
    base = f"{index:010d}"
    for k in (base + ".jpg", base):
        if k in emb:
            return emb[k]
    raise KeyError(f"No embedding found for frame index {index}")


def frame_key(index: int, ext: str = ".jpg") -> str:
    """The key the frame extractor's file names give (scripts/decord_frame_extraction.py: %010d.jpg)."""
    return f"{index:010d}{ext}"


# ---- flat sibling -----------------------------------------------------------------------
class FlatEmbeddingStore:
    """float32 [N,4,h,w] in one plain .npy + keys in <path>.keys.json."""

    def __init__(self, keys: list[str], latents):
        self.keys = list(keys)
        self.latents = latents                       # np.ndarray / np.memmap / torch.Tensor, [N,4,h,w]
        if len(self.keys) != self.latents.shape[0]:
            raise ValueError(f"{len(self.keys)} keys for {self.latents.shape[0]} latents")
        self._row = {k: i for i, k in enumerate(self.keys)}

    @staticmethod
    def _paths(path):
        path = str(path)
        if not path.endswith(".npy"):
            path += ".npy"
        return path, path[:-4] + ".keys.json"

    def save(self, path):
        npy, idx = self._paths(path)
        lat = self.latents.detach().cpu().numpy() if isinstance(self.latents, torch.Tensor) else np.asarray(self.latents)
        np.save(npy, np.ascontiguousarray(lat, dtype=np.float32))
        with open(idx, "w") as f:
            json.dump(self.keys, f)

    @classmethod
    def load(cls, path, mmap=True):
        npy, idx = cls._paths(path)
        with open(idx) as f:
            keys = json.load(f)
        return cls(keys, np.load(npy, mmap_mode="r" if mmap else None))

    @classmethod
    def from_pickled(cls, emb_or_path):
        emb = load_embeddings_npy(emb_or_path) if isinstance(emb_or_path, (str, Path)) else emb_or_path
        keys = list(emb.keys())
        lat = np.concatenate([np.asarray(emb[k], dtype=np.float32).reshape((1,) + np.asarray(emb[k]).shape[-3:])
                              for k in keys]) if keys else np.zeros((0, 4, 0, 0), np.float32)
        return cls(keys, lat)

    def to_pickled(self) -> dict:
        lat = self.latents.detach().cpu().numpy() if isinstance(self.latents, torch.Tensor) else np.asarray(self.latents)
        return {k: np.ascontiguousarray(lat[i:i + 1]).astype(np.float32) for i, k in enumerate(self.keys)}

    def row(self, frame_index: int) -> int:
        base = f"{frame_index:010d}"
        for k in (base + ".jpg", base):
            if k in self._row:
                return self._row[k]
        raise KeyError(f"No embedding found for frame index {frame_index}")

    def get(self, frame_index: int):
        return self.latents[self.row(frame_index)][None]

    def to_device(self, device="cuda") -> "FlatEmbeddingStore":
        lat = self.latents if isinstance(self.latents, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(self.latents))
        return FlatEmbeddingStore(self.keys, lat.to(device=device, dtype=torch.float32))

    def __len__(self):
        return len(self.keys)


# ---- HBM-resident training dataset ------------------------------------------------------
class ShuffledStatePairDataset(torch.utils.data.Dataset):
    """percep_RBVAE_train.py:181-360; embeddings live on `device` as one tensor."""

    def __init__(self, input_embeddings, state_segments, test_pct=0.1, val_pct=0.1, transform=None, mode="train",
                 device="cuda"):
        super().__init__()
        if isinstance(input_embeddings, FlatEmbeddingStore):
            store = input_embeddings
        elif isinstance(input_embeddings, (str, Path)) and os.path.exists(FlatEmbeddingStore._paths(input_embeddings)[1]):
            store = FlatEmbeddingStore.load(input_embeddings)
        else:
            store = FlatEmbeddingStore.from_pickled(input_embeddings)
        self.store = store.to_device(device)
        self.input_embeddings = self.store            # the reference's attribute name
        self.state_segments = state_segments
        self.transform = transform
        self.mode = mode.lower().strip()
        self.num_states = len(self.state_segments)
        self.train_indices_per_state = []
        self.test_indices_per_state = []
        self.val_indices_per_state = []
        self.pairs_per_state = []
        for (start, end) in self.state_segments:                      # :226-262 contiguous middle chunk = test+val
            full = list(range(start, end))
            n = len(full)
            test_val_count = int(n * (test_pct + val_pct))
            margin = (n - test_val_count) // 2
            middle = full[margin:margin + test_val_count]
            train = full[:margin] + full[margin + test_val_count:]
            if test_val_count > 0:
                test_count = int(round(test_pct / (test_pct + val_pct) * test_val_count))
                test, val = middle[:test_count], middle[test_count:]
            else:
                test, val = [], []
            self.train_indices_per_state.append(train)
            self.test_indices_per_state.append(test)
            self.val_indices_per_state.append(val)
        self._build_pairs()

    def _build_pairs(self):
        """:269-318 -- draws from the global ``random`` in the reference's order."""
        try:
            per_state = {"train": self.train_indices_per_state, "test": self.test_indices_per_state,
                         "val": self.val_indices_per_state}[self.mode]
        except KeyError:
            raise ValueError(f"Unknown mode={self.mode}")
        self.pairs_per_state = []
        max_frames = max([len(ix) for ix in per_state] + [0])
        for indices in per_state:
            if 0 < len(indices) < max_frames:
                padded = indices.copy() + random.choices(indices, k=max_frames - len(indices))
            else:
                padded = indices.copy()
            random.shuffle(padded)
            pairs = [(padded[2 * i], padded[2 * i + 1]) for i in range(len(padded) // 2)]
            if len(padded) % 2 == 1:
                leftover = padded[-1]
                candidate = random.choice([x for x in indices if x != leftover]) if len(indices) > 1 else leftover
                pairs.append((leftover, candidate))
            self.pairs_per_state.append(pairs)
        self.num_items = max(len(p) for p in self.pairs_per_state)
        # row table [num_items, 2, T] into the resident tensor (the modulo wrap of __getitem__ :331 baked in)
        if all(len(p) > 0 for p in self.pairs_per_state):
            tab = [[[self.store.row(p[i % len(p)][j]) for p in self.pairs_per_state] for j in (0, 1)]
                   for i in range(self.num_items)]
            self._rows = torch.tensor(tab, dtype=torch.long, device=self.store.latents.device).reshape(
                self.num_items, 2, self.num_states)
        else:
            self._rows = None

    def __len__(self):
        return self.num_items

    def _gather(self, rows):
        out = self.store.latents[rows]
        return self.transform(out) if self.transform is not None else out

    def __getitem__(self, idx):
        """-> [2, T, C, H, W] on the device (:320-335)."""
        if self._rows is None:
            raise ValueError(f"State {idx} has no pairs")
        return self._gather(self._rows[idx])

    def batch(self, indices):
        """-> [B, 2, T, C, H, W] in one gather (what DataLoader's default collate stacks)."""
        if self._rows is None:
            raise ValueError("a state has no pairs")
        idx = torch.as_tensor(indices, dtype=torch.long, device=self._rows.device)
        return self._gather(self._rows[idx])

    def _load_embedding(self, frame_index):
        """:337-360 -- one embedding, squeezed, on the device."""
        return self.store.latents[self.store.row(frame_index)].squeeze()
