"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the evaluation helpers around the hot path
(scripts/evaluation/state_consistency_eval/embedding_matching.py):
  state_consistency   :275-297   np.unique(axis=0) per state, share equal to the most common vector
  gaussian_noise_u8   :141-161 + ToTensor / ToPILImage round trip (:240-243)
  occlusion_u8        :165-193 + the same round trip
The reference's own add_gaussian_noise / add_occlusion / assign_label are imported by the tests when
/root/reference is present (oracle.ref_shim.embedding_matching_functions) to pin these restatements.
"""
from __future__ import annotations

import numpy as np
import torch


def state_consistency(latent_vectors: np.ndarray, labels: np.ndarray, n_states: int):
    """latent_vectors float {0,1} [N,L]; labels int [N] -> (weighted_avg, percentages)."""
    percentages = []
    for label in range(n_states):
        v = latent_vectors[labels == label]
        if len(v) == 0:
            percentages.append(0.0)
            continue
        uniq, counts = np.unique(v, axis=0, return_counts=True)
        most = uniq[np.argmax(counts)]
        percentages.append(float(np.mean(np.all(v == most, axis=1))))
    counts = [int(np.sum(labels == label)) for label in range(n_states)]
    total = sum(counts)
    weighted = float(np.dot(percentages, counts) / total) if total > 0 else 0
    return weighted, percentages


def to_tensor(u8_hwc: np.ndarray) -> torch.Tensor:
    """torchvision ToTensor on a uint8 HWC image: CHW float32 / 255."""
    return torch.from_numpy(u8_hwc).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


def to_pil_u8(t_chw: torch.Tensor) -> np.ndarray:
    """torchvision ToPILImage on a float CHW tensor: mul(255).byte(), HWC."""
    return t_chw.mul(255).byte().permute(1, 2, 0).contiguous().numpy()


def gaussian_noise_u8(u8_hwc: np.ndarray, noise_chw: torch.Tensor, mean=0.0, std=0.1) -> np.ndarray:
    t = to_tensor(u8_hwc).unsqueeze(0)
    noisy = torch.clamp(t + (noise_chw.reshape(t.shape) * std + mean), 0, 1)
    return to_pil_u8(noisy.squeeze())


def occlusion_u8(u8_hwc: np.ndarray, x: int, y: int, size: int) -> np.ndarray:
    t = to_tensor(u8_hwc).unsqueeze(0).clone()
    t[:, :, y:y + size, x:x + size] = 0.5
    return to_pil_u8(t.squeeze())
