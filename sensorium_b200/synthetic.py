"""Seeded synthetic workloads of the DwiseNeuro hot path for bench.py and examples.

The same generators exist in ``oracle/dwiseneuro_oracle.py`` for the tests (``tests/test_host.py`` asserts that both
produce identical tensors); this copy keeps the GPU arm of ``bench.py`` free of any ``oracle`` import."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F


def synthetic_clip(batch: int, frames: int = 16, size: int = 64, seed: int = 0) -> torch.Tensor:
    """(B, 5, T, size, size) fp32 mimicking StackInputsProcessor (inputs.py:24-33): channel 0 = raw video 0..255 on the
    36/64 central rows, zero padding elsewhere; channels 1-4 = behaviour / pupil scalars broadcast over H, W."""
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(batch, 5, frames, size, size)
    lo, hi = (size - 36 * size // 64) // 2, (size - 36 * size // 64) // 2 + 36 * size // 64
    x[:, 0, :, lo:hi, :] = torch.randint(0, 256, (batch, frames, hi - lo, size), generator=g).float()
    scal = torch.rand(batch, 4, frames, generator=g) * torch.tensor([5.0, 10.0, 40.0, 30.0])[None, :, None]
    x[:, 1:] = scal[:, :, :, None, None]
    return x


def synthetic_targets(batch: int, n_out: Sequence[int], frames: int = 16, seed: int = 1):
    """Dense batch format of datasets.py:172-187: one labelled mouse per sample, zero tensors + weight 0 for the rest."""
    g = torch.Generator().manual_seed(seed)
    mice = torch.randint(0, len(n_out), (batch,), generator=g)
    w = F.one_hot(mice, len(n_out)).float()
    tg = []
    for m, n in enumerate(n_out):
        t = torch.relu(torch.randn(batch, n, frames, generator=g)) * 3.0
        t = t * (mice == m).float()[:, None, None]
        tg.append(t)
    return tg, w


def compact_from_dense(targets, weights):
    """Dense per-mouse targets + one-hot weights -> (compact (B, n_max, T), mouse_ids (B,) int64): the form
    ``MouseModel.train_step`` accepts to upload 17 MB instead of 160 MB per batch of 32."""
    ids = weights.argmax(1)
    B, T = weights.shape[0], targets[0].shape[-1]
    n_max = max(t.shape[1] for t in targets)
    comp = torch.zeros(B, n_max, T)
    for b in range(B):
        t = targets[int(ids[b])][b]
        comp[b, :t.shape[0]] = t
    return comp, ids


def raw_from_dense(x: torch.Tensor):
    """Dense clip batch (B, 5, T, S, S) of ``synthetic_clip`` -> the raw form ``MouseModel.train_step`` also accepts:
    (video (B, T, 36*S/64, S) uint8, scalars (B, 4, T) fp32)."""
    size = x.shape[-1]
    lo, hi = (size - 36 * size // 64) // 2, (size - 36 * size // 64) // 2 + 36 * size // 64
    video = x[:, 0, :, lo:hi, :].to(torch.uint8).contiguous()
    scalars = x[:, 1:, :, 0, 0].contiguous()
    return video, scalars


def synthetic_trial(length: int = 300, seed: int = 0):
    """Raw trial as the reference stores it (predictors.py:36-40): video (36, 64, L) uint8, behaviour (2, L), pupil
    centre (2, L) float32."""
    rs = np.random.RandomState(seed)
    video = rs.randint(0, 256, (36, 64, length)).astype(np.uint8)
    behavior = (rs.rand(2, length) * np.array([[5.0], [10.0]])).astype(np.float32)
    pupil = (rs.rand(2, length) * np.array([[40.0], [30.0]])).astype(np.float32)
    return video, behavior, pupil
