"""CutMix and batch collation on the device.

The reference mixes samples inside the DataLoader workers on the CPU (mixers.py:52-67 called from
datasets.py:124-129) and collates ten mostly-zero target tensors per sample (datasets.py:172-187), so 1.5x samples are
loaded and ~160 MB of zeros are uploaded per batch of 32.  Here the *decisions* stay on the host and draw the same
numpy calls per sample as ``CutMix`` itself (``use()`` -> ``np.random.random()``, then ``np.random.beta``, then the two
``np.random.randint`` of ``rand_bbox``).  This reproduces ``CutMix.__call__`` under a given numpy state; it does NOT
reproduce the reference *dataset's* stream, which reseeds numpy inside ``get_sample_tensors(index + 1)`` between
``use()`` and the beta draw (datasets.py:124-129, utils.py:12-15).  The *work* — box copy, target lerp, scatter of the compact per-sample
targets into the per-mouse tensors — is three kernels on the device batch."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from ._lib import call


def rand_bbox(height: int, width: int, lam: float):
    """mixers.py:36-49, verbatim semantics (bbx* are clipped to ``width`` but index the H axis in CutMix)."""
    cut_rat = np.sqrt(lam)
    cut_w = (width * cut_rat).astype(int)
    cut_h = (height * cut_rat).astype(int)
    cx = np.random.randint(width)
    cy = np.random.randint(height)
    bbx1 = np.clip(cx - cut_w // 2, 0, width)
    bby1 = np.clip(cy - cut_h // 2, 0, height)
    bbx2 = np.clip(cx + cut_w // 2, 0, width)
    bby2 = np.clip(cy + cut_h // 2, 0, height)
    return bbx1, bby1, bbx2, bby2


class DeviceCutMix:
    """CutMix(alpha, prob) of the reference applied to a whole device batch."""

    def __init__(self, alpha: float = 1.0, prob: float = 1.0):
        self.alpha = alpha
        self.prob = prob

    def sample(self, batch_size: int, height: int, width: int) -> Tuple[np.ndarray, np.ndarray]:
        """Per sample, in order: use() (mixers.py:13-14), beta, rand_bbox.  Returns boxes (B, 4) int32
        {bbx1, bby1, bbx2, bby2} (all zero for a sample that is not mixed) and the target weight lam (B,) float32 =
        box area / (h * w) (mixers.py:65)."""
        boxes = np.zeros((batch_size, 4), dtype=np.int32)
        lams = np.zeros((batch_size,), dtype=np.float32)
        for b in range(batch_size):
            if np.random.random() < self.prob:
                lam = np.random.beta(self.alpha, self.alpha)
                bbx1, bby1, bbx2, bby2 = rand_bbox(height, width, lam)
                boxes[b] = (bbx1, bby1, bbx2, bby2)
                lams[b] = (bbx2 - bbx1) * (bby2 - bby1) / (height * width)
        return boxes, lams

    def __call__(self, inputs1: torch.Tensor, inputs2: torch.Tensor, targets1: torch.Tensor, targets2: torch.Tensor,
                 boxes: np.ndarray | None = None, lams: np.ndarray | None = None):
        """inputs*: (B, C, T, H, W) fp32 on the device, targets*: (B, n, T) compact per-sample targets (same mouse).
        Returns (mixed inputs, mixed targets)."""
        if not inputs1.is_cuda:
            raise RuntimeError("DeviceCutMix runs on CUDA tensors: no CPU fallback")
        B, H, W = inputs1.shape[0], inputs1.shape[-2], inputs1.shape[-1]
        if boxes is None:
            boxes, lams = self.sample(B, H, W)
        dev = inputs1.device
        st = torch.cuda.current_stream(dev).cuda_stream
        x1, x2 = inputs1.float().contiguous(), inputs2.float().contiguous()
        t1, t2 = targets1.float().contiguous(), targets2.float().contiguous()
        bx = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32)).to(dev, non_blocking=True)
        lm = torch.from_numpy(np.ascontiguousarray(lams, dtype=np.float32)).to(dev, non_blocking=True)
        out = torch.empty_like(x1)
        call("dwn_cutmix", x1, x2, bx, out, B, x1[0].numel() // (H * W), H, W, st)
        tout = torch.empty_like(t1)
        call("dwn_lerp_rows", t1, t2, lm, tout, B, t1[0].numel(), st)
        return out, tout


def collate_on_device(compact_targets: torch.Tensor, mouse_ids: torch.Tensor,
                      num_neurons: Sequence[int]) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """ConcatMiceVideoDataset.construct_mice_sample + default collate (datasets.py:172-187) on the device:
    compact_targets (B, n_max, T) holds each sample's own-mouse responses (rows >= n_m unused), mouse_ids (B,) int.
    Returns ([ (B, n_m, T) for every mouse ], mice_weights (B, num_mice) one-hot)."""
    if not compact_targets.is_cuda:
        raise RuntimeError("collate_on_device runs on CUDA tensors: no CPU fallback")
    dev = compact_targets.device
    st = torch.cuda.current_stream(dev).cuda_stream
    comp = compact_targets.float().contiguous()
    ids = mouse_ids.to(device=dev, dtype=torch.int32).contiguous()
    B, n_max, T = comp.shape
    targets = []
    for m, n_m in enumerate(num_neurons):
        out = torch.empty((B, n_m, T), dtype=torch.float32, device=dev)
        call("dwn_scatter_mouse_targets", comp, ids, m, out, B, n_m, n_max, T, st)
        targets.append(out)
    weights = torch.zeros((B, len(num_neurons)), dtype=torch.float32, device=dev)
    weights.scatter_(1, ids.long()[:, None], 1.0)
    return targets, weights
