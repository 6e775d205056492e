"""Validation metric — same surface as /root/reference/src/metrics.py:11-74, accumulated on the device.

The reference's ``CorrelationMetric.update`` boolean-indexes every prediction, transposes it and copies it to host
numpy arrays on every validation batch (one sync per mouse per batch), then concatenates everything at epoch end.
Here ``update`` launches one kernel per mouse that adds {sum x, sum y, sum xy, sum x^2, sum y^2} per neuron to fp64
accumulators (dwn_corr_update; samples with weight 0 are skipped inside the kernel, no host sync), and ``compute``
finalizes the per-neuron correlation and its mean on the device (dwn_corr_finalize).  SURVEY.md §8(f2)."""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import torch

from ._lib import call

try:  # pragma: no cover
    from argus.metrics import Metric  # type: ignore
except ImportError:
    from .argus_shim import Metric


def corr(y1: np.ndarray, y2: np.ndarray, axis: Union[None, int, Tuple[int]] = -1, eps: float = 1e-8, **kwargs) -> np.ndarray:
    """Host utility with the reference's signature (metrics.py:11-31): correlation of two numpy arrays along ``axis``
    with biased std and ``eps`` added to the std.  Used by scripts on host arrays, not by the training path."""
    y1 = (y1 - y1.mean(axis=axis, keepdims=True)) / (y1.std(axis=axis, keepdims=True, ddof=0) + eps)
    y2 = (y2 - y2.mean(axis=axis, keepdims=True)) / (y2.std(axis=axis, keepdims=True, ddof=0) + eps)
    return (y1 * y2).mean(axis=axis, **kwargs)


class CorrelationMetric(Metric):
    name: str = "corr"
    better: str = "max"

    def __init__(self, eps: float = 1e-8):
        super().__init__()
        self.eps = float(eps)
        self.reset()

    def reset(self):
        self._acc: Dict[int, torch.Tensor] = {}   # mouse -> (n, 5) fp64 accumulators
        self._cnt: Dict[int, torch.Tensor] = {}   # mouse -> (1,) fp64 masked sample count * T

    def update(self, step_output: dict):
        pred_tensors = step_output["prediction"]
        target_tensors, mice_weights = step_output["target"]
        if not mice_weights.is_cuda:
            raise RuntimeError("CorrelationMetric runs on the device: step outputs must be CUDA tensors")
        dev = mice_weights.device
        st = torch.cuda.current_stream(dev).cuda_stream
        w = mice_weights.detach().float().contiguous()
        n_mice = w.shape[-1]
        for m, (pred, target) in enumerate(zip(pred_tensors, target_tensors)):
            pred = pred.detach().float().contiguous()
            target = target.detach().float().contiguous()
            B, n = pred.shape[0], pred.shape[1]
            T = pred.shape[2] if pred.dim() == 3 else 1
            if m not in self._acc:
                self._acc[m] = torch.zeros((n, 5), dtype=torch.float64, device=dev)
                self._cnt[m] = torch.zeros((1,), dtype=torch.float64, device=dev)
            call("dwn_corr_update", pred, target, w[:, m:], n_mice, B, n, T, self._acc[m], self._cnt[m], st)

    def compute_per_neuron(self) -> Dict[int, torch.Tensor]:
        """Per-neuron correlations (device tensors) of every mouse seen so far."""
        out = {}
        for m, acc in self._acc.items():
            if float(self._cnt[m].item()) == 0.0:
                continue
            st = torch.cuda.current_stream(acc.device).cuda_stream
            per = torch.empty((acc.shape[0],), dtype=torch.float32, device=acc.device)
            mean = torch.empty((1,), dtype=torch.float32, device=acc.device)
            call("dwn_corr_finalize", acc, self._cnt[m], acc.shape[0], self.eps, per, mean, st)
            out[m] = (per, mean)
        return out

    def compute(self) -> Dict[int, np.floating]:
        # mice that never had a sample are absent, like the reference's defaultdict (metrics.py:55-66)
        return {m: np.float32(mean.item()) for m, (_, mean) in self.compute_per_neuron().items()}

    def epoch_complete(self, state):
        with torch.no_grad():
            mice_corr = self.compute()
        name_prefix = f"{state.phase}_" if state.phase else ""
        for mouse_index, mouse_corr in mice_corr.items():
            state.metrics[name_prefix + self.name + f"_mouse_{mouse_index}"] = mouse_corr
        state.metrics[name_prefix + self.name] = np.mean(list(mice_corr.values()))
