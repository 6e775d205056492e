"""Input assembly — same API as /root/reference/src/inputs.py:1-46 (StackInputsProcessor)."""
from __future__ import annotations

import numpy as np
import torch


class StackInputsProcessor:
    """(H, W, L) video + (2, L) behaviour + (2, L) pupil centre -> float32 (5, L, size[1], size[0]),
    the video centred and padded with ``pad_fill_value`` (inputs.py:22-36)."""

    def __init__(self, size: tuple[int, int], pad_fill_value: int = 0):
        self.size = size
        self.pad_fill_value = pad_fill_value

    def __call__(self, frames: np.ndarray, behavior: np.ndarray, pupil_center: np.ndarray) -> torch.Tensor:
        length = frames.shape[-1]
        out = np.full((5, length, self.size[1], self.size[0]), self.pad_fill_value, dtype=np.float32)
        video = np.moveaxis(frames.astype(np.float32), -1, 0)
        h, w = video.shape[-2:]
        top, left = (self.size[1] - h) // 2, (self.size[0] - w) // 2
        out[0, :, top:top + h, left:left + w] = video
        out[1:3] = behavior[:, :, None, None]
        out[3:] = pupil_center[:, :, None, None]
        return torch.from_numpy(out)


_REGISTRY = {"stack_inputs": StackInputsProcessor}


def get_inputs_processor(name: str, processor_params: dict):
    assert name in _REGISTRY
    return _REGISTRY[name](**processor_params)
