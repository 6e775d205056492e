"""Predictor — same surface as /root/reference/src/predictors.py:14-55, batched on the device.

The reference runs ~270 batch-1 forwards per trial, each followed by a device->host copy, and blends on
the CPU.  Here the windows of a trial are gathered on the device (dwn_window_gather), pushed through the
eval-mode network in chunks and overlap-added by one deterministic gather kernel (dwn_window_blend); the
result is copied to the host once.  Eval-mode BatchNorm is per-sample, so batching windows is exact."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

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
                 precision: str = "auto", _model=None):
        self.model: MouseModel = _model if _model is not None else _load_model(model_path, device=device,
                                                                                optimizer=None, loss=None)
        self.model.eval()
        self.model.nn_module._dwn_frozen = True     # inference-only weights: folded BatchNorm tables are never stale
        self.model.nn_module.precision = precision
        self.inputs_processor = get_inputs_processor(*self.model.params["inputs_processor"])
        self.frame_stack_size = self.model.params["frame_stack"]["size"]
        self.frame_stack_step = self.model.params["frame_stack"]["step"]
        assert self.model.params["frame_stack"]["position"] == "last"
        assert self.model.params["responses_processor"][0] == "identity"
        self.indexes_generator = IndexesGenerator(self.frame_stack_size, self.frame_stack_step)
        self.blend_weights = get_blend_weights(blend_weights, self.frame_stack_size)
        self.window_batch = int(window_batch)

    @classmethod
    def from_model(cls, model: "MouseModel", blend_weights="ones", window_batch: int = 32, precision: str = "auto"):
        """Predictor over an already constructed MouseModel (no checkpoint file)."""
        return cls(None, device=str(model.device), blend_weights=blend_weights, window_batch=window_batch,
                   precision=precision, _model=model)

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
    def predict_trial_raw_device(self, video: torch.Tensor, behavior: torch.Tensor, pupil_center: torch.Tensor,
                                 mouse_index: int) -> torch.Tensor:
        """Raw trial on the device -> responses (n, L) on the device.  video: (Hv, Wv, L) uint8 or fp32, behaviour and
        pupil centre: (2, L) fp32.  The clips of every window batch are assembled straight from the raw arrays
        (dwn_assemble_clips = StackInputsProcessor + window gather); the padded (5, L, H, W) tensor is never built."""
        dev = video.device
        st = torch.cuda.current_stream(dev).cuda_stream
        if video.dtype not in (torch.uint8, torch.float32):
            video = video.float()
        video = video.contiguous()
        behavior = behavior.float().contiguous()
        pupil_center = pupil_center.float().contiguous()
        Hv, Wv, L = video.shape
        W, H = self.inputs_processor.size
        fill = float(self.inputs_processor.pad_fill_value)
        size, step = self.frame_stack_size, self.frame_stack_step
        behind, ahead = self.indexes_generator.behind, self.indexes_generator.ahead
        n_out = self.model.nn_module.cfg["readout_outputs"][mouse_index]
        nwin = max(L - ahead - behind, 0)
        preds = torch.empty((max(nwin, 1), n_out, size), dtype=torch.float32, device=dev)
        vcode = 2 if video.dtype == torch.uint8 else 0
        for w0 in range(0, nwin, self.window_batch):
            nw = min(self.window_batch, nwin - w0)
            clips = torch.empty((nw, 5, size, H, W), dtype=torch.float32, device=dev)
            call("dwn_assemble_clips", video, vcode, behavior, pupil_center, clips, L, Hv, Wv, H, W, fill, size, step,
                 behind + w0, nw, st)
            preds[w0:w0 + nw] = self.model.predict(clips, mouse_index)
        blend = torch.as_tensor(np.asarray(self.blend_weights, dtype=np.float32), device=dev)
        out = torch.empty((n_out, L), dtype=torch.float32, device=dev)
        call("dwn_window_blend", preds, blend, out, n_out, L, size, step, 0, nwin, n_out * size, st)
        return out

    @torch.no_grad()
    def predict_trial(self, video: np.ndarray, behavior: np.ndarray, pupil_center: np.ndarray,
                      mouse_index: int) -> np.ndarray:
        # upload the raw arrays (uint8 video: ~36x fewer bytes than the padded fp32 stack) and assemble on the device;
        # position == "last" is asserted in __init__, so every window ends at its own frame
        dev = self.model.device
        if video.dtype != np.uint8:
            video = video.astype(np.float32, copy=False)
        out = self.predict_trial_raw_device(torch.from_numpy(np.ascontiguousarray(video)).to(dev),
                                            torch.from_numpy(np.ascontiguousarray(behavior, dtype=np.float32)).to(dev),
                                            torch.from_numpy(np.ascontiguousarray(pupil_center, dtype=np.float32)).to(dev),
                                            mouse_index)
        return out.cpu().numpy()


class EnsemblePredictor:
    """The fold ensemble of /root/reference/scripts/predict.py:43-49,65-72 — ``np.mean([predictor.predict_trial(...) for
    predictor in predictors], axis=0)`` — with every model resident on the device.

    Per trial the raw arrays are uploaded once; every window batch is assembled once (dwn_assemble_clips) and pushed
    through all models; the per-window predictions are summed over the models on the device (the overlap-add blend is
    linear, so blending the sum equals the mean of the blended responses), blended once and copied to the host once.
    ``predict_trials`` shards a list of trials round-robin over the ranks of a process group (inference
    is embarrassingly parallel, all models on every GPU, no collective on the hot path)."""

    def __init__(self, model_paths, device: str = "cuda:0", blend_weights="ones", window_batch: int = 32,
                 precision: str = "auto"):
        model_paths = list(model_paths)
        if not model_paths:
            raise ValueError("EnsemblePredictor needs at least one model path")
        self.predictors = [p if isinstance(p, Predictor) else
                           Predictor(p, device=device, blend_weights=blend_weights, window_batch=window_batch,
                                     precision=precision) for p in model_paths]
        p0 = self.predictors[0]
        for p in self.predictors[1:]:  # the reference builds every fold from the same config
            assert (p.frame_stack_size, p.frame_stack_step) == (p0.frame_stack_size, p0.frame_stack_step)
            assert p.model.nn_module.cfg["readout_outputs"] == p0.model.nn_module.cfg["readout_outputs"]
        self.device = p0.model.device
        self.window_batch = int(window_batch)

    def set_precision(self, precision: str) -> None:
        for p in self.predictors:
            p.model.nn_module.precision = precision

    @torch.no_grad()
    def predict_trial_raw_device(self, video: torch.Tensor, behavior: torch.Tensor, pupil_center: torch.Tensor,
                                 mouse_index: int) -> torch.Tensor:
        p0 = self.predictors[0]
        dev = video.device
        st = torch.cuda.current_stream(dev).cuda_stream
        if video.dtype not in (torch.uint8, torch.float32):
            video = video.float()
        video = video.contiguous()
        behavior = behavior.float().contiguous()
        pupil_center = pupil_center.float().contiguous()
        Hv, Wv, L = video.shape
        W, H = p0.inputs_processor.size
        fill = float(p0.inputs_processor.pad_fill_value)
        size, step = p0.frame_stack_size, p0.frame_stack_step
        behind, ahead = p0.indexes_generator.behind, p0.indexes_generator.ahead
        n_out = p0.model.nn_module.cfg["readout_outputs"][mouse_index]
        nwin = max(L - ahead - behind, 0)
        preds = torch.zeros((max(nwin, 1), n_out, size), dtype=torch.float32, device=dev)
        vcode = 2 if video.dtype == torch.uint8 else 0
        for w0 in range(0, nwin, self.window_batch):
            nw = min(self.window_batch, nwin - w0)
            clips = torch.empty((nw, 5, size, H, W), dtype=torch.float32, device=dev)
            call("dwn_assemble_clips", video, vcode, behavior, pupil_center, clips, L, Hv, Wv, H, W, fill, size, step,
                 behind + w0, nw, st)
            for p in self.predictors:
                preds[w0:w0 + nw] += p.model.predict(clips, mouse_index)
        blend = torch.as_tensor(np.asarray(p0.blend_weights, dtype=np.float32), device=dev)
        out = torch.empty((n_out, L), dtype=torch.float32, device=dev)
        call("dwn_window_blend", preds, blend, out, n_out, L, size, step, 0, nwin, n_out * size, st)
        return out.div_(float(len(self.predictors)))

    @torch.no_grad()
    def predict_trial(self, video: np.ndarray, behavior: np.ndarray, pupil_center: np.ndarray,
                      mouse_index: int) -> np.ndarray:
        dev = self.device
        if video.dtype != np.uint8:
            video = video.astype(np.float32, copy=False)
        out = self.predict_trial_raw_device(torch.from_numpy(np.ascontiguousarray(video)).to(dev),
                                            torch.from_numpy(np.ascontiguousarray(behavior, dtype=np.float32)).to(dev),
                                            torch.from_numpy(np.ascontiguousarray(pupil_center, dtype=np.float32)).to(dev),
                                            mouse_index)
        return out.cpu().numpy()

    def predict_trials(self, trials, group=None):
        """trials: list of dicts {"video", "behavior", "pupil_center", "mouse_index"}.  Returns {trial index: (n, L)
        float32 array} on every rank; rank r computes trials r, r + world, ... (parallel.predict_trials_sharded)."""
        from .parallel import predict_trials_sharded
        return predict_trials_sharded(
            lambda t: self.predict_trial(t["video"], t["behavior"], t["pupil_center"], t["mouse_index"]), trials, group)
