"""Minimal stand-in for the parts of pytorch-argus==1.0.0 that the hot path touches.

pytorch-argus is a third-party dependency of the reference (requirements.txt:7) that is neither vendored
under /root/reference nor installed here.  It does no arithmetic on this path: it is the class registry,
the device plumbing and the checkpoint container.  If the real package is importable it is used instead
(see ``argus_models.py``).  Names mirror the call sites argus_models.py:3-6,13 / predictors.py:25 / ema.py:61-73.
"""
from __future__ import annotations

import collections.abc
from types import SimpleNamespace
from typing import Any

import torch

State = SimpleNamespace


def deep_to(obj: Any, device, **kw):
    if torch.is_tensor(obj):
        return obj.to(device, **kw)
    if isinstance(obj, (str, bytes)):
        return obj
    if isinstance(obj, collections.abc.Mapping):
        return {k: deep_to(v, device, **kw) for k, v in obj.items()}
    if isinstance(obj, collections.abc.Sequence):
        return [deep_to(v, device, **kw) for v in obj]
    if isinstance(obj, torch.nn.Module):
        return obj.to(device, **kw)
    return obj


def deep_detach(obj: Any):
    if torch.is_tensor(obj):
        return obj.detach()
    if isinstance(obj, (str, bytes)):
        return obj
    if isinstance(obj, collections.abc.Mapping):
        return {k: deep_detach(v) for k, v in obj.items()}
    if isinstance(obj, collections.abc.Sequence):
        return [deep_detach(v) for v in obj]
    return obj


def deep_chunk(obj: Any, chunks: int, dim: int = 0):
    if torch.is_tensor(obj):
        return list(torch.chunk(obj, chunks, dim))
    if isinstance(obj, collections.abc.Sequence) and not isinstance(obj, (str, bytes)):
        parts = [deep_chunk(v, chunks, dim) for v in obj]
        return [[p[i] for p in parts] for i in range(min(len(p) for p in parts))]
    if isinstance(obj, collections.abc.Mapping):
        parts = {k: deep_chunk(v, chunks, dim) for k, v in obj.items()}
        n = min(len(p) for p in parts.values())
        return [{k: p[i] for k, p in parts.items()} for i in range(n)]
    return [obj for _ in range(chunks)]


class Model:
    """argus.Model look-alike: builds nn_module / loss / optimizer from ``params`` via class registries."""
    nn_module: dict = {}
    loss: dict = {}
    optimizer: dict = {}
    prediction_transform = staticmethod(lambda x: x)

    def __init__(self, params: dict):
        self.params = dict(params)
        self.device = torch.device(params.get("device", "cuda:0"))
        name, kw = params["nn_module"]
        self.nn_module = type(self).nn_module[name](**kw).to(self.device)
        self.loss = None
        if params.get("loss") is not None:
            lname, lkw = params["loss"] if not isinstance(params["loss"], str) else (params["loss"], {})
            self.loss = type(self).loss[lname](**lkw)
        self.optimizer = None
        if params.get("optimizer") is not None:
            oname, okw = params["optimizer"]
            reg = type(self).optimizer
            cls = reg[oname] if oname in reg else getattr(torch.optim, oname)
            self.optimizer = cls(self.nn_module.parameters(), **okw)
        self.prediction_transform = lambda x: x

    def train(self, mode: bool = True):
        self.nn_module.train(mode)

    def eval(self):
        self.nn_module.eval()

    def get_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def set_lr(self, lr):
        for g in self.optimizer.param_groups:
            g["lr"] = lr

    def _check_predict_ready(self):
        assert self.nn_module is not None

    def save(self, file_path):
        torch.save({"model_name": type(self).__name__, "params": self.params,
                    "nn_state_dict": deep_to(self.nn_module.state_dict(), "cpu")}, file_path)


class Metric:
    """argus.metrics.Metric look-alike (reset / update / compute / epoch_complete protocol, metrics.py:34)."""
    name: str = ""
    better: str = "min"

    def reset(self):
        pass

    def update(self, step_output: dict):
        pass

    def compute(self):
        raise NotImplementedError

    def epoch_start(self, state):
        self.reset()

    def iteration_complete(self, state):
        self.update(state.step_output)

    def epoch_complete(self, state):
        state.metrics[self.name] = self.compute()


_MODEL_REGISTRY = {}


def register_model(cls):
    _MODEL_REGISTRY[cls.__name__] = cls
    return cls


def load_model(file_path, device=None, optimizer=None, loss=None, **kw):
    """argus.load_model for the checkpoint layout {'model_name','params','nn_state_dict'} (ema.py:67-72)."""
    state = torch.load(file_path, map_location="cpu", weights_only=False)
    params = dict(state["params"])
    if device is not None:
        params["device"] = device
    if optimizer is None:
        params["optimizer"] = None
    if loss is None:
        params["loss"] = None
    cls = _MODEL_REGISTRY[state["model_name"]]
    model = cls(params)
    model.nn_module.load_state_dict(state["nn_state_dict"])
    model.eval()
    return model


pytorch_losses = {
    "MSELoss": torch.nn.MSELoss, "L1Loss": torch.nn.L1Loss, "PoissonNLLLoss": torch.nn.PoissonNLLLoss,
    "SmoothL1Loss": torch.nn.SmoothL1Loss, "HuberLoss": torch.nn.HuberLoss,
}
