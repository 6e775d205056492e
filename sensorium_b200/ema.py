"""ModelEma — same surface as /root/reference/src/ema.py:12-58 (timm ModelEmaV2 semantics: the EMA runs over
*every* state_dict entry in order; int64 buffers go through fp32 and are truncated on copy), but one
multi-tensor kernel launch instead of ~1.4k tiny launches per step."""
from __future__ import annotations

from copy import deepcopy

import torch
from torch import nn

from . import _lib
from ._lib import call
from .optim import _chunk_tables


class ModelEma(nn.Module):
    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.ema = deepcopy(model)
        self.ema.eval()
        self.decay = decay
        self.device = device
        if self.device is not None:
            self.ema.to(device=device)
        self._key = None
        self._tab = None

    def _tables(self, model):
        ev = list(self.ema.state_dict().values())
        mv = list(model.state_dict().values())
        key = tuple((e.data_ptr(), m.data_ptr()) for e, m in zip(ev, mv))
        if key != self._key:
            dev = ev[0].device
            rows = []
            for e, m in zip(ev, mv):
                if e.device != m.device:
                    raise RuntimeError("sensorium_b200.ModelEma: EMA and model must live on the same CUDA device")
                if e.dtype == torch.float32:
                    flag = 0
                elif e.dtype == torch.int64:
                    flag = 1
                else:
                    raise RuntimeError(f"unsupported state dtype {e.dtype}")
                rows.append([m.data_ptr(), 0, 0, 0, 0, e.data_ptr(), e.numel(), flag])
            self._tab = torch.tensor(rows, dtype=torch.int64).to(dev)
            self._chunks = _chunk_tables([e.numel() for e in ev], _lib.lib().dwn_opt_chunk(), dev)
            self._key = key
            self._nelem = sum(e.numel() for e in ev)
            # the same update split in two launches: the readout entries (95 % of the bytes) and the rest — MouseModel
            # runs the first part early, under backward (see FusedAdamW.early_step)
            names = list(self.ema.state_dict().keys())
            self._parts = {}
            for part, sel in (("readouts", [i for i, n in enumerate(names) if n.startswith("readouts.")]),
                              ("rest", [i for i, n in enumerate(names) if not n.startswith("readouts.")])):
                if sel:
                    self._parts[part] = (torch.tensor([rows[i] for i in sel], dtype=torch.int64).to(dev),
                                         _chunk_tables([ev[i].numel() for i in sel], _lib.lib().dwn_opt_chunk(), dev),
                                         sum(ev[i].numel() for i in sel))
        return ev[0].device

    @torch.no_grad()
    def update(self, model, part: str = "all"):
        """``part``: "all" (the reference's update), or "readouts" / "rest" — the same update in two launches."""
        dev = self._tables(model)
        if dev.type != "cuda":
            raise RuntimeError("sensorium_b200.ModelEma runs on CUDA only: no CPU fallback")
        tab, (ct, co, nch), nelem = (self._tab, self._chunks, self._nelem) if part == "all" else self._parts[part]
        call("dwn_ema", tab, ct, co, nch, float(self.decay), torch.cuda.current_stream(dev).cuda_stream,
             _tag="ema" if part == "all" else "ema_" + part, _bytes=nelem * 12)
        from .engine import bump_generation
        bump_generation()
        # weights of the EMA module changed behind autograd's back: drop stale bf16 shadows
        for p in self.ema.parameters():
            if getattr(p, "_dwn_shadow", None) is not None:
                p._dwn_shadow = None

    @torch.no_grad()
    def set(self, model):
        for e, m in zip(self.ema.state_dict().values(), model.state_dict().values()):
            e.copy_(m)


def save_ema_model(model, file_path) -> None:
    """Write the EMA weights in the argus checkpoint layout the reference's ``EmaCheckpoint.save_model`` produces
    (ema.py:61-73): ``{'model_name', 'params', 'nn_state_dict'}`` with the state dict on the CPU, so that
    ``argus.load_model`` / ``Predictor`` load it like any other checkpoint."""
    nn_module = model.model_ema.ema
    if isinstance(nn_module, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
        nn_module = nn_module.module
    torch.save({"model_name": model.__class__.__name__, "params": model.params,
                "nn_state_dict": {k: v.detach().to("cpu") for k, v in nn_module.state_dict().items()}}, file_path)


try:  # pragma: no cover - pytorch-argus is absent in this image
    from argus.callbacks import Checkpoint as _CheckpointBase  # type: ignore
except ImportError:
    class _CheckpointBase:
        """Minimal stand-in for argus.callbacks.Checkpoint: only the hook the reference overrides."""

        def save_model(self, state, file_path):
            state.model.save(file_path)


class EmaCheckpoint(_CheckpointBase):
    """argus Checkpoint callback that stores the EMA weights instead of the raw ones (ema.py:61-73)."""

    def save_model(self, state, file_path):
        save_ema_model(state.model, file_path)
        logger = getattr(state, "logger", None)
        if logger is not None:
            logger.info(f"Model saved to '{file_path}'")
