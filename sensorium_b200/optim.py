"""FusedAdamW — multi-tensor AdamW with torch.optim.AdamW semantics (configs/true_batch_001.py:45-48):
decoupled weight decay on every parameter, per-tensor step counters, tensors whose ``grad is None`` are
skipped entirely (no decay, step not advanced).  One kernel launch updates all parameters and refreshes
the bf16 weight shadows used by the tcgen05 GEMMs."""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib
from ._lib import call
from .engine import bump_generation, set_shadow


def _h2d(host: torch.Tensor, dev, keep: dict) -> torch.Tensor:
    """Host table -> device.  Under CUDA-graph capture the copy becomes a memcpy node that re-reads the HOST buffer at
    every replay, so the buffer is pinned and kept alive (``keep["pinned"]``); pageable copies are illegal in capture."""
    if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
        host = host.pin_memory()
        keep.setdefault("pinned", []).append(host)
    return host.to(dev, non_blocking=True)


def _chunk_tables(sizes, chunk, dev):
    ct, co = [], []
    for t, n in enumerate(sizes):
        for off in range(0, max(n, 1), chunk):
            ct.append(t)
            co.append(off)
    return (torch.tensor(ct, dtype=torch.int32, device=dev), torch.tensor(co, dtype=torch.int64, device=dev), len(ct))


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        # launch caches (flat moment buffers, pointer tables, chunk tables) live on the optimizer, keyed by group index —
        # NOT in param_groups, so state_dict() carries only torch's own {exp_avg, exp_avg_sq, step} per parameter
        self._cache = {}
        self._active_cache = {}
        self._early_done: set = set()

    def _flatten(self, gi, params):
        """Flat exp_avg / exp_avg_sq / step buffers; ``state[p]`` entries are views into them.  State that is already
        present (load_state_dict) is copied in, so a loaded optimizer continues exactly where the saved one stopped."""
        dev = params[0].device
        sizes = [p.numel() for p in params]
        total = sum(sizes)
        c = {"m": torch.zeros(total, dtype=torch.float32, device=dev),
             "v": torch.zeros(total, dtype=torch.float32, device=dev),
             "steps": torch.zeros(len(params), dtype=torch.int32, device=dev),
             "chunks": _chunk_tables(sizes, _lib.lib().dwn_opt_chunk(), dev), "key": None, "tab": None,
             # device copy of the learning rate for captured steps.  Allocated HERE, outside any capture: a tensor
             # allocated inside a capture shares its address with earlier temporaries of the same graph, which would
             # overwrite a value written before the replay
             "lr_dev": torch.zeros((1,), dtype=torch.float32, device=dev),
             "bc": torch.zeros((2 * len(params),), dtype=torch.float32, device=dev)}
        off = 0
        for i, p in enumerate(params):
            st = self.state[p]
            mv, vv, sv = c["m"][off:off + sizes[i]].view_as(p), c["v"][off:off + sizes[i]].view_as(p), c["steps"][i]
            if "exp_avg" in st:
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                sv.copy_(torch.as_tensor(st["step"]).to(torch.int32))
            st["exp_avg"], st["exp_avg_sq"], st["step"] = mv, vv, sv
            off += sizes[i]
        self._cache[gi] = c
        return c

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._cache.clear()          # the loaded state tensors no longer alias the flat buffers: re-flatten on next step

    def __setstate__(self, state):
        super().__setstate__(state)
        self._cache = {}
        self._active_cache = {}

    def _table(self, c, params, dev, grads=None):
        rows = []
        for p in params:
            st = self.state[p]
            sh = getattr(p, "_dwn_shadow", None)
            g = p.grad if grads is None else grads.get(p)
            rows.append([p.data_ptr(), g.data_ptr() if g is not None else 0, st["exp_avg"].data_ptr(),
                         st["exp_avg_sq"].data_ptr(), sh[1].data_ptr() if sh is not None else 0, 0, p.numel(), 0])
        return _h2d(torch.tensor(rows, dtype=torch.int64), dev, c)

    @torch.no_grad()
    def early_step(self, subset, grads) -> None:
        """AdamW update of ``subset`` (parameters of group 0) from ``grads[p]`` BEFORE the rest of backward has run —
        MouseModel issues it for the readouts on a side stream as soon as their gradients exist.  The following
        ``step()`` skips these tensors.  Same kernel, same per-tensor step counters: the result is identical to a
        single late step."""
        group = self.param_groups[0]
        params = [p for p in group["params"] if p.requires_grad]
        dev = params[0].device
        c = self._cache.get(0)
        if c is None or c["steps"].numel() != len(params):
            c = self._flatten(0, params)
        chosen = {id(p) for p in subset if grads.get(p) is not None}
        key = tuple(grads[p].data_ptr() if id(p) in chosen else 0 for p in params)
        ent = c.get("early")
        if ent is None or ent[0] != key:
            tab = self._table(c, params, dev, grads={p: (grads[p] if id(p) in chosen else None) for p in params})
            act = _h2d(torch.tensor([1 if id(p) in chosen else 0 for p in params], dtype=torch.int32), dev, c)
            ent = c["early"] = (key, tab, act)
        if torch.cuda.is_current_stream_capturing() and ent[1].device.type == "cuda":
            pass  # tables created inside this capture are replayed by its memcpy nodes
        ct, co, nch = c["chunks"]
        b1, b2 = group["betas"]
        lr_dev = c["lr_dev"] if torch.cuda.is_current_stream_capturing() else None
        call("dwn_adamw", ent[1], ct, co, nch, c["steps"], ent[2], len(params), float(group["lr"]),
             float(group["weight_decay"]), float(b1), float(b2), float(group["eps"]), 0.0, lr_dev, c["bc"],
             torch.cuda.current_stream(dev).cuda_stream, _tag="adamw_early",
             _bytes=sum(p.numel() for p in params if id(p) in chosen) * 30)
        self._early_done = chosen

    def sync_graph_lr(self) -> None:
        """Push the current learning rates to the device scalars that captured optimizer steps read."""
        for gi, group in enumerate(self.param_groups):
            c = self._cache.get(gi)
            if c is not None and c.get("lr_val") != group["lr"]:
                c["lr_dev"].fill_(float(group["lr"]))
                c["lr_val"] = group["lr"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.requires_grad]
            if not params:
                continue
            dev = params[0].device
            if not params[0].is_cuda:
                raise RuntimeError("FusedAdamW runs on CUDA only: no CPU fallback")
            c = self._cache.get(gi)
            if c is None or c["steps"].numel() != len(params):
                c = self._flatten(gi, params)

            def key_of():
                return tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0,
                              id(getattr(p, "_dwn_shadow", None) and p._dwn_shadow[1])) for p in params)

            if c["key"] != key_of():
                for p in params:
                    if p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous()):
                        p.grad = p.grad.float().contiguous()
                c["tab"] = self._table(c, params, dev)
                c["key"] = key_of()
            early = self._early_done if gi == 0 else set()
            act = tuple(p.grad is not None and id(p) not in early for p in params)
            active = None
            provider = getattr(self, "active_provider", None)
            if provider is not None and provider.active is not None and provider.active.numel() == len(params):
                active = provider.active  # device-side flags from the data-parallel exchange
            elif not all(act):
                active = self._active_cache.get(act)
                if active is None:
                    active = _h2d(torch.tensor(act, dtype=torch.int32), dev, c)
                    # a table created inside a capture lives in that graph's pool and is only valid inside its replay
                    if len(self._active_cache) < 256 and not torch.cuda.is_current_stream_capturing():
                        self._active_cache[act] = active
            ct, co, nch = c["chunks"]
            b1, b2 = group["betas"]
            lr_dev = None
            if torch.cuda.is_current_stream_capturing():
                # a captured step reads the learning rate from device memory; GraphedTrainStep refreshes it
                # (sync_graph_lr) before every replay, so LR schedulers keep working
                lr_dev = c["lr_dev"]
            call("dwn_adamw", c["tab"], ct, co, nch, c["steps"], active, len(params), float(group["lr"]),
                 float(group["weight_decay"]), float(b1), float(b2), float(group["eps"]), 0.0, lr_dev, c["bc"],
                 torch.cuda.current_stream(dev).cuda_stream, _tag="adamw", _bytes=sum(p.numel() for p in params) * 30)
            for p in params:
                sh = getattr(p, "_dwn_shadow", None)
                if sh is not None:
                    set_shadow(p, sh[1])
        self._early_done = set()
        bump_generation()
        provider = getattr(self, "active_provider", None)
        if provider is not None and hasattr(provider, "consumed"):
            provider.consumed()
        return loss
