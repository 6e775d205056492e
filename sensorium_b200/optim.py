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
        self._tab = None
        self._key = None
        self._active_cache = {}

    def _build(self, group, params):
        dev = params[0].device
        sizes = [p.numel() for p in params]
        total = sum(sizes)
        if "m" not in group:
            group["m"] = torch.zeros(total, dtype=torch.float32, device=dev)
            group["v"] = torch.zeros(total, dtype=torch.float32, device=dev)
            group["steps"] = torch.zeros(len(params), dtype=torch.int32, device=dev)
            off = 0
            for i, p in enumerate(params):
                st = self.state[p]
                st["exp_avg"] = group["m"][off:off + sizes[i]].view_as(p)
                st["exp_avg_sq"] = group["v"][off:off + sizes[i]].view_as(p)
                st["step"] = group["steps"][i]
                off += sizes[i]
            group["chunks"] = _chunk_tables(sizes, _lib.lib().dwn_opt_chunk(), dev)
        rows = []
        for p in params:
            st = self.state[p]
            sh = getattr(p, "_dwn_shadow", None)
            shp = 0
            if sh is not None:
                shp = sh[1].data_ptr()
            g = p.grad
            rows.append([p.data_ptr(), g.data_ptr() if g is not None else 0, st["exp_avg"].data_ptr(),
                         st["exp_avg_sq"].data_ptr(), shp, 0, p.numel(), 0])
        return torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            params = [p for p in group["params"] if p.requires_grad]
            if not params:
                continue
            dev = params[0].device
            if not params[0].is_cuda:
                raise RuntimeError("FusedAdamW runs on CUDA only: no CPU fallback")
            key = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0,
                         id(getattr(p, "_dwn_shadow", None) and p._dwn_shadow[1])) for p in params)
            if group.get("_key") != key:
                for p in params:
                    if p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous()):
                        p.grad = p.grad.float().contiguous()
                group["_tab"] = self._build(group, params)
                group["_key"] = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0,
                                       id(getattr(p, "_dwn_shadow", None) and p._dwn_shadow[1])) for p in params)
            act = tuple(p.grad is not None for p in params)
            active = None
            provider = getattr(self, "active_provider", None)
            if provider is not None and provider.active is not None and provider.active.numel() == len(params):
                active = provider.active  # device-side flags from the data-parallel exchange
            elif not all(act):
                active = self._active_cache.get(act)
                if active is None:
                    active = torch.tensor(act, dtype=torch.int32).to(dev)
                    if len(self._active_cache) < 256:
                        self._active_cache[act] = active
            ct, co, nch = group["chunks"]
            b1, b2 = group["betas"]
            call("dwn_adamw", group["_tab"], ct, co, nch, group["steps"], active, len(params), float(group["lr"]),
                 float(group["weight_decay"]), float(b1), float(b2), float(group["eps"]), 0.0,
                 torch.cuda.current_stream(dev).cuda_stream, _tag="adamw", _bytes=sum(p.numel() for p in params) * 30)
            for p in params:
                sh = getattr(p, "_dwn_shadow", None)
                if sh is not None:
                    set_shadow(p, sh[1])
        bump_generation()
        provider = getattr(self, "active_provider", None)
        if provider is not None and hasattr(provider, "consumed"):
            provider.consumed()
        return loss
