"""MicePoissonLoss — same surface as /root/reference/src/losses.py:5-21, computed by fused CUDA kernels.

loss = sum_m sum_{b: w_hat[b,m] != 0} w_hat[b,m] * sum_{n,t} (p - y*log(p + eps)),  w_hat = w / sum(w).
Mice without a live sample are skipped, so their readout gradients stay ``None``."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from ._lib import call

_J = 16


class _PoissonFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eps, wn, live, n_mice, *tensors):
        preds, tgts = tensors[:n_mice], tensors[n_mice:]
        dev = wn.device
        st = torch.cuda.current_stream(dev).cuda_stream
        ctx.set_materialize_grads(False)
        idx = [m for m in range(n_mice) if live[m]]
        B0 = preds[0].shape[0]
        partial = torch.zeros((max(len(idx), 1), B0 * _J), dtype=torch.float64, device=dev)
        saved = []
        for j, m in enumerate(idx):
            p = preds[m].detach().contiguous()
            y = tgts[m].detach().contiguous().float()
            B = p.shape[0]
            per_b = p.numel() // B
            call("dwn_poisson_fwd", p, y, wn[:, m], wn.shape[1], B, per_b, eps, partial[j], _J, st)
            saved.append((m, p, y))
        ctx.saved = saved
        ctx.wn = wn
        ctx.eps = eps
        ctx.n_mice = n_mice
        return partial.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        grads: List[Optional[torch.Tensor]] = [None] * (2 * ctx.n_mice)
        if gout is None:
            return (None, None, None, None, *grads)
        wn = ctx.wn
        st = torch.cuda.current_stream(wn.device).cuda_stream
        g = gout.detach().to(torch.float32).contiguous()
        for m, p, y in ctx.saved:
            B = p.shape[0]
            dp = torch.empty_like(p)
            call("dwn_poisson_bwd", p, y, wn[:, m], wn.shape[1], g, B, p.numel() // B, ctx.eps, dp, st)
            grads[m] = dp
        return (None, None, None, None, *grads)


class MicePoissonLoss(nn.Module):
    def __init__(self, log_input: bool = False, full: bool = False, eps: float = 1e-8):
        super().__init__()
        if log_input or full:
            raise NotImplementedError("sensorium_b200.MicePoissonLoss implements log_input=False, full=False")
        self.eps = float(eps)
        self._live_hint: Optional[Sequence[bool]] = None

    def set_live_hint(self, live: Optional[Sequence[bool]]) -> None:
        """Optional host-side knowledge of which mice have a non-zero weight in the next batch; avoids the
        device->host sync that ``torch.any(mask)`` costs the reference (losses.py:17)."""
        self._live_hint = None if live is None else [bool(v) for v in live]

    def forward(self, inputs, targets):
        target_tensors, mice_weights = targets
        if not mice_weights.is_cuda:
            raise RuntimeError("sensorium_b200.MicePoissonLoss runs on CUDA only: no CPU fallback")
        wn = (mice_weights.float() / mice_weights.float().sum()).contiguous()
        live = self._live_hint
        self._live_hint = None
        if live is None:
            live = (wn != 0.0).any(dim=0).tolist()
        n = len(inputs)
        return _PoissonFn.apply(self.eps, wn, list(live), n, *inputs, *target_tensors)
