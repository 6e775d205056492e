"""autograd bridge: one Function for the whole network (forward saves raw intermediates, backward runs
the hand-written backward kernels and returns the parameter gradients in ``mod.parameters()`` order)."""
from __future__ import annotations

import torch


class DwiseNeuroFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, index, mode, x, *params):
        from . import engine
        outs, saved = engine.run_forward(mod, x, index, mode, mod.training, save=True)
        ctx.set_materialize_grads(False)
        ctx.mod = mod
        ctx.saved = saved
        ctx.n_params = len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        from . import engine_bwd
        grads = engine_bwd.run_backward(ctx.mod, ctx.saved, grad_outs)
        ctx.saved = None
        return (None, None, None, None, *grads)
