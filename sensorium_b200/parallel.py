"""Data-parallel gradient exchange for DwiseNeuro training (BASELINE.json configs[2]).

One process per GPU (``torch.distributed``, NCCL over NVLink 5 / NVSwitch).  The reference has no
distributed path at all; this is DDP semantics: local BatchNorm statistics, gradients
averaged over ranks.  The exchange is *bucketed and overlapped with backward*: ``engine_bwd.run_backward``
hands every finished group of gradients to ``reduce()`` as soon as its kernels are enqueued — the ten readout
weight gradients (95 % of the bytes) go first, as one ~65 MB all-reduce each, and overlap with the whole
core backward; the ~190 small core/cortex gradients are coalesced into one flat bucket per block.
Mice absent from a rank's batch contribute a zero bucket; a per-mouse "has-grad" flag is MAX-reduced so a
mouse absent on *every* rank is skipped by the optimizer exactly like ``grad is None`` in the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

_BIG = 1 << 20


class DataParallelGrads:
    def __init__(self, mod, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.works: List = []
        self._post: List = []
        self.mod = mod
        params = list(mod.parameters())
        self.param_mouse = torch.full((len(params),), -1, dtype=torch.int64)
        index = {id(p): i for i, p in enumerate(params)}
        for m, r in enumerate(mod.readouts):
            for p in r.parameters():
                self.param_mouse[index[id(p)]] = m
        self.n_mice = len(mod.readouts)
        self.active: Optional[torch.Tensor] = None
        self._avg = dist.get_backend(group) == "nccl"  # gloo (CPU tests) has no ReduceOp.AVG
        self._pm_dev = None
        # SMs left to the NCCL kernels while backward runs next to the exchange (engine_bwd sizes its persistent grids for
        # the rest); 0 = none
        self.sm_reserve = int(__import__("os").environ.get("DWN_DP_SM_RESERVE", "0"))
        self._pinned: List = []
        self.bytes_reduced = 0

    def __deepcopy__(self, memo):
        return None  # copies of the module (ModelEma) are not trained: no exchange state to carry

    @staticmethod
    def attach(mod, optimizer, group=None) -> "DataParallelGrads":
        """Broadcast rank 0's parameters / buffers, enable the overlapped gradient exchange and wire the optimizer's
        skip flags.  ``optimizer`` is mandatory: without the flags a mouse that is absent on EVERY rank would receive a
        zero (non-None) gradient and AdamW would still decay its readout, unlike the reference where ``grad is None``
        skips the tensor.  Pass ``optimizer=False`` only when the caller consumes ``.active`` itself."""
        if optimizer is None:
            raise ValueError("DataParallelGrads.attach needs the optimizer (its has-grad flags come from the exchange)")
        with torch.no_grad():
            for t in mod.state_dict().values():
                dist.broadcast(t, src=0, group=group)
        from .engine import bump_generation
        bump_generation()
        mod._dp = DataParallelGrads(mod, group)
        if optimizer is not False:
            optimizer.active_provider = mod._dp
        return mod._dp

    # called by engine_bwd.run_backward ---------------------------------------------------------------
    def begin(self, local_live: List[bool], dev, mice: Optional[List[int]] = None) -> None:
        """``local_live[j]``: the j-th readout output of this forward has a gradient; ``mice[j]`` is its mouse index
        (forward with ``index=m`` returns one output).  The flags always cover all ``n_mice``."""
        self.bytes_reduced = 0
        if mice is None:
            mice = list(range(len(local_live)))
        row = [0] * self.n_mice
        for j, v in enumerate(local_live):
            if v:
                row[mice[j]] = 1
        host = torch.tensor(row, dtype=torch.int32)
        if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
            host = host.pin_memory()               # memcpy node of a captured step: re-read at every replay
            self._pinned.append(host)
        flags = host.to(dev, non_blocking=True)
        self.works.append(dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group, async_op=True))
        self._flags = flags

    def consumed(self) -> None:
        """Called by the optimizer after a step: the next backward starts a new accumulation window."""
        self._flags_acc = None
        self.active = None

    def reduce(self, grads: Dict[torch.Tensor, torch.Tensor], keys) -> None:
        keys = [k for k in keys if grads.get(k) is not None]
        small = []
        for k in keys:
            g = grads[k]
            if g.numel() >= _BIG:
                self.works.append(self._allreduce_mean(g))
                self.bytes_reduced += g.numel() * 4
            else:
                small.append(k)
        if small:
            flat = torch.cat([grads[k].reshape(-1) for k in small])
            self.works.append(self._allreduce_mean(flat))
            self.bytes_reduced += flat.numel() * 4
            off = 0
            for k in small:
                n = grads[k].numel()
                grads[k] = flat[off:off + n].view(grads[k].shape)
                off += n

    def reduce_readouts(self, grads: Dict[torch.Tensor, torch.Tensor], mice: Optional[List[int]] = None) -> None:
        """Exchange the readout gradients.  Every rank must issue the SAME sequence of collectives, so the
        readouts go in mouse order on all ranks; a mouse without a local sample contributes a zero bucket
        (its has-grad flag was MAX-reduced in ``begin``).  ``mice``: the mice of this forward (all, or [index])."""
        readouts = self.mod.readouts if mice is None else [self.mod.readouts[m] for m in mice]
        for r in readouts:
            ps = list(r.parameters())
            for p in ps:
                if grads.get(p) is None:
                    grads[p] = torch.zeros_like(p)
            self.reduce(grads, ps)

    def _allreduce_mean(self, t):
        if self._avg:
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._post.append(t)
        return w

    def finish(self, dev) -> None:
        for w in self.works:
            w.wait()  # stream-level wait: the compute stream waits for the NCCL stream, the host does not block
        self.works.clear()
        for t in self._post:
            t.mul_(1.0 / self.world)
        self._post.clear()
        if self._pm_dev is None or self._pm_dev.device != dev:
            self._pm_dev = self.param_mouse.to(dev)
        pm = self._pm_dev
        one = torch.ones((), dtype=torch.int32, device=dev)
        # iter_size > 1: a mouse is live for the optimizer step if ANY micro-batch since the last step had it
        acc = getattr(self, "_flags_acc", None)
        self._flags_acc = self._flags if acc is None else torch.maximum(acc, self._flags)
        self.active = torch.where(pm >= 0, self._flags_acc[pm.clamp(min=0)], one).to(torch.int32).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# Inference (C5): embarrassingly parallel — trials are sharded across ranks, every rank holds all models.
# ---------------------------------------------------------------------------------------------------------------------
def shard_trials(num_trials: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """Trial indices handled by ``rank``: round-robin, so that long and short trials spread evenly.  Without arguments
    the rank / world size of the default process group are used (1 process: every trial)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return list(range(rank, num_trials, world))


def predict_trials_sharded(predict_fn, trials, group=None) -> Dict[int, "torch.Tensor"]:
    """Run ``predict_fn(trial) -> Tensor`` on this rank's shard of ``trials`` and gather every result on every rank
    (no collective on the hot path; one ``all_gather_object`` of host arrays at the end).  Returns {trial index: result}."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = {i: predict_fn(trials[i]) for i in shard_trials(len(trials), rank, world)}
    mine = {i: (v.detach().cpu() if torch.is_tensor(v) else v) for i, v in mine.items()}
    if world == 1:
        return mine
    parts: List = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    out: Dict[int, "torch.Tensor"] = {}
    for p in parts:
        out.update(p)
    return out
