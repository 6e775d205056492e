"""MouseModel — the argus model wrapper of /root/reference/src/argus_models.py:13-99 over the B200 engine.

Same attributes (iter_size, amp, grad_scaler, model_ema, distill_model, distill_ratio) and the same
train_step / val_step / predict / add_distill_predictions contracts.  Differences, all on the device side:
 * AMP uses bf16 autocast (BASELINE.json north_star) so no loss scaling is needed; ``grad_scaler`` is kept
   as a disabled GradScaler for attribute compatibility;
 * the distillation target fill is three kernel launches instead of a 288-iteration Python loop;
 * the set of live mice is taken from the host copy of the weights, so the loss needs no device sync;
 * ``train_step`` also accepts the compact batch form ``(input, (responses (B, n_max, T), mouse_ids (B,)))`` and builds
   the reference's dense per-mouse targets on the device.
"""
from __future__ import annotations

import torch

try:  # pragma: no cover - the real package is absent in this image
    import argus  # type: ignore
    from argus.engine import State  # type: ignore
    from argus.loss import pytorch_losses  # type: ignore
    from argus.utils import deep_to, deep_detach, deep_chunk  # type: ignore
    _Base = argus.Model
    _register = lambda c: c  # noqa: E731
except ImportError:
    from .argus_shim import Model as _Base, State, pytorch_losses, deep_to, deep_detach, deep_chunk
    from .argus_shim import register_model as _register

from ._lib import call
from .dwiseneuro import DwiseNeuro
from .ema import ModelEma
from .losses import MicePoissonLoss
from .mixers import collate_on_device
from .optim import FusedAdamW

# live-mouse hints of device-resident batches (DevicePrefetcher), keyed by id() of the device weight tensor of the batch:
# a Python attribute on the tensor would be lost by deep_chunk (torch.chunk returns new tensor objects)
_LIVE_HINTS: dict = {}


@_register
class MouseModel(_Base):
    nn_module = {"dwiseneuro": DwiseNeuro}
    loss = {**pytorch_losses, "mice_poisson": MicePoissonLoss}
    optimizer = {"FusedAdamW": FusedAdamW, "AdamW": FusedAdamW}

    def __init__(self, params: dict):
        super().__init__(params)
        self.iter_size = int(params.get("iter_size", 1))
        self.amp = bool(params.get("amp", False))
        self.grad_scaler = torch.amp.GradScaler("cuda", enabled=False)
        self.model_ema: ModelEma | None = None
        self.distill_model: torch.nn.Module | None = None
        self.distill_ratio: float = 0.0
        # params["cuda_graph"] = True: the whole train step (distillation fill, forward, loss, backward, AdamW, EMA) is
        # captured once per (shapes, set of mice present) and replayed — ~500 kernel launches become one graph launch
        self.cuda_graph = bool(params.get("cuda_graph", False))
        # data-parallel steps are captured too (the NCCL all-reduces become graph nodes); every rank must then see the same
        # sequence of graph keys, which holds with distillation (all mice live) or when the batch form is the same on all
        # ranks and the loader hands every rank batches with the same set of mice
        self.cuda_graph_dp = bool(params.get("cuda_graph_dp", False))
        # iter_size == 1, FusedAdamW: the AdamW + EMA update of the readouts (95 % of the parameters) can be issued on a
        # side stream as soon as their gradients exist, to run under the rest of backward.  Bit-identical, but measured
        # no gain on B200 (26.18 vs 26.12 ms): the backward kernels hold the whole register file (2 CTAs x 128 registers
        # per thread per SM), so the update's CTAs only become resident in the gaps.  Off by default.
        self.early_readout_step = bool(params.get("early_readout_step", False))
        self._opt_stream = None
        self._early_ev = None
        self._graphs: dict = {}
        self._graph_seen: dict = {}
        self._graph_pool = None
        self.max_graphs = 12

    # argus_models.py:31-41
    @torch.no_grad()
    def add_distill_predictions(self, input, target):
        if self.distill_model is not None and self.distill_ratio:
            # the teacher is frozen by contract (eval, no_grad, never stepped): its folded BatchNorm tables stay valid
            self.distill_model._dwn_frozen = True
            teacher = self.distill_model(input)
            target_tensors, mice_weights = target
            dev = mice_weights.device
            st = torch.cuda.current_stream(dev).cuda_stream
            B, nm = mice_weights.shape
            mask = torch.empty((B, nm), dtype=torch.uint8, device=dev)
            dweight = torch.empty((1,), dtype=torch.float32, device=dev)
            call("dwn_distill_prepare", mice_weights, B * nm, float(self.distill_ratio), mask, dweight, st)
            for m in range(nm):
                t = target_tensors[m]
                call("dwn_distill_fill", t, teacher[m].float().contiguous(), mask, nm, m, B, t.numel() // B, st)
            call("dwn_distill_weights", mice_weights, mask, dweight, B * nm, st)

    def _assemble_input(self, x):
        """The network input from either batch form: the reference's dense clip tensor (B, 5, T, H, W), or the raw form
        ``(video (B, T, Hv, Wv) uint8 / fp32, scalars (B, 4, T) fp32)`` — unpadded frames plus behaviour / pupil-centre
        values per frame — which is padded and stacked on the device exactly like StackInputsProcessor (inputs.py:22-36,
        ``params["inputs_processor"]`` gives the canvas size and fill value).  1.2 MB instead of 42 MB of H2D per batch."""
        if torch.is_tensor(x):
            return x
        video, scalars = x
        if not video.is_cuda:
            raise RuntimeError("raw clips must be on the CUDA device before assembly")
        name, kw = self.params.get("inputs_processor", ("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0}))
        if name != "stack_inputs":
            raise ValueError(f"raw clip batches need the 'stack_inputs' processor, got '{name}'")
        W, H = kw["size"]
        if video.dtype not in (torch.uint8, torch.float32):
            video = video.float()
        video = video.contiguous()
        scalars = scalars.float().contiguous()
        B, T, Hv, Wv = video.shape
        out = torch.empty((B, 5, T, H, W), dtype=torch.float32, device=video.device)
        call("dwn_assemble_batch", video, 2 if video.dtype == torch.uint8 else 0, scalars, out, B, T, Hv, Wv, H, W,
             float(kw.get("pad_fill_value", 0.0)), torch.cuda.current_stream(video.device).cuda_stream)
        return out

    def _to_device_overlapped(self, chunk_batch):
        """deep_to(batch, device, non_blocking=True) (argus_models.py:49) with the target / weight copies (80 % of the
        host->device bytes, not needed before the loss) issued on a side stream so they overlap the forward pass.
        Returns (input, target, ready) where ready() makes the compute stream wait for the side-stream copies."""
        x, target = chunk_batch
        dev = self.device
        if dev.type != "cuda" or not torch.is_tensor(x) or x.is_cuda:  # already resident (DevicePrefetcher) or no GPU
            inp, tgt = deep_to(chunk_batch, dev, non_blocking=True)
            return inp, tgt, (lambda: None)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        inp = x.to(dev, non_blocking=True)
        side = self._copy_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tgt = deep_to(target, dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)

        def _rec(o):
            if torch.is_tensor(o):
                o.record_stream(main)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    _rec(v)

        _rec(tgt)
        state = {"done": False}

        def ready():
            if not state["done"]:
                main.wait_event(ev)
                state["done"] = True

        return inp, tgt, ready

    # ---------------------------------------------------------------------------------------------------------
    # CUDA-graph replay of the train step
    # ---------------------------------------------------------------------------------------------------------
    def _graph_key(self, batch, pre_hint=None):
        """Key of the captured step this batch can replay, or None when the step must run eagerly."""
        if not (self.cuda_graph and self.iter_size == 1 and self.device.type == "cuda"
                and isinstance(self.loss, MicePoissonLoss) and isinstance(self.optimizer, FusedAdamW)
                and (getattr(self.nn_module, "_dp", None) is None or self.cuda_graph_dp)
                and getattr(self.nn_module, "_rng_device", None) is None):
            return None
        try:
            x, (t, w) = batch
        except (TypeError, ValueError):
            return None
        raw = not torch.is_tensor(x)
        if raw and not (isinstance(x, (list, tuple)) and len(x) == 2 and all(torch.is_tensor(v) for v in x)):
            return None
        if not torch.is_tensor(w):
            return None
        compact = torch.is_tensor(t)
        n_mice = len(self.nn_module.cfg["readout_outputs"])
        distill = bool(self.distill_model is not None and self.distill_ratio)
        if distill:
            live = (True,) * n_mice
        elif w.is_cuda:
            if pre_hint is None:                   # device-resident batch: DevicePrefetcher registers the hint
                return None                        # the set of mice present must be known on the host
            live = tuple(bool(v) for v in pre_hint)
        elif compact:
            present = set(w.tolist())
            live = tuple(m in present for m in range(n_mice))
        else:
            live = tuple((w != 0).any(0).tolist())
        tshape = tuple(t.shape) if compact else tuple(tuple(v.shape) for v in t)
        xshape = tuple((tuple(v.shape), str(v.dtype)) for v in x) if raw else (tuple(x.shape), str(x.dtype))
        return (xshape, tuple(w.shape), str(w.dtype), compact, tshape, live, distill, self.amp,
                self.model_ema is not None)

    # ---------------------------------------------------------------------------------------------------------
    # early optimizer step of the readouts
    # ---------------------------------------------------------------------------------------------------------
    def _early_ok(self) -> bool:
        return (self.early_readout_step and self.iter_size == 1 and isinstance(self.optimizer, FusedAdamW)
                and not self.grad_scaler.is_enabled() and getattr(self.nn_module, "_dp", None) is None
                and self.device.type == "cuda")

    def _early_readout_step(self, grads, params) -> None:
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=dev)
        ev = torch.cuda.Event()
        ev.record(main)
        self._opt_stream.wait_event(ev)
        with torch.cuda.stream(self._opt_stream):
            self.optimizer.early_step(params, grads)
            if self.model_ema is not None:
                self.model_ema.update(self.nn_module, part="readouts")
        self._early_ev = torch.cuda.Event()
        self._early_ev.record(self._opt_stream)

    def _backward_step_ema(self, loss) -> None:
        """backward + optimizer step + EMA (argus_models.py:55-62) with the readout update overlapped."""
        early = self._early_ok()
        self._early_ev = None
        if early:
            self.nn_module._early_step = self._early_readout_step   # transient: never visible to deepcopy (ModelEma)
        try:
            self.grad_scaler.scale(loss).backward()
        finally:
            if early:
                del self.nn_module._early_step
        if self._early_ev is not None:
            torch.cuda.current_stream(self.device).wait_event(self._early_ev)
        self.grad_scaler.step(self.optimizer)
        self.grad_scaler.update()
        if self.model_ema is not None:
            self.model_ema.update(self.nn_module, part="rest" if self._early_ev is not None else "all")
        self._early_ev = None

    def release_graphs(self) -> None:
        """Drop every captured step (graph memory pool, static buffers).  Call before destroying the process group when
        data-parallel steps were captured: the graphs hold NCCL kernels (also registered with atexit)."""
        self._graphs.clear()
        self._graph_last = None

    def _capture_step(self, key, batch):
        from . import _lib
        if not self._graphs:
            import atexit
            import weakref
            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref().release_graphs())
        x, (t, w) = batch
        dev = self.device
        compact, live = key[3], key[5]
        ent = State()
        ent.x = torch.empty(x.shape, dtype=x.dtype, device=dev) if torch.is_tensor(x) else [
            torch.empty(v.shape, dtype=v.dtype, device=dev) for v in x]
        ent.t = torch.empty(t.shape, dtype=t.dtype, device=dev) if compact else [
            torch.empty(v.shape, dtype=v.dtype, device=dev) for v in t]
        ent.w = torch.empty(w.shape, dtype=w.dtype, device=dev)
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        self.optimizer.zero_grad(set_to_none=True)
        params = [p for p in self.nn_module.parameters()]
        launches0 = _lib.LAUNCHES
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self._graph_pool, capture_error_mode="relaxed"):
            target = (ent.t, ent.w)
            self.loss.set_live_hint(list(live))
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                if compact:
                    target = collate_on_device(ent.t, ent.w, self.nn_module.cfg["readout_outputs"])
                xin = self._assemble_input(ent.x)
                self.add_distill_predictions(xin, target)
                prediction = self.nn_module(xin)
                loss = self.loss(prediction, target)
            self._backward_step_ema(loss)
        ent.graph = graph
        ent.launches = _lib.LAUNCHES - launches0
        ent.loss = loss.detach()
        ent.prediction = deep_detach(prediction)
        ent.target = deep_detach(target)
        ent.grads = [p.grad for p in params]     # keeps the graph's gradient buffers alive; restored after a replay
        ent.params = params
        return ent

    def _replay_step(self, ent, batch, sync: bool = True) -> dict:
        from . import _lib
        from .engine import bump_generation
        x, (t, w) = batch
        if torch.is_tensor(x):
            ent.x.copy_(x, non_blocking=True)
        else:
            for d, s_ in zip(ent.x, x):
                d.copy_(s_, non_blocking=True)
        if torch.is_tensor(t):
            ent.t.copy_(t, non_blocking=True)
        else:
            for d, s_ in zip(ent.t, t):
                d.copy_(s_, non_blocking=True)
        ent.w.copy_(w, non_blocking=True)
        self.optimizer.sync_graph_lr()
        ent.graph.replay()
        _lib.LAUNCHES += ent.launches
        bump_generation()                        # weights / running statistics / EMA changed behind torch's back
        if self._graph_last is not ent:
            for p, g in zip(ent.params, ent.grads):
                p.grad = g
            self._graph_last = ent
        prediction = self.prediction_transform(ent.prediction)
        return {"prediction": prediction, "target": ent.target, "loss": ent.loss.item() if sync else ent.loss}

    _graph_last = None

    def train_step_async(self, batch, state: State = None) -> dict:
        """train_step without the device->host read of the loss: ``loss`` is returned as a 0-dim device tensor, so the
        host can enqueue the next step while this one runs (an extension; the reference's train_step always syncs)."""
        return self.train_step(batch, state, _sync=False)

    # argus_models.py:43-71
    def train_step(self, batch, state: State, _sync: bool = True) -> dict:
        self.train()
        try:  # hint registered by DevicePrefetcher for this (device-resident) batch; looked up BEFORE deep_chunk, which
            pre_hint = _LIVE_HINTS.pop(id(batch[1][1]), None)  # returns new tensor objects
        except (TypeError, IndexError, KeyError):
            pre_hint = None
        key = self._graph_key(batch, pre_hint)
        if key is not None:
            ent = self._graphs.get(key)
            if ent is None:
                seen = self._graph_seen.get(key, 0)
                self._graph_seen[key] = seen + 1
                if seen >= 1 and len(self._graphs) < self.max_graphs:   # first occurrence runs eagerly (warm-up)
                    ent = self._graphs[key] = self._capture_step(key, batch)
            if ent is not None:
                return self._replay_step(ent, batch, _sync)
            self._graph_last = None
        self.optimizer.zero_grad()
        chunk_losses = []
        for i, chunk_batch in enumerate(deep_chunk(batch, self.iter_size)):
            host_t, host_w = chunk_batch[1]
            # compact batch form (an extension of datasets.py:172-187): target = (responses (B, n_max, T) of each sample's
            # own mouse, mouse_ids (B,) integer) instead of ten mostly-zero tensors + one-hot weights; it is scattered
            # into the reference's dense form on the device, 17 MB instead of 160 MB of H2D per batch
            compact = torch.is_tensor(host_t)
            n_mice = len(self.nn_module.cfg["readout_outputs"])
            distill = self.distill_model is not None and self.distill_ratio
            if isinstance(self.loss, MicePoissonLoss):
                if distill:
                    self.loss.set_live_hint([True] * n_mice)
                elif compact and not host_w.is_cuda:
                    present = set(host_w.tolist())
                    self.loss.set_live_hint([m in present for m in range(n_mice)])
                elif not compact and not host_w.is_cuda:
                    self.loss.set_live_hint((host_w != 0).any(0).tolist())
                elif pre_hint is not None and self.iter_size == 1:  # batch prefetched by DevicePrefetcher
                    self.loss.set_live_hint(pre_hint)
            input, target, ready = self._to_device_overlapped(chunk_batch)
            input = self._assemble_input(input)

            def dense():
                nonlocal target, compact
                ready()
                if compact:
                    target = collate_on_device(target[0], target[1], self.nn_module.cfg["readout_outputs"])
                    compact = False

            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                if distill:
                    dense()
                self.add_distill_predictions(input, target)
                prediction = self.nn_module(input)
                dense()  # targets / weights are only needed from here on
                loss = self.loss(prediction, target)
                loss = loss / self.iter_size
            if self.iter_size == 1:
                self._backward_step_ema(loss)
            else:
                self.grad_scaler.scale(loss).backward()
            chunk_losses.append(loss.detach())
        if self.iter_size != 1:
            self.grad_scaler.step(self.optimizer)
            self.grad_scaler.update()
            if self.model_ema is not None:
                self.model_ema.update(self.nn_module)
        # the reference reads loss.item() right after each backward (argus_models.py:56); reading the same values once
        # the optimizer and EMA kernels are enqueued returns the same number without idling the GPU at the sync
        if _sync:
            loss_value = 0
            for l in chunk_losses:
                loss_value += l.item()
        else:
            loss_value = torch.stack(chunk_losses).sum()
        prediction = deep_detach(prediction)
        target = deep_detach(target)
        prediction = self.prediction_transform(prediction)
        return {"prediction": prediction, "target": target, "loss": loss_value}

    # argus_models.py:73-87
    def val_step(self, batch, state: State) -> dict:
        self.eval()
        with torch.no_grad():
            input, target = deep_to(batch, device=self.device, non_blocking=True)
            if self.model_ema is None:
                prediction = self.nn_module(input)
            else:
                prediction = self.model_ema.ema(input)
            loss = self.loss(prediction, target)
            prediction = self.prediction_transform(prediction)
            return {"prediction": prediction, "target": target, "loss": loss.item()}

    # argus_models.py:89-99
    def predict(self, input, mouse_index: int | None = None):
        self._check_predict_ready()
        with torch.no_grad():
            self.eval()
            input = deep_to(input, self.device)
            if self.model_ema is None:
                prediction = self.nn_module(input, mouse_index)
            else:
                prediction = self.model_ema.ema(input, mouse_index)
            prediction = self.prediction_transform(prediction)
            return prediction
