"""GPU timing (not a pytest): the torch-eager restatement of the reference (oracle, cuDNN/cuBLAS kernels, bf16 autocast)
on the SAME B200, next to this repo's kernels — SURVEY.md §8(d) "the real bar to beat".  Also the C5 numbers: eval
forward windows/s and Predictor.predict_trial on a 300-frame trial.  CUDA-event timing after warm-up."""
import json
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import dwiseneuro_oracle as O  # noqa: E402
from sensorium_b200 import DwiseNeuro, constants  # noqa: E402
from sensorium_b200.utils import init_weights  # noqa: E402
from tests.shapes import TRUE_BATCH_KW  # noqa: E402

dev = torch.device("cuda:0")
B = 32
out = {}


def timed(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW)
init_weights(net)
net = net.to(dev)
cfg = O.make_cfg(constants.num_neurons, **TRUE_BATCH_KW)
x = O.synthetic_clip(B, 16, 64, seed=1000).to(dev)
tg, w = O.synthetic_targets(B, constants.num_neurons, 16, seed=1001)
tg, w = [t.to(dev) for t in tg], w.to(dev)

# ---- torch eager train step: fwd + loss + bwd + AdamW (fused) under bf16 autocast
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
names = [k for k, _ in net.named_parameters()]
leaves = []
for k in names:
    sd[k].requires_grad_(True)
    leaves.append(sd[k])
opt = torch.optim.AdamW(leaves, lr=2.4e-3, weight_decay=0.05, fused=True)


def eager_step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred = O.dwiseneuro_forward(x, sd, cfg, None, True)
        loss = O.mice_poisson_loss(pred, tg, w)
    loss.backward()
    opt.step()


torch.cuda.reset_peak_memory_stats()
ms = timed(eager_step, 2, 5)
out["eager_train_bf16_autocast"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3,
                                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
print("torch eager train step (bf16 autocast, fused AdamW): %.1f ms  %.1f clips/s  peak %.1f GB" %
      (ms, B / ms * 1e3, out["eager_train_bf16_autocast"]["peak_mem_gb"]), flush=True)
del opt, leaves
sd = {k: v.detach() for k, v in net.state_dict().items()}
torch.cuda.empty_cache()

# ---- eval forward, one mouse (what Predictor runs per batch of 32 windows)
for mode, ctx in (("fp32", None), ("bf16", torch.bfloat16)):
    def eager_eval():
        with torch.no_grad():
            if ctx is None:
                return O.dwiseneuro_forward(x, sd, cfg, 0, False)
            with torch.autocast("cuda", dtype=ctx):
                return O.dwiseneuro_forward(x, sd, cfg, 0, False)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ms = timed(eager_eval, 2, 5)
    out[f"eager_eval_{mode}"] = {"ms_per_batch": ms, "windows_per_s": B / ms * 1e3}
    print(f"torch eager eval fwd {mode} (index 0): {ms:.2f} ms  {B / ms * 1e3:.0f} windows/s", flush=True)

net.eval()
for mode in ("fp32", "bf16"):
    net.precision = mode

    def ours_eval():
        with torch.no_grad():
            return net(x, 0)
    ms = timed(ours_eval, 3, 10)
    out[f"ours_eval_{mode}"] = {"ms_per_batch": ms, "windows_per_s": B / ms * 1e3}
    print(f"ours eval fwd {mode} (index 0): {ms:.2f} ms  {B / ms * 1e3:.0f} windows/s", flush=True)

# ---- our train step for the same batch (device-resident), for the ratio
from sensorium_b200.losses import MicePoissonLoss  # noqa: E402
from sensorium_b200.optim import FusedAdamW  # noqa: E402
net.train()
net.precision = "auto"
loss_fn = MicePoissonLoss()
fopt = FusedAdamW(net.parameters(), lr=2.4e-3, weight_decay=0.05)


def ours_step():
    fopt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(net(x), (tg, w))
    loss.backward()
    fopt.step()


ms = timed(ours_step, 3, 10)
out["ours_train_bf16"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
print("ours train step (bf16, FusedAdamW, no EMA): %.1f ms  %.1f clips/s" % (ms, B / ms * 1e3), flush=True)

# ---- Predictor on one 300-frame trial (C5): 270 windows, one mouse
from sensorium_b200.argus_models import MouseModel  # noqa: E402
from sensorium_b200.predictors import Predictor  # noqa: E402
del fopt
torch.cuda.empty_cache()
params = {"nn_module": ("dwiseneuro", {"readout_outputs": constants.num_neurons, **TRUE_BATCH_KW}),
          "loss": ("mice_poisson", {"log_input": False, "full": False, "eps": 1e-8}),
          "optimizer": ("FusedAdamW", {"lr": 2.4e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
          "iter_size": 1, "inputs_processor": ("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0}),
          "responses_processor": ("identity", {}), "frame_stack": {"size": 16, "step": 2, "position": "last"}}
model = MouseModel(params)
model.nn_module.load_state_dict(net.state_dict())
rng = np.random.default_rng(0)
video = rng.integers(0, 256, size=(36, 64, 300)).astype(np.float32)
beh = rng.random((2, 300), dtype=np.float32) * 10
pup = rng.random((2, 300), dtype=np.float32) * 30
with tempfile.TemporaryDirectory() as d:
    path = d + "/model.pth"
    model.save(path)
    del model
    for mode in ("fp32", "bf16"):
        pr = Predictor(path, device="cuda:0", blend_weights="ones", precision=mode)
        pr.predict_trial(video, beh, pup, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for i in range(n):
            res = pr.predict_trial(video, beh, pup, i % 5)
        dt = (time.perf_counter() - t0) / n
        out[f"predict_trial_{mode}"] = {"s_per_trial": dt, "windows_per_s": 270 / dt}
        print(f"Predictor.predict_trial {mode}: {dt * 1e3:.1f} ms per 300-frame trial = {270 / dt:.0f} windows/s "
              f"(host in, host out)", flush=True)
        del pr
print(json.dumps(out))
