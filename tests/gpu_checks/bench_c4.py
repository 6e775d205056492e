"""C4: distillation train step — expansion-6 student, frozen expansion-7 teacher in eval mode,
distill_ratio 0.36, drop-path 0.1, dropout 0.4, EMA on — through MouseModel.train_step with pinned host batches.
CUDA-event timing after warm-up; not a pytest."""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import dwiseneuro_oracle as O  # noqa: E402  (synthetic inputs only)
from sensorium_b200 import DwiseNeuro, constants  # noqa: E402
from sensorium_b200.argus_models import MouseModel  # noqa: E402
from sensorium_b200.ema import ModelEma  # noqa: E402
from sensorium_b200.utils import init_weights  # noqa: E402
from tests.shapes import TRUE_BATCH_KW  # noqa: E402

dev = torch.device("cuda:0")
B = 32
kw_student = dict(TRUE_BATCH_KW, expansion_ratio=6)
params = {"nn_module": ("dwiseneuro", {"readout_outputs": constants.num_neurons, **kw_student}),
          "loss": ("mice_poisson", {"log_input": False, "full": False, "eps": 1e-8}),
          "optimizer": ("FusedAdamW", {"lr": 2.4e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
          "iter_size": 1}
torch.manual_seed(0)
model = MouseModel(params)
init_weights(model.nn_module)
teacher = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW).to(dev)
init_weights(teacher)
teacher.eval()
for p in teacher.parameters():
    p.requires_grad_(False)
model.distill_model, model.distill_ratio = teacher, 0.36
model.model_ema = ModelEma(model.nn_module, decay=0.999)
x = O.synthetic_clip(B, 16, 64, seed=1000)
tg, w = O.synthetic_targets(B, constants.num_neurons, 16, seed=1001)
x, tg, w = x.pin_memory(), [t.pin_memory() for t in tg], w.pin_memory()


def step():
    # distillation mutates targets / weights in place on the device copy; the host batch stays pristine
    return model.train_step((x, (tg, w)), None)["loss"]


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 8
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out = {"config": "C4 distillation: student er=6 + frozen teacher er=7 (eval), distill_ratio 0.36, EMA 0.999, batch 32, bf16",
       "ms_per_step_e2e": ms, "clips_per_s_e2e": B / ms * 1e3, "last_loss": loss}
print(json.dumps(out))
