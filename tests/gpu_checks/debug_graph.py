import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import dwiseneuro_oracle as O
from tests.shapes import TINY_KW, TINY_OUTS
from sensorium_b200.argus_models import MouseModel
from sensorium_b200.ema import ModelEma
from sensorium_b200.utils import init_weights

kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)

def batches():
    out = []
    for i in range(6):
        x = O.synthetic_clip(4, 16, 32, seed=60 + i % 3)
        tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=70 + i % 3)
        if i % 2 == 1:
            keep = (w[:, 2] == 0).float()
            w = w * keep[:, None]
            w[:, 0] = torch.clamp(w[:, 0] + (1 - keep), max=1.0)
            tg[2].zero_()
        out.append((x, (tg, w)))
    return out

def make(graph):
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw}), "loss": ("mice_poisson", {}),
              "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
              "iter_size": 1, "cuda_graph": graph}
    torch.manual_seed(0)
    m = MouseModel(params)
    init_weights(m.nn_module)
    m.model_ema = ModelEma(m.nn_module, decay=0.9)
    return m

m1, m2 = make(False), make(True)
for i, b in enumerate(batches()):
    l1 = m1.train_step(b, None)["loss"]
    l2 = m2.train_step(b, None)["loss"]
    torch.cuda.synchronize()
    bad = [k for (k, a), c in zip(m1.nn_module.state_dict().items(), m2.nn_module.state_dict().values()) if not torch.equal(a, c)]
    gbad = [k for (k, p1), p2 in zip(m1.nn_module.named_parameters(), m2.nn_module.parameters())
            if (p1.grad is None) != (p2.grad is None) or (p1.grad is not None and not torch.equal(p1.grad, p2.grad))]
    print(f"step {i}: loss {l1} {l2} graphs={len(m2._graphs)} state mismatches {len(bad)} {bad[:6]} grad mismatches {len(gbad)} {gbad[:6]}")
