"""GPU check (not a pytest): per-parameter bf16 gradient errors of the tiny train step of
tests/test_parity_gpu.py::test_train_step_vs_oracle_shared_rng, printed for the sdw_bwd kernel families (DWN_SDW_TMA=0/1/2)
so that a borderline bound can be told from a wrong kernel.  Usage: python tests/gpu_checks/check_tiny_grads.py"""
import math
import os
import sys

import torch

sys.path.insert(0, ".")
import oracle.dwiseneuro_oracle as O  # noqa: E402
from tests.shapes import TINY_KW, TINY_OUTS  # noqa: E402
from tests.test_parity_gpu import _tiny  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for (B, T, HW, seed) in [(3, 8, 16, 1), (2, 16, 32, 2)]:
    ref_cache = None
    for mode in ("0", "1", "2"):
        os.environ["DWN_SDW_TMA"] = mode
        net = _tiny(dev, seed)
        for n_, p in net.named_parameters():
            if p.dim() == 1:
                torch.nn.init.uniform_(p, 0.5, 1.5) if n_.endswith("bn.weight") else torch.nn.init.uniform_(p, -0.3, 0.3)
        net.train()
        net.precision = "bf16"
        net._mask_dtype = torch.float32
        x = O.synthetic_clip(B, T, HW, seed=seed).to(dev)
        tg, w = O.synthetic_targets(B, TINY_OUTS, T, seed=seed + 1)
        tg, w = [t.to(dev) for t in tg], w.to(dev)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        names = [k for k, _ in net.named_parameters()]
        for k in names:
            sd[k].requires_grad_(True)
        cfg = O.make_cfg(TINY_OUTS, **TINY_KW)
        torch.manual_seed(11)
        ref = O.dwiseneuro_forward(x, sd, cfg, None, True)
        O.mice_poisson_loss(ref, tg, w).backward()
        torch.manual_seed(11)
        out = net(x)
        O.mice_poisson_loss(out, tg, w).backward()
        gmax = max(float(sd[k].grad.abs().max()) for k in names if sd[k].grad is not None)
        rows = []
        for k, p in net.named_parameters():
            if sd[k].grad is None:
                continue
            gr = sd[k].grad
            err = float((p.grad - gr).abs().max()) / max(float(gr.abs().max()), 3e-3 * gmax)
            l2 = float((p.grad.double() - gr.double()).norm()) / max(float(gr.double().norm()), 3e-3 * gmax * math.sqrt(gr.numel()))
            rows.append((err, l2, k))
        rows.sort(reverse=True)
        print(f"B={B} T={T} HW={HW} DWN_SDW_TMA={mode}: worst max-norm " + "  ".join(f"{k}:{e:.3f}/{l:.3f}" for e, l, k in rows[:4]), flush=True)
