"""GPU diagnostic: how accurate is the bf16 pipeline compared with torch's own bf16 autocast of the oracle?
(both measured against the fp32 oracle)."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import dwiseneuro_oracle as O  # noqa: E402
from sensorium_b200 import DwiseNeuro, constants, engine  # noqa: E402
from sensorium_b200.utils import init_weights  # noqa: E402
from tests.shapes import TINY_KW, TINY_OUTS, TRUE_BATCH_KW  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def rms(a, b):
    return float(((a.double() - b.double()) ** 2).mean().sqrt() / ((b.double() ** 2).mean().sqrt() + 1e-30))


# ---------------- C1 eval ----------------
torch.manual_seed(0)
net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW)
init_weights(net)
net = net.to(dev).eval()
x = O.synthetic_clip(1, 16, 64, seed=0).to(dev)
cfg = O.make_cfg(constants.num_neurons, **TRUE_BATCH_KW)
sd = {k: v.detach() for k, v in net.state_dict().items()}
with torch.no_grad():
    ref = O.dwiseneuro_forward(x, sd, cfg, 0, False)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref16 = O.dwiseneuro_forward(x, sd, cfg, 0, False)
    net.precision = "bf16"
    mine16 = net(x, 0)
    net.precision = "fp32"
    mine32 = net(x, 0)
    print(f"C1 eval: torch-autocast-bf16 vs fp32: max {rel(ref16.float(), ref):.3e} rms {rms(ref16.float(), ref):.3e}")
    print(f"C1 eval: ours bf16 vs fp32:            max {rel(mine16, ref):.3e} rms {rms(mine16, ref):.3e}")
    print(f"C1 eval: ours fp32 vs fp32:            max {rel(mine32, ref):.3e}")
    # per-block trunk error
    _, s16 = engine.run_forward(net, x, 0, "bf16", False, True)
    _, s32 = engine.run_forward(net, x, 0, "fp32", False, True)
    for i, (a, b) in enumerate(zip(s16.blocks, s32.blocks)):
        print(f"  blk{i}: X max {rel(a.X, b.X):.3e} rms {rms(a.X, b.X):.3e} | E rms {rms(a.E.float(), b.E):.3e} S rms {rms(a.S.float(), b.S):.3e} "
              f"Tm rms {rms(a.Tm.float(), b.Tm):.3e} A rms {rms(a.A.float(), b.A):.3e} gate rms {rms(a.gate, b.gate):.3e} Y rms {rms(a.Y.float(), b.Y):.3e}")
    print(f"  cortex in rms {rms(s16.cortex[0].x, s32.cortex[0].x):.3e}; cortex out rms {rms(s16.cx, s32.cx):.3e}")
del net

# ---------------- tiny train: gradient accuracy yardstick ----------------
for B, T, HW, seed in ((4, 16, 32, 0), (2, 16, 32, 2)):
    torch.manual_seed(seed)
    net = DwiseNeuro(readout_outputs=TINY_OUTS, **TINY_KW)
    init_weights(net)
    net = net.to(dev).train()
    x = O.synthetic_clip(B, T, HW, seed=seed).to(dev)
    tg, w = O.synthetic_targets(B, TINY_OUTS, T, seed=seed + 1)
    tg, w = [t.to(dev) for t in tg], w.to(dev)
    cfg = O.make_cfg(TINY_OUTS, **TINY_KW)
    names = [k for k, _ in net.named_parameters()]

    def oracle_grads(autocast):
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        for k in names:
            sd[k].requires_grad_(True)
        torch.manual_seed(11)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = O.dwiseneuro_forward(x, sd, cfg, None, True)
            loss = O.mice_poisson_loss(out, tg, w)
        loss.backward()
        return {k: sd[k].grad for k in names}, [o.detach().float() for o in out]

    g32, o32 = oracle_grads(False)
    g16, o16 = oracle_grads(True)
    net.precision = "bf16"
    net._mask_dtype = torch.float32
    torch.manual_seed(11)
    out = net(x)
    loss = O.mice_poisson_loss(out, tg, w)
    loss.backward()
    gmax = max(float(v.abs().max()) for v in g32.values() if v is not None)
    print(f"tiny train B={B}: pred max-rel torch-bf16 {max(rel(a, b) for a, b in zip(o16, o32)):.3e} ours {max(rel(a.detach(), b) for a, b in zip(out, o32)):.3e}")
    rows = []
    for k, p in net.named_parameters():
        if g32[k] is None or float(g32[k].abs().max()) < 1e-3 * gmax:
            continue
        rows.append((rel(p.grad, g32[k]), rel(g16[k].float(), g32[k]), rms(p.grad, g32[k]), rms(g16[k].float(), g32[k]), k))
    rows.sort(reverse=True)
    for e_m, e_t, r_m, r_t, k in rows[:8]:
        print(f"   {k:45s} ours max {e_m:.3e} rms {r_m:.3e} | torch-bf16 max {e_t:.3e} rms {r_t:.3e}")
    import statistics
    print(f"   median max-rel: ours {statistics.median(r[0] for r in rows):.3e} torch-bf16 {statistics.median(r[1] for r in rows):.3e}")
