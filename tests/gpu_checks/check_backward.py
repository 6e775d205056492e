"""GPU diagnostic (not a pytest): parameter gradients of the CUDA path vs torch autograd through the oracle."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import dwiseneuro_oracle as O  # noqa: E402
from sensorium_b200 import DwiseNeuro  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def run(mode, cfgkw, B, T, HW, seed=0):
    torch.manual_seed(seed)
    net = DwiseNeuro(**cfgkw).to(dev)
    for n_, p in net.named_parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, 0, 0.7 / (p[0].numel() ** 0.5))
        elif n_.endswith("bn.weight"):
            torch.nn.init.uniform_(p, 0.5, 1.5)
        else:
            torch.nn.init.uniform_(p, -0.3, 0.3)
    net.train()
    net.precision = mode
    outs_n = cfgkw["readout_outputs"]
    x = O.synthetic_clip(B, T, HW, seed=seed).to(dev)
    tg, w = O.synthetic_targets(B, outs_n, T, seed=seed + 1)
    tg = [t.to(dev) for t in tg]
    w = w.to(dev)
    # oracle on a detached copy of the state
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    for k in names:
        sd[k].requires_grad_(True)
    cfg = O.make_cfg(outs_n, **{k: v for k, v in cfgkw.items() if k != "readout_outputs"})
    torch.manual_seed(77)
    ref_out = O.dwiseneuro_forward(x, sd, cfg, None, True)
    ref_loss = O.mice_poisson_loss(ref_out, tg, w)
    ref_loss.backward()
    torch.manual_seed(77)
    out = net(x)
    loss = O.mice_poisson_loss(out, tg, w)
    loss.backward()
    torch.cuda.synchronize()
    tol = 2e-4 if mode == "fp32" else 6e-2
    ok = True
    # biases that feed a batch-stat BatchNorm have an analytically zero gradient: compare those (and
    # everything else) with an absolute floor relative to the largest gradient of the network
    gscale = max(float(sd[k].grad.abs().max()) for k in names if sd[k].grad is not None)
    atol = (1e-6 if mode == "fp32" else 2e-3) * gscale
    print(f"[{mode}] loss {float(loss):.6f} ref {float(ref_loss):.6f}  pred rel {max(rel(a, b) for a, b in zip(out, ref_out)):.3e}")
    worst = []
    for (k, p) in net.named_parameters():
        g_ref = sd[k].grad
        if g_ref is None:
            if p.grad is not None:
                print(f"FAIL {k}: grad should be None")
                ok = False
            continue
        if p.grad is None:
            print(f"FAIL {k}: grad missing")
            ok = False
            continue
        e = rel(p.grad, g_ref)
        adiff = float((p.grad - g_ref).abs().max())
        if adiff <= atol:
            continue
        worst.append((e, k))
        if e > tol or not torch.isfinite(p.grad).all():
            ok = False
            print(f"FAIL [{mode}] {k}: rel={e:.3e} |ref|max={float(g_ref.abs().max()):.3e}")
    worst.sort(reverse=True)
    print(f"[{mode}] worst grads:", [(f"{e:.2e}", k) for e, k in worst[:6]])
    # running stats parity
    es = max(rel(net.state_dict()[k].float(), sd[k].float()) for k in sd if "running" in k)
    print(f"[{mode}] running-stat rel err {es:.3e}")
    return ok


small = dict(readout_outputs=(37, 64, 129), core_features=(16, 16, 32), spatial_strides=(2, 1, 2), expansion_ratio=4,
             se_reduce_ratio=8, cortex_features=(64, 128), groups=2, drop_path_rate=0.3)
allok = True
for mode in ("fp32", "bf16"):
    allok &= run(mode, small, B=4, T=16, HW=32)
    allok &= run(mode, small, B=3, T=8, HW=16, seed=1)
print("BACKWARD CHECK", "PASSED" if allok else "FAILED")
sys.exit(0 if allok else 1)
