"""GPU diagnostic (not a pytest): stage-by-stage check of the forward kernels.  Each stage is recomputed with
torch ops from the *engine's own* previous intermediate, so the first failing kernel is pinpointed."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from sensorium_b200 import DwiseNeuro, engine  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
worst = {}


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def report(tag, a, b, tol):
    e = rel(a.float(), b.float())
    bad = e > tol or not torch.isfinite(a.float()).all()
    print(f"{'FAIL' if bad else 'ok  '} {tag}: rel={e:.3e}")
    worst[tag] = e
    return not bad


def cl2ncdhw(t, B, T, H, W):
    return t.float().view(B, T, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


def bn_train(x, bn):
    return F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.1, 1e-5)


def run(mode, cfgkw, B, T, HW, seed=0):
    tol = 2e-5 if mode == "fp32" else 3e-2
    torch.manual_seed(seed)
    net = DwiseNeuro(**cfgkw).to(dev)
    for p in net.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, 0, 0.5 / (p[0].numel() ** 0.5))
    for n_, p in net.named_parameters():
        if n_.endswith("bn.weight"):
            torch.nn.init.uniform_(p, 0.5, 1.5)
        elif n_.endswith("bias"):
            torch.nn.init.uniform_(p, -0.3, 0.3)
    net.train()
    x = torch.randn(B, cfgkw.get("in_channels", 5), T, HW, HW, device=dev) * 2 + 0.5
    torch.manual_seed(123)
    with torch.no_grad():
        preds, sv = engine.run_forward(net, x, None, mode, True, True)
    torch.cuda.synchronize()
    ok = True
    feats = net.cfg["core_features"]
    # stem
    stem = net.core.stem
    ref = bn_train(F.conv3d(x, stem[0].weight), stem[1].bn)
    pe = engine.pe_tables(net.core.blocks[0], feats[0], T, HW, HW, dev)
    ref = ref + (pe[0].t()[None, :, :, None, None] + pe[1].t()[None, :, None, :, None] + pe[2].t()[None, :, None, None, :])
    ok &= report(f"[{mode}] stem+pe", cl2ncdhw(sv.blocks[0].X, B, T, HW, HW), ref, 2e-5 if mode == "fp32" else 1e-4)
    for i, b in enumerate(sv.blocks):
        blk = net.core.blocks[2 * i + 1]
        Xn = cl2ncdhw(b.X, B, T, b.Hi, b.Wi)
        Xg = cl2ncdhw(b.Xb if b.Xb is not None else b.X, B, T, b.Hi, b.Wi)
        wq = (lambda w: w.to(torch.bfloat16).float()) if mode == "bf16" else (lambda w: w)
        E_ref = F.conv3d(Xg, wq(blk.conv_pw[0].weight))
        ok &= report(f"[{mode}] blk{i} E (pw gemm)", cl2ncdhw(b.E, B, T, b.Hi, b.Wi), E_ref, tol)
        Em = cl2ncdhw(b.E, B, T, b.Hi, b.Wi)
        ok &= report(f"[{mode}] blk{i} bn1 mean", b.coef1[2], Em.mean((0, 2, 3, 4)), 1e-3)
        a1 = F.silu(bn_train(Em, blk.conv_pw[1].bn))
        S_ref = F.conv3d(a1, blk.spat_covn_dw[0].weight, stride=(1, b.s, b.s), padding=(0, 1, 1), groups=b.mid)
        ok &= report(f"[{mode}] blk{i} S (spatial dw)", cl2ncdhw(b.S, B, T, b.Ho, b.Wo), S_ref, tol)
        Sm = cl2ncdhw(b.S, B, T, b.Ho, b.Wo)
        a2 = F.silu(bn_train(Sm, blk.spat_covn_dw[1].bn))
        T_ref = F.conv3d(a2, blk.temp_covn_dw[0].weight, padding=(2, 0, 0), groups=b.mid)
        ok &= report(f"[{mode}] blk{i} Tm (temporal dw)", cl2ncdhw(b.Tm, B, T, b.Ho, b.Wo), T_ref, tol)
        Tmm = cl2ncdhw(b.Tm, B, T, b.Ho, b.Wo)
        a3 = F.silu(bn_train(Tmm, blk.temp_covn_dw[1].bn))
        ok &= report(f"[{mode}] blk{i} A (bn3+silu)", cl2ncdhw(b.A, B, T, b.Ho, b.Wo), a3, tol)
        Am = cl2ncdhw(b.A, B, T, b.Ho, b.Wo)
        se = Am.mean((2, 3, 4), keepdim=True)
        se = F.silu(F.conv3d(se, blk.se.conv_reduce.weight, blk.se.conv_reduce.bias))
        se = torch.sigmoid(F.conv3d(se, blk.se.conv_expand.weight, blk.se.conv_expand.bias))
        ok &= report(f"[{mode}] blk{i} gate", b.gate, se.flatten(1), 1e-4 if mode == "fp32" else 1e-2)
        Y_ref = F.conv3d(Am * b.gate[:, :, None, None, None], blk.conv_pwl[0].weight)
        ok &= report(f"[{mode}] blk{i} Y (pwl gemm)", cl2ncdhw(b.Y, B, T, b.Ho, b.Wo), Y_ref, tol)
        Ym = cl2ncdhw(b.Y, B, T, b.Ho, b.Wo)
        y = bn_train(Ym, blk.conv_pwl[1].bn)
        if b.dp is not None:
            y = y * b.dp[:, None, None, None, None]
        sc = Xn
        if b.s > 1:
            sc = F.interpolate(sc, size=(T, b.Ho, b.Wo), mode="nearest")
        if b.ci != b.co:
            sc = torch.tile(sc, (1, -(-b.co // b.ci), 1, 1, 1))[:, :b.co]
        out_ref = y + bn_train(sc, blk.bn_sc.bn)
        if i + 1 < len(sv.blocks):
            pe = engine.pe_tables(net.core.blocks[2 * i + 2], b.co, T, b.Ho, b.Wo, dev)
            out_ref = out_ref + (pe[0].t()[None, :, :, None, None] + pe[1].t()[None, :, None, :, None] +
                                 pe[2].t()[None, :, None, None, :])
            got = cl2ncdhw(sv.blocks[i + 1].X, B, T, b.Ho, b.Wo)
            ok &= report(f"[{mode}] blk{i} out (residual+pe)", got, out_ref, 2e-5 if mode == "fp32" else 2e-2)
        else:
            pooled = out_ref.mean((3, 4))  # (B, C, T)
            got = sv.cortex[0].x.view(B, T, -1).permute(0, 2, 1)
            ok &= report(f"[{mode}] blk{i} out+pool", got, pooled, 2e-5 if mode == "fp32" else 2e-2)
    # cortex
    G = net.cfg["groups"]
    for j, c in enumerate(sv.cortex):
        layer = net.cortex.layers[j]
        xin = c.x.view(B, T, -1).permute(0, 2, 1).contiguous()
        xg = (c.xb if c.xb is not None else c.x).float().view(B, T, -1).permute(0, 2, 1).contiguous()
        w = layer.conv.weight
        if mode == "bf16":
            w = w.to(torch.bfloat16).float()
        y = F.conv1d(xg, w, groups=G)
        ok &= report(f"[{mode}] cortex{j} Y (grouped gemm)", c.Y.float().view(B, T, -1).permute(0, 2, 1), y, tol)
        ym = c.Y.float().view(B, T, -1).permute(0, 2, 1).contiguous()
        y = F.silu(bn_train(ym, layer.bn.bn))
        y = y.view(B, G, -1, T).transpose(1, 2).reshape(B, -1, T)
        if c.dp is not None:
            y = y * c.dp[:, None, None]
        sc = xin
        if c.I != c.O:
            sc = torch.tile(sc, (1, -(-c.O // c.I), 1))[:, :c.O]
        ref = y + bn_train(sc, layer.bn_sc.bn)
        nxt = sv.cortex[j + 1].x if j + 1 < len(sv.cortex) else sv.cx
        ok &= report(f"[{mode}] cortex{j} out", nxt.view(B, T, -1).permute(0, 2, 1), ref, 2e-5 if mode == "fp32" else 2e-2)
    # readouts
    for r in sv.readouts:
        conv = net.readouts[r.m].layer[1]
        xin = sv.cx.view(B, T, -1).permute(0, 2, 1)
        if r.mask is not None:
            xin = xin * r.mask
        w = conv.weight
        if mode == "bf16":
            w = w.to(torch.bfloat16).float()
            xin = xin.to(torch.bfloat16).float()
        ref = F.softplus(F.conv1d(xin, w, conv.bias, groups=G)[:, :r.n_out], beta=net.cfg["softplus_beta"])
        ok &= report(f"[{mode}] readout{r.m}", r.pred, ref, 2e-5 if mode == "fp32" else 2e-2)
        if r.xt is not None:
            ok &= report(f"[{mode}] readout{r.m} xT", r.xt.float(), r.xm.float().t(), 0.0 + 1e-9)
    return ok


small = dict(readout_outputs=(37, 64, 129), core_features=(16, 16, 32), spatial_strides=(2, 1, 2), expansion_ratio=4,
             se_reduce_ratio=8, cortex_features=(64, 128), groups=2, drop_path_rate=0.3)
allok = True
for mode in ("fp32", "bf16"):
    allok &= run(mode, small, B=4, T=16, HW=32)
    allok &= run(mode, small, B=2, T=8, HW=16, seed=1)
print("LAYER CHECK", "PASSED" if allok else "FAILED")
bad = {k: v for k, v in worst.items() if v > 3e-2}
print("worst offenders:", bad)
sys.exit(0 if allok else 1)
