"""GPU parity at the BENCHMARKED architecture (-m gpu): true_batch_001 (expansion 7, 64x64 clips, 10 readouts with
the real neuron counts) in TRAIN mode — forward, MicePoissonLoss, backward, BatchNorm running statistics — against the
fp32 oracle run on the same GPU (TF32 off) with the same RNG seed (same drop-path / dropout masks).

Bounds (BASELINE.json north_star).  Every bound below is an absolute number; nothing is relaxed against a yardstick:
  fp32 mode: outputs / loss / running stats <= 1e-4 relative, every parameter gradient <= 1e-4 (max-norm, relative to
             the tensor's own largest gradient, with an absolute floor for gradients that are analytically zero);
  bf16 mode: predicted responses, loss and running statistics <= 2e-2 relative; the single-trial correlation METRIC
             (per-neuron corr averaged over the neurons of a mouse, metrics.py:66-70 — the number the reference
             reports) within 1e-3.
             NOT met, and stated as such: the worst single neuron's correlation moves by up to 1.6e-2 (bound asserted:
             CORR_NEURON_BF16).  With random-init weights the responses are almost constant in time, so a neuron's
             correlation measures the error relative to the response's *variation*, a few % of its value; 8 mantissa
             bits cannot hold that to 1e-3 (torch's own bf16 autocast of the reference moves it by 2.0-2.9e-2).
             Gradients ("checked on a fixed batch", no number in north_star): relative L2 error per tensor
             <= GRAD_L2_BF16, max-norm <= GRAD_MAX_BF16, median L2 over all tensors <= GRAD_L2_MEDIAN_BF16
             (measured 0.13 / 0.17 / 0.025 on C2 and 0.18 / 0.24 / 0.046 on the C4 distillation step, where all ten
             readouts are live and dropout / drop-path are on; torch autocast on C2: 0.23 / 0.41 / 0.044).
Every run writes a per-stage error table (block inputs, outputs, loss, every gradient; the same quantities for the
oracle under torch's bf16 autocast as a DIAGNOSTIC column, never as a bound) to gpurun_out/ — committed copies live in
profiles/ — so that a failure names the offending stage."""
import math
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from oracle import dwiseneuro_oracle as O
from tests.shapes import TINY_KW, TINY_OUTS, TRUE_BATCH_KW

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
FP32_TOL, BF16_TOL, CORR_TOL = 1e-4, 2e-2, 1e-3
CORR_NEURON_BF16 = 3e-2
GRAD_L2_BF16, GRAD_MAX_BF16, GRAD_L2_MEDIAN_BF16 = 0.2, 0.3, 6e-2
NUM_NEURONS = (7863, 7908, 8202, 7939, 8122, 7440, 7928, 8285, 7671, 7495)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _report(name, lines):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / name).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


def _net(dev, kw, outs, seed=0, perturb=True):
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.utils import init_weights
    torch.manual_seed(seed)
    net = DwiseNeuro(readout_outputs=outs, **kw)
    init_weights(net)
    if perturb:  # non-trivial BatchNorm affine parameters and biases
        for n_, p in net.named_parameters():
            if p.dim() == 1:
                torch.nn.init.uniform_(p, 0.5, 1.5) if n_.endswith("bn.weight") else torch.nn.init.uniform_(p, -0.3, 0.3)
    return net.to(dev)


def _capture_saved():
    """Monkeypatch hook: keeps the saved forward state of the next engine.run_forward(save=True) call."""
    from sensorium_b200 import engine
    box = []
    orig = engine.run_forward

    def wrapped(*a, **k):
        outs, sv = orig(*a, **k)
        if sv is not None:
            box.append(sv)
        return outs, sv

    engine.run_forward = wrapped
    return box, lambda: setattr(engine, "run_forward", orig)


def _corr_gap(pred, ref, target):
    n = pred.shape[1]
    f = lambda t: t.permute(1, 0, 2).reshape(n, -1)  # noqa: E731
    return float((O.corr(f(pred), f(target)) - O.corr(f(ref), f(target))).abs().max())


def _oracle_train(x, sd0, names, cfg, tg, w, seed, taps=None):
    sd = {k: v.detach().clone() for k, v in sd0.items()}
    for k in names:
        sd[k].requires_grad_(True)
    torch.manual_seed(seed)
    ref = O.dwiseneuro_forward(x, sd, cfg, None, True, taps=taps)
    loss = O.mice_poisson_loss(ref, tg, w)
    loss.backward()
    return ref, loss, sd


def _oracle_autocast(x, sd0, names, cfg, tg, w, seed, taps=None):
    """DIAGNOSTIC ONLY (never part of a bound): the oracle under torch's own bf16 autocast, i.e. what the reference's
    AMP path would compute in bf16 — printed next to our errors so a reader can tell inherent bf16 noise from defects."""
    sd = {k: v.detach().clone() for k, v in sd0.items()}
    for k in names:
        sd[k].requires_grad_(True)
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.dwiseneuro_forward(x, sd, cfg, None, True, taps=taps)
        loss = O.mice_poisson_loss(out, tg, w)
    loss.backward()
    return out, loss, sd


def _grad_table(net, sd_ref, lines, sd_yard=None):
    """Per-tensor gradient errors.  Returns (worst max-norm error, worst L2 error, list of offending names)."""
    names = [k for k, _ in net.named_parameters()]
    gmax = max(float(sd_ref[k].grad.abs().max()) for k in names if sd_ref[k].grad is not None)
    rows = []
    for k, p in net.named_parameters():
        gr = sd_ref[k].grad
        if gr is None:
            assert p.grad is None, f"{k}: the reference has no gradient (absent mouse) but the CUDA path produced one"
            continue
        assert p.grad is not None, k
        ref_max = float(gr.abs().max())
        # gradients of parameters in front of a batch-stat BatchNorm (conv biases, BN-normalised scales) are analytically
        # zero; the floor keeps pure round-off from being compared with itself
        e_max = float((p.grad - gr).abs().max()) / max(ref_max, 3e-3 * gmax)
        e_l2 = float((p.grad.double() - gr.double()).norm()) / max(float(gr.double().norm()),
                                                                  3e-3 * gmax * math.sqrt(gr.numel()))
        y_max = y_l2 = float("nan")
        if sd_yard is not None and sd_yard[k].grad is not None:
            gy = sd_yard[k].grad.float()
            y_max = float((gy - gr).abs().max()) / max(ref_max, 3e-3 * gmax)
            y_l2 = float((gy.double() - gr.double()).norm()) / max(float(gr.double().norm()),
                                                                   3e-3 * gmax * math.sqrt(gr.numel()))
        rows.append((k, e_max, e_l2, ref_max, y_max, y_l2))
    lines.append(f"{'parameter gradient':60s} {'max-norm err':>12s} {'L2 err':>10s} {'max|g_ref|':>11s}"
                 f" {'[torch-autocast max-norm':>25s} {'L2]':>10s}")
    for k, e_max, e_l2, ref_max, y_max, y_l2 in sorted(rows, key=lambda r: -r[1])[:60]:
        lines.append(f"{k:60s} {e_max:12.3e} {e_l2:10.3e} {ref_max:11.3e} {y_max:25.3e} {y_l2:10.3e}")
    import statistics
    lines.append(f"median over {len(rows)} tensors: max-norm {statistics.median(r[1] for r in rows):.3e}, "
                 f"L2 {statistics.median(r[2] for r in rows):.3e}"
                 + (f"; torch-autocast: max-norm {statistics.median(r[4] for r in rows):.3e}, "
                    f"L2 {statistics.median(r[5] for r in rows):.3e}" if sd_yard is not None else ""))
    return [r[:4] for r in rows]


def _assert_bf16_grads(rows):
    import statistics
    for k, e_max, e_l2, _ in rows:
        assert e_l2 <= GRAD_L2_BF16 and e_max <= GRAD_MAX_BF16, (k, e_max, e_l2)
    assert statistics.median(r[2] for r in rows) <= GRAD_L2_MEDIAN_BF16


def _calibrate_running_stats(x, sd, cfg, iters=30):
    """A trained model's BatchNorm running statistics match its activations.  Random running stats (or the 0 / 1
    defaults) let the eval-mode activations drift from layer to layer, which is not what the reference's checkpoints
    look like; 30 train-mode oracle passes (momentum 0.1) bring the running statistics to the batch statistics."""
    with torch.no_grad():
        for _ in range(iters):
            O.dwiseneuro_forward(x, sd, cfg, 0, True)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_true_batch_001_train_parity(dev, mode):
    """B = 4 clips of the C2 workload (every kernel instance of the benchmark runs: W = 64 stride-2 spatial dw,
    K = 1792 / N = 256 GEMM tiles, Gram BatchNorm statistics, split-K wgrad, ...)."""
    from sensorium_b200.losses import MicePoissonLoss
    B = 4
    net = _net(dev, TRUE_BATCH_KW, NUM_NEURONS)
    net.train()
    net.precision = mode
    net._mask_dtype = torch.float32          # draw the masks with the oracle's dtype so the values agree
    x = O.synthetic_clip(B, 16, 64, seed=3).to(dev)
    tg, w = O.synthetic_targets(B, NUM_NEURONS, 16, seed=4)
    tg, w = [t.to(dev) for t in tg], w.to(dev)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    cfg = O.make_cfg(NUM_NEURONS, **TRUE_BATCH_KW)
    taps = []
    ref, ref_loss, sd = _oracle_train(x, sd0, names, cfg, tg, w, 11, taps)

    box, restore = _capture_saved()
    try:
        torch.manual_seed(11)
        out = net(x)
    finally:
        restore()
    loss = MicePoissonLoss()(out, (tg, w))
    loss.backward()
    sv = box[0]

    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    lines = [f"true_batch_001 train parity, B={B}, mode={mode} (tolerance {tol:g})", "stage                      rel err (max-norm)"]
    stage_err = []
    yard = ytaps = sd_y = None
    if mode == "bf16":
        ytaps = []
        yard, _, sd_y = _oracle_autocast(x, sd0, names, cfg, tg, w, 11, ytaps)
        lines[1] += "   [torch-autocast bf16, diagnostic]"
    for i, b in enumerate(sv.blocks):
        got = b.X.view(B, 16, b.Hi, b.Wi, b.ci).permute(0, 4, 1, 2, 3)
        stage_err.append((f"block {i} input", rel(got, taps[i]), rel(ytaps[i].float(), taps[i]) if ytaps else None))
    for m, (a, r_) in enumerate(zip(out, ref)):
        stage_err.append((f"readout {m} output", rel(a, r_), rel(yard[m].float(), r_) if yard else None))
    lerr = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    stage_err.append(("loss", lerr, None))
    for k, e, y in stage_err:
        lines.append(f"{k:26s} {e:.3e}" + (f"            {y:.3e}" if y is not None else ""))
    gaps, mgaps = [], []
    if mode == "bf16":
        gen = torch.Generator().manual_seed(0)
        ygaps = []
        for m in range(len(NUM_NEURONS)):
            r_ = ref[m].detach().cpu()
            noisy = torch.relu(r_ * (1 + 0.3 * torch.randn(r_.shape, generator=gen)))  # single-trial-like target
            n = r_.shape[1]
            f = lambda t: t.permute(1, 0, 2).reshape(n, -1)  # noqa: E731
            c_ref, c_got = O.corr(f(r_), f(noisy)), O.corr(f(out[m].detach().cpu()), f(noisy))
            c_y = O.corr(f(yard[m].detach().float().cpu()), f(noisy))
            gaps.append(float((c_got - c_ref).abs().max()))
            mgaps.append(abs(float(c_got.mean() - c_ref.mean())))          # the metric itself: mean over neurons
            ygaps.append((float((c_y - c_ref).abs().max()), abs(float(c_y.mean() - c_ref.mean()))))
        lines.append("single-trial correlation (metrics.py:11-31), |ours - fp32 oracle| per mouse:")
        lines.append("  max over neurons : " + " ".join(f"{g:.2e}" for g in gaps))
        lines.append("  mean over neurons: " + " ".join(f"{g:.2e}" for g in mgaps))
        lines.append("  [torch-autocast max : " + " ".join(f"{g[0]:.2e}" for g in ygaps) + "]")
        lines.append("  [torch-autocast mean: " + " ".join(f"{g[1]:.2e}" for g in ygaps) + "]")
    run_err = []
    for k, v in net.state_dict().items():
        if v.dtype == torch.int64:
            assert torch.equal(v, sd[k]), k                     # num_batches_tracked: bit-exact
        elif "running_" in k:
            run_err.append((k, rel(v, sd[k])))
    lines.append(f"running statistics: worst {max(run_err, key=lambda r: r[1])}")
    rows = _grad_table(net, sd, lines, sd_y)
    _report(f"parity_fullsize_train_{mode}.txt", lines)

    for k, e, _ in stage_err:
        # the tolerance is defined on the predicted responses (and the loss); in fp32 mode the trunk meets it too, in
        # bf16 mode the per-block rows are the diagnostic that names where the error accumulates
        if mode == "fp32" or not k.startswith("block"):
            assert e < tol, (k, e)
    for k, e in run_err:
        assert e < tol, (k, e)
    if mode == "fp32":
        for k, e_max, e_l2, _ in rows:
            assert e_max <= FP32_TOL, (k, e_max)
    else:
        assert max(mgaps) < CORR_TOL, mgaps
        assert max(gaps) < CORR_NEURON_BF16, gaps
        _assert_bf16_grads(rows)


def test_gram_batchnorm_statistics_full_rows(dev):
    """conv_pw BatchNorm statistics come from the Gram matrix of the block input (dwn_pw_algebra.cu), never from E.  At
    the benchmarked size (batch 32: M = 2.1 M rows in block 0) they must equal the statistics of E = Xb W^T computed
    directly (fp32 matmul, fp64 moments): |d mean| / sigma <= 1e-4 and |d rstd| / rstd <= 1e-4 for every block."""
    B = 32
    net = _net(dev, TRUE_BATCH_KW, NUM_NEURONS)
    net.train()
    net.precision = "bf16"
    x = O.synthetic_clip(B, 16, 64, seed=5).to(dev)
    box, restore = _capture_saved()
    try:
        with torch.no_grad():
            from sensorium_b200 import engine
            engine.run_forward(net, x, 0, "bf16", True, save=True)
    finally:
        restore()
    sv = box[0]
    lines = ["Gram-derived BatchNorm statistics of conv_pw vs direct statistics of E (B=32)",
             "block      M    ci   mid   max|dmean|/sigma   max|drstd|/rstd"]
    worst = 0.0
    for i, b in enumerate(sv.blocks):
        wq = net.core.blocks[2 * i + 1].conv_pw[0].weight.detach().reshape(b.mid, b.ci).to(torch.bfloat16).float()
        M = b.Xb.shape[0]
        s1 = torch.zeros(b.mid, dtype=torch.float64, device=dev)
        s2 = torch.zeros(b.mid, dtype=torch.float64, device=dev)
        for r0 in range(0, M, 1 << 18):
            e = (b.Xb[r0:r0 + (1 << 18)].float() @ wq.t()).double()
            s1 += e.sum(0)
            s2 += (e * e).sum(0)
        mean = s1 / M
        var = s2 / M - mean * mean
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        dm = float(((b.coef1[2].double() - mean).abs() * rstd).max())
        dr = float((b.coef1[3].double() / rstd - 1).abs().max())
        lines.append(f"{i:5d} {M:8d} {b.ci:4d} {b.mid:5d}   {dm:.3e}          {dr:.3e}")
        worst = max(worst, dm, dr)
    _report("parity_gram_stats.txt", lines)
    assert worst <= 1e-4, worst


def test_c4_distillation_bf16_full_size(dev):
    """BASELINE configs[3] at the real architecture: expansion-6 student (drop-path 0.1, dropout 0.4 ON), frozen
    expansion-7 teacher, distill_ratio 0.36, bf16 autocast, through MouseModel.train_step on a HOST batch; the oracle
    runs the same step in fp32 with the same seed.  argus_models.py:31-71."""
    from sensorium_b200.argus_models import MouseModel
    B = 4
    kw_s = dict(TRUE_BATCH_KW, expansion_ratio=6)
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": NUM_NEURONS, **kw_s}), "loss": ("mice_poisson", {}),
              "optimizer": ("AdamW", {"lr": 1e-4, "weight_decay": 0.05}), "device": "cuda:0", "amp": True, "iter_size": 1}
    torch.manual_seed(0)
    m = MouseModel(params)
    student = _net(dev, kw_s, NUM_NEURONS, seed=0)
    m.nn_module.load_state_dict(student.state_dict())
    m.nn_module._mask_dtype = torch.float32
    teacher = _net(dev, TRUE_BATCH_KW, NUM_NEURONS, seed=1).eval()
    x = O.synthetic_clip(B, 16, 64, seed=6)
    tg, w = O.synthetic_targets(B, NUM_NEURONS, 16, seed=7)
    cal = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    _calibrate_running_stats(torch.cat([x, O.synthetic_clip(4, 16, 64, seed=16)]).to(dev), cal,
                             O.make_cfg(NUM_NEURONS, **TRUE_BATCH_KW))
    teacher.load_state_dict(cal)                       # a teacher whose running statistics fit its activations
    m.distill_model, m.distill_ratio = teacher, 0.36
    sd_s = {k: v.detach().clone() for k, v in m.nn_module.state_dict().items()}
    sd_t = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    names = [k for k, _ in m.nn_module.named_parameters()]
    cfg_s, cfg_t = O.make_cfg(NUM_NEURONS, **kw_s), O.make_cfg(NUM_NEURONS, **TRUE_BATCH_KW)
    xd = x.to(dev)
    tg_o, w_o = [t.clone().to(dev) for t in tg], w.clone().to(dev)
    with torch.no_grad():
        t_out = O.dwiseneuro_forward(xd, sd_t, cfg_t, None, False)
    O.distill_fill(tg_o, w_o, t_out, 0.36)
    ref, ref_loss, sd = _oracle_train(xd, sd_s, names, cfg_s, tg_o, w_o, 21)
    torch.manual_seed(21)
    out = m.train_step((x, ([t.clone() for t in tg], w.clone())), None)
    lines = ["C4 distillation step, bf16, B=4 (student er=6 with drop rates on, teacher er=7)"]
    lerr = abs(out["loss"] - float(ref_loss)) / abs(float(ref_loss))
    lines.append(f"loss rel err {lerr:.3e}")
    got_t, got_w = out["target"]
    terr = max(rel(a, b) for a, b in zip(got_t, tg_o))
    perr = max(rel(a, b) for a, b in zip(out["prediction"], ref))
    lines.append(f"filled targets rel err {terr:.3e}; predictions rel err {perr:.3e}; weights rel err {rel(got_w, w_o):.3e}")
    rows = _grad_table(m.nn_module, sd, lines)
    _report("parity_fullsize_c4_bf16.txt", lines)
    assert rel(got_w, w_o) < 1e-6
    assert lerr < BF16_TOL and terr < BF16_TOL and perr < BF16_TOL
    _assert_bf16_grads(rows)


def _tiny_model(dev, iter_size, amp=False, ema=False):
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.ema import ModelEma
    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw}), "loss": ("mice_poisson", {}),
              "optimizer": ("AdamW", {"lr": 1e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": amp,
              "iter_size": iter_size}
    torch.manual_seed(0)
    m = MouseModel(params)
    m.nn_module.load_state_dict(_net(dev, kw, TINY_OUTS, seed=2).state_dict())
    if ema:
        m.model_ema = ModelEma(m.nn_module, decay=0.9)
    return m, kw


def test_iter_size_2_gradient_accumulation(dev):
    """argus_models.py:46-56: the batch is chunked, every chunk runs forward / loss / iter_size / backward and the
    gradients accumulate before ONE optimizer step.  Mouse 1 has samples only in the first chunk, mouse 2 in none:
    mouse 1 keeps its accumulated gradient (and is stepped), mouse 2 stays grad=None (and untouched)."""
    m, kw = _tiny_model(dev, 2)
    B = 6
    x = O.synthetic_clip(B, 16, 32, seed=8)
    tg, w = O.synthetic_targets(B, TINY_OUTS, 16, seed=9)
    w.zero_()
    w[:, 0] = 1.0
    w[0, 1] = 1.0                                                  # mouse 1: only sample 0 (chunk 0)
    sd0 = {k: v.detach().clone() for k, v in m.nn_module.state_dict().items()}
    names = [k for k, _ in m.nn_module.named_parameters()]
    cfg = O.make_cfg(TINY_OUTS, **kw)
    sd = {k: v.detach().clone() for k, v in sd0.items()}
    params = [sd[k].requires_grad_(True) for k in names]
    ref_loss = 0.0
    for c in range(2):
        sl = slice(3 * c, 3 * c + 3)
        xo = x[sl].to(dev)
        out = O.dwiseneuro_forward(xo, sd, cfg, None, True)
        l = O.mice_poisson_loss(out, [t[sl].to(dev) for t in tg], w[sl].to(dev)) / 2
        l.backward()
        ref_loss += float(l)
    res = m.train_step((x, (tg, w)), None)
    assert abs(res["loss"] - ref_loss) / abs(ref_loss) < FP32_TOL
    gmax = max(float(p.grad.abs().max()) for p in params if p.grad is not None)
    for k, p in m.nn_module.named_parameters():
        if sd[k].grad is None:
            assert p.grad is None and k.startswith("readouts.2."), k
            assert torch.equal(p.detach(), sd0[k]), k             # never stepped, never decayed
            continue
        err = float((p.grad - sd[k].grad).abs().max())
        assert err <= FP32_TOL * max(float(sd[k].grad.abs().max()), 3e-3 * gmax) + 1e-6 * gmax, (k, err)
    # ONE AdamW step was taken from the accumulated gradients: torch.optim.AdamW fed with the same (accumulated)
    # gradients must land on the same weights (the first Adam step is lr*sign(g), so the oracle's own round-off-level
    # gradients cannot be used for tensors whose gradient is analytically zero)
    for k, p in m.nn_module.named_parameters():
        if p.grad is not None:
            sd[k].grad = p.grad.detach().clone()
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05)
    opt.step()
    for k, p in m.nn_module.named_parameters():
        assert rel(p.detach(), sd[k].detach()) < 1e-5, k
    for k, v in m.nn_module.state_dict().items():                 # running stats saw both chunks
        if v.dtype == torch.int64:
            assert torch.equal(v, sd[k]) and int(v) == 2, k
        elif "running_" in k:
            assert rel(v, sd[k]) < FP32_TOL, k


@pytest.mark.parametrize("ema", [False, True])
def test_val_step_numerics(dev, ema):
    """argus_models.py:73-87: eval mode, no autocast (fp32), EMA weights when present; loss and predictions vs the
    oracle's eval forward + MicePoissonLoss."""
    m, kw = _tiny_model(dev, 1, amp=True, ema=ema)
    with torch.no_grad():
        for k, v in m.nn_module.state_dict().items():
            if "running_mean" in k:
                v.uniform_(-0.2, 0.2)
            if "running_var" in k:
                v.uniform_(0.5, 1.5)
    if ema:
        m.model_ema.set(m.nn_module)
        with torch.no_grad():  # EMA weights differ from the raw ones: val_step must use the EMA module
            for p in m.model_ema.ema.parameters():
                p.mul_(1.01)
    x = O.synthetic_clip(3, 16, 32, seed=12)
    tg, w = O.synthetic_targets(3, TINY_OUTS, 16, seed=13)
    src = m.model_ema.ema if ema else m.nn_module
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    with torch.no_grad():
        ref = O.dwiseneuro_forward(x.to(dev), sd, O.make_cfg(TINY_OUTS, **kw), None, False)
        ref_loss = O.mice_poisson_loss(ref, [t.to(dev) for t in tg], w.to(dev))
    res = m.val_step((x, (tg, w)), None)
    assert abs(res["loss"] - float(ref_loss)) / abs(float(ref_loss)) < FP32_TOL
    for a, b in zip(res["prediction"], ref):
        assert a.dtype == torch.float32 and rel(a, b) < FP32_TOL
    for k, v in src.state_dict().items():                          # eval mode: no state is touched
        assert torch.equal(v, sd[k]), k


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_exchange_two_gpus():
    """tests/gpu_checks/check_dp.py under torchrun (NCCL, 2 ranks, rank-dependent mice): exchanged gradients == mean of
    the per-shard gradients, has-grad flags skip only the mouse absent on every rank."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(ROOT / "tests/gpu_checks/check_dp.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "check_dp: PASS" in r.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_entry_points_behind_the_c_abi_two_gpus():
    """tests/gpu_checks/check_comm_cabi.py under torchrun: dwn_comm_init / dwn_allreduce_bucket (include/dwn_b200.h,
    SURVEY.md 8b) - fp32 mean, bf16 sum, int32 max, grouped buckets - against closed-form expectations."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29641",
                        str(ROOT / "tests/gpu_checks/check_comm_cabi.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "check_comm_cabi: PASS" in r.stdout
