"""GPU parity tests (-m gpu): the CUDA path (through the C ABI) against the oracle and the committed golden fixtures.

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-4 relative; bf16 mode <= 2e-2 relative on the predicted responses
with the single-trial correlation metric (per-neuron corr averaged over a mouse's neurons, metrics.py:66-70) within 1e-3;
integer / index work bit-exact.  All bounds are absolute numbers (none is relaxed against torch's bf16 autocast); the
ones bf16 cannot meet — the worst single neuron's correlation, gradients — carry their own stated bound, see
tests/test_parity_fullsize_gpu.py for the measurements behind them."""
import copy
import math

import numpy as np
import pytest
import torch

from oracle import dwiseneuro_oracle as O
from tests.shapes import TINY_KW, TINY_OUTS, TRUE_BATCH_KW

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-4, 2e-2
CORR_TOL, CORR_NEURON_BF16 = 1e-3, 3e-2          # metric (mean over neurons) / worst single neuron
GRAD_L2_BF16, GRAD_MAX_BF16 = 0.2, 0.3           # per-tensor relative L2 / max-norm error of bf16 gradients


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _tiny(dev, seed=0):
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.utils import init_weights
    torch.manual_seed(seed)
    net = DwiseNeuro(readout_outputs=TINY_OUTS, **TINY_KW)
    init_weights(net)
    return net.to(dev)


def _corr_gap(pred, ref, target):
    """|corr(pred, target) - corr(ref, target)| (metrics.py:11-31 semantics): (worst neuron, mean-over-neurons metric)."""
    a = O.corr(pred.permute(1, 0, 2).reshape(pred.shape[1], -1), target.permute(1, 0, 2).reshape(pred.shape[1], -1))
    b = O.corr(ref.permute(1, 0, 2).reshape(pred.shape[1], -1), target.permute(1, 0, 2).reshape(pred.shape[1], -1))
    return float((a - b).abs().max()), abs(float(a.mean() - b.mean()))


def test_tiny_golden_forward_loss_grads(dev, golden_dir):
    from sensorium_b200.losses import MicePoissonLoss
    g = torch.load(golden_dir / "tiny_forward_backward.pt", weights_only=False)
    x = O.synthetic_clip(4, 16, 32, seed=0).to(dev)
    tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=1)
    tgd, wd = [t.to(dev) for t in tg], w.to(dev)
    for mode, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        net = _tiny(dev)
        net.precision = mode
        net.eval()
        with torch.no_grad():
            ev = net(x)
        for a, b in zip(ev, g["eval_out"]):
            assert rel(a.cpu(), b) < tol
        net.train()
        net._rng_device, net._mask_dtype = "cpu", torch.float32  # the golden masks came from the CPU generator
        torch.manual_seed(5)
        tr = net(x)
        for a, b in zip(tr, g["train_out"]):
            assert rel(a.detach().cpu(), b) < tol
        loss = MicePoissonLoss()(tr, (tgd, wd))
        assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < tol
        loss.backward()
        gmax = max(float(v.abs().max()) for v in g["grads"].values())
        for k, p in net.named_parameters():
            if k in g["none_grads"]:
                assert p.grad is None, k
                continue
            err = float((p.grad.cpu() - g["grads"][k]).abs().max())
            # biases in front of a batch-stat BatchNorm have an analytically zero gradient (pure round-off in the
            # reference too): absolute floor relative to the largest gradient of the network
            floor = (3e-2 if mode == "bf16" else 3e-3) * gmax
            assert err <= (3 if mode == "bf16" else 1) * tol * max(float(g["grads"][k].abs().max()), floor) + 1e-6 * gmax, k
        for k, v in g["running"].items():
            got = net.state_dict()[k].cpu()
            if v.dtype == torch.int64:
                assert torch.equal(got, v), k           # num_batches_tracked: bit-exact
            else:
                assert rel(got, v) < tol, k
        if mode == "bf16":
            # single-trial correlation (metrics.py:11-31) of the train-mode outputs against a noisy single-trial-like
            # target: worst single neuron.  (The 1e-3 bound on the METRIC — the mean over a mouse's >= 7440 neurons — is
            # asserted on the real architecture in test_c1_full_architecture_golden and test_parity_fullsize_gpu.py; a
            # mean over 37 neurons of this tiny model does not average the per-neuron noise down.)
            gen = torch.Generator().manual_seed(0)
            for m in range(len(TINY_OUTS)):
                ref_m = g["train_out"][m]
                noisy = torch.relu(ref_m * (1 + torch.randn(ref_m.shape, generator=gen)))
                worst, metric = _corr_gap(tr[m].detach().cpu(), ref_m, noisy)
                assert worst < CORR_NEURON_BF16, (m, worst, metric)


def test_c1_full_architecture_golden(dev, golden_dir):
    """BASELINE configs[0]: true_batch_001 architecture, batch 1, 16x64x64 clip, single-mouse readout."""
    from sensorium_b200 import DwiseNeuro, constants
    from sensorium_b200.utils import init_weights
    g = torch.load(golden_dir / "c1_forward_index0.pt", weights_only=False)
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW)
    init_weights(net)
    net = net.to(dev).eval()
    g_sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = O.synthetic_clip(1, 16, 64, seed=0).to(dev)
    with torch.no_grad():
        y32 = net(x, 0)
        assert y32.shape == (1, 7863, 16) and y32.dtype == torch.float32
        assert rel(y32.cpu(), g["out_index0"]) < FP32_TOL
        # bf16: on a model whose running statistics fit its activations (what a trained checkpoint looks like; with the
        # 0 / 1 defaults eval-mode activations drift from layer to layer and NO bf16 implementation holds 2e-2 —
        # torch's own autocast of the reference is at 7e-2 there).  30 train-mode oracle passes calibrate them.
        cfg = O.make_cfg(constants.num_neurons, **TRUE_BATCH_KW)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        xc = O.synthetic_clip(8, 16, 64, seed=21).to(dev)
        for _ in range(30):
            O.dwiseneuro_forward(xc, sd, cfg, 0, True)
        net.load_state_dict(sd)
        ref = O.dwiseneuro_forward(x3 := torch.cat([x, O.synthetic_clip(3, 16, 64, seed=7).to(dev)]), sd, cfg, 0, False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = net(x3, 0)
        assert y16.dtype == torch.float32 and rel(y16, ref) < BF16_TOL, rel(y16, ref)
        noisy = ref * (1 + 0.3 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(0)).to(dev))
        worst, metric = _corr_gap(y16.cpu(), ref.cpu(), noisy.cpu())
        assert metric < CORR_TOL and worst < CORR_NEURON_BF16, (worst, metric)
        net.load_state_dict({k: v.to(dev) for k, v in g_sd.items()})
        # index / list consistency and batch invariance of the eval graph (windows can be batched, predictors.py)
        full = net(x)
        assert len(full) == 10 and torch.equal(full[0], y32)
        x3 = torch.cat([x, O.synthetic_clip(2, 16, 64, seed=7).to(dev)])
        y3 = net(x3, 0)
        assert rel(y3[:1].cpu(), g["out_index0"]) < FP32_TOL


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T,HW,seed", [(3, 8, 16, 1), (2, 16, 32, 2)])
def test_train_step_vs_oracle_shared_rng(dev, mode, B, T, HW, seed):
    """Ragged shapes (odd batch, T != 16, odd neuron counts); drop-path and Dropout1d masks come from the same torch
    RNG calls as the reference, so a shared seed gives the same masks."""
    net = _tiny(dev, seed)
    for n_, p in net.named_parameters():  # non-trivial BN affine / biases
        if p.dim() == 1:
            torch.nn.init.uniform_(p, 0.5, 1.5) if n_.endswith("bn.weight") else torch.nn.init.uniform_(p, -0.3, 0.3)
    net.train()
    net.precision = mode
    net._mask_dtype = torch.float32  # same mask values as the fp32 oracle
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
    ref_loss = O.mice_poisson_loss(ref, tg, w)
    ref_loss.backward()
    torch.manual_seed(11)
    out = net(x)
    loss = O.mice_poisson_loss(out, tg, w)
    loss.backward()
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    for a, b in zip(out, ref):
        assert rel(a, b) < tol
    gmax = max(float(sd[k].grad.abs().max()) for k in names if sd[k].grad is not None)
    for k, p in net.named_parameters():
        if sd[k].grad is None:
            assert p.grad is None
            continue
        gr = sd[k].grad
        err = float((p.grad - gr).abs().max())
        if mode == "fp32":
            assert err <= tol * max(float(gr.abs().max()), 3e-3 * gmax) + 1e-6 * gmax, (k, err)
        else:
            l2 = float((p.grad.double() - gr.double()).norm()) / max(float(gr.double().norm()),
                                                                      3e-3 * gmax * math.sqrt(gr.numel()))
            assert err <= GRAD_MAX_BF16 * max(float(gr.abs().max()), 3e-3 * gmax) and l2 <= GRAD_L2_BF16, (k, err, l2)


def test_loss_kernels_and_absent_mice(dev):
    from sensorium_b200.losses import MicePoissonLoss
    torch.manual_seed(0)
    outs = (33, 20, 7)
    preds = [torch.rand(5, n, 16, device=dev) * 3 + 0.01 for n in outs]
    for p in preds:
        p.requires_grad_(True)
    tg, w = O.synthetic_targets(5, outs, 16, seed=3)
    w[:, 1] = 0  # mouse 1 absent
    w[:, 0] = torch.tensor([0.5, 0, 1, 0, 0.25])
    w[:, 2] = torch.tensor([0, 1, 0, 0, 0.75])
    tg, w = [t.to(dev) for t in tg], w.to(dev)
    loss = MicePoissonLoss()(preds, (tg, w))
    loss.backward()
    ref_p = [p.detach().clone().requires_grad_(True) for p in preds]
    ref = O.mice_poisson_loss(ref_p, tg, w)
    ref.backward()
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-6
    assert preds[1].grad is None and ref_p[1].grad is None
    for a, b in ((preds[0], ref_p[0]), (preds[2], ref_p[2])):
        assert rel(a.grad, b.grad) < 1e-6
        assert float(a.grad[3].abs().max()) == 0.0  # sample 3 is masked for both mice: exactly zero


def test_fused_adamw_and_ema_golden(dev, golden_dir):
    from sensorium_b200.ema import ModelEma
    from sensorium_b200.optim import FusedAdamW
    a = torch.load(golden_dir / "adamw_steps.pt", weights_only=False)
    p = torch.nn.Parameter(a["p0"].clone().to(dev))
    q = torch.nn.Parameter(torch.ones(7, device=dev))  # never receives a gradient: must stay untouched
    opt = FusedAdamW([p, q], lr=a["lr"], weight_decay=a["wd"])
    for i, g in enumerate(a["grads"]):
        p.grad = g.clone().to(dev)
        opt.step()
        assert rel(p.detach().cpu(), a["traj"][i]) < 1e-6
    assert torch.equal(q.detach().cpu(), torch.ones(7)) and int(opt.state[q]["step"]) == 0 and int(opt.state[p]["step"]) == 3
    # EMA over every state entry, int64 buffers through fp32 with truncation (ema.py:49-55)
    g = torch.load(golden_dir / "ema_update.pt", weights_only=False)
    from sensorium_b200 import DwiseNeuro
    net = DwiseNeuro(readout_outputs=(5, 4), core_features=(8,), spatial_strides=(1,), expansion_ratio=2,
                     se_reduce_ratio=4, cortex_features=(8,), groups=2)
    net.load_state_dict(g["before"])
    net = net.to(dev)
    ema = ModelEma(net, decay=g["decay"])
    net.load_state_dict(g["model"])
    ema.update(net)
    for k, v in ema.ema.state_dict().items():
        if v.dtype == torch.int64:
            assert torch.equal(v.cpu(), g["after"][k]), k
        else:
            assert rel(v.cpu(), g["after"][k]) < 1e-6 or float(g["after"][k].abs().max()) == 0.0, k


def test_distill_fill_golden(dev, golden_dir):
    from sensorium_b200.argus_models import MouseModel
    d = torch.load(golden_dir / "distill_fill.pt", weights_only=False)
    mm = MouseModel.__new__(MouseModel)
    mm.distill_ratio = d["ratio"]
    teacher = [t.to(dev) for t in d["teacher"]]
    mm.distill_model = lambda inp: teacher
    tg = [t.clone().to(dev) for t in d["targets_in"]]
    w = d["weights_in"].clone().to(dev)
    mm.add_distill_predictions(None, (tg, w))
    assert rel(w.cpu(), d["weights_out"]) < 1e-6
    for a, b in zip(tg, d["targets_out"]):
        assert torch.equal(a.cpu(), b)


def test_window_kernels_golden(dev, golden_dir):
    """Integer index work of the predictor is bit-exact: gather == fancy indexing, blend == reference loop."""
    from sensorium_b200._lib import call
    g = torch.load(golden_dir / "predictor_blend.pt", weights_only=False)
    st = torch.cuda.current_stream(dev).cuda_stream
    inputs = O.stack_inputs(g["video"], g["behavior"], g["pupil"]).to(dev)
    Cn, L, H, W = inputs.shape
    size, step, behind = 16, 2, 30
    nwin = L - behind
    clips = torch.empty(nwin, Cn, size, H, W, device=dev)
    call("dwn_window_gather", inputs, clips, Cn, L, H * W, size, step, behind, nwin, st)
    for wi in (0, 5, nwin - 1):
        idx = O.make_window_indexes(behind + wi, size, step)
        assert torch.equal(clips[wi], inputs[:, idx])
    n_out = g["n_out"]
    feat = clips[:, :, :, 20:24, 30:34].mean((1, 3, 4))                       # (nwin, 16)
    preds = (feat[:, None, :] * torch.arange(1, n_out + 1, device=dev)[None, :, None]).float().contiguous()
    for blend in ("ones", "linear"):
        bw = torch.ones(size, device=dev) if blend == "ones" else torch.linspace(0, 1, size, dtype=torch.float64).float().to(dev)
        out = torch.empty(n_out, L, device=dev)
        call("dwn_window_blend", preds, bw, out, n_out, L, size, step, 0, nwin, n_out * size, st)
        assert rel(out.cpu(), g["responses"][blend]) < 1e-5
        assert float(out[:, 1].abs().max()) == float(g["responses"][blend][:, 1].abs().max())


def test_predictor_end_to_end(dev, tmp_path):
    """Batched device predictor == the reference's per-window loop (oracle.predict_trial) on a real (tiny) model."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.predictors import Predictor
    from sensorium_b200.utils import init_weights
    kw = {"readout_outputs": (9, 6), "core_features": (8, 16), "spatial_strides": (2, 2), "expansion_ratio": 2,
          "se_reduce_ratio": 4, "cortex_features": (16,), "groups": 2}
    params = {"nn_module": ("dwiseneuro", kw), "loss": None, "optimizer": None, "device": "cuda:0",
              "frame_stack": {"size": 16, "step": 2, "position": "last"},
              "inputs_processor": ("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0}),
              "responses_processor": ("identity", {}), "amp": True, "iter_size": 1}
    torch.manual_seed(3)
    m = MouseModel(params)
    init_weights(m.nn_module)
    with torch.no_grad():  # non-trivial running stats
        for k, v in m.nn_module.state_dict().items():
            if "running_mean" in k:
                v.uniform_(-0.2, 0.2)
            if "running_var" in k:
                v.uniform_(0.5, 1.5)
    path = tmp_path / "model-001-0.100000.pth"
    m.save(path)
    pr = Predictor(path, device="cuda:0", blend_weights="ones", window_batch=7)
    rs = np.random.RandomState(0)
    L = 41
    video = rs.randint(0, 256, (36, 64, L)).astype(np.uint8)
    beh, pup = rs.rand(2, L).astype(np.float32), rs.rand(2, L).astype(np.float32)
    got = pr.predict_trial(video, beh, pup, 1)
    assert got.shape == (6, L) and got.dtype == np.float32
    sd = {k: v.detach().cpu() for k, v in m.nn_module.state_dict().items()}
    cfg = O.make_cfg(kw["readout_outputs"], **{k: v for k, v in kw.items() if k != "readout_outputs"})
    inputs = O.stack_inputs(torch.from_numpy(video), torch.from_numpy(beh), torch.from_numpy(pup))
    with torch.no_grad():
        want = O.predict_trial(lambda c: O.dwiseneuro_forward(c, sd, cfg, 1, False), inputs, 6, 16, 2, "ones")
    assert rel(torch.from_numpy(got), want) < FP32_TOL
    assert float(np.abs(got[:, :1]).max()) >= 0.0 and np.all(got[:, [1, 3]] == 0.0) == bool((want[:, [1, 3]] == 0).all())


def test_mouse_model_train_step_and_determinism(dev):
    """The public wrapper (argus_models.py:43-71 contract): host batch in, dict out, weights move, EMA follows,
    absent mice keep their readouts untouched; two identical runs are bit-identical (deterministic reductions)."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.ema import ModelEma
    from sensorium_b200.utils import init_weights

    def run():
        params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **TINY_KW}),
                  "loss": ("mice_poisson", {}), "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}),
                  "device": "cuda:0", "amp": True, "iter_size": 1}
        torch.manual_seed(0)
        m = MouseModel(params)
        init_weights(m.nn_module)
        m.model_ema = ModelEma(m.nn_module, decay=0.9)
        start = {k: v.detach().clone() for k, v in m.nn_module.state_dict().items()}
        x = O.synthetic_clip(4, 16, 32, seed=0)
        tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=1)
        w[:, 2] = 0
        w[:, 0] = 1
        losses = []
        torch.manual_seed(9)
        for _ in range(3):
            out = m.train_step((x, (tg, w)), None)
            losses.append(out["loss"])
        return m, out, losses, start

    m1, out, l1, start = run()
    m2, _, l2, _ = run()
    assert set(out) == {"prediction", "target", "loss"} and len(out["prediction"]) == 3
    assert l1 == l2 and all(math.isfinite(v) for v in l1) and l1[2] < l1[0]
    for a, b in zip(m1.nn_module.state_dict().values(), m2.nn_module.state_dict().values()):
        assert torch.equal(a, b)
    assert torch.equal(m1.nn_module.readouts[2].layer[1].weight, start["readouts.2.layer.1.weight"])  # absent mouse
    assert not torch.equal(m1.nn_module.readouts[0].layer[1].weight, start["readouts.0.layer.1.weight"])
    assert int(m1.nn_module.core.stem[1].bn.num_batches_tracked) == 3
    v = m1.val_step((O.synthetic_clip(2, 16, 32, seed=5), O.synthetic_targets(2, TINY_OUTS, 16, seed=6)), None)
    assert math.isfinite(v["loss"])


def test_full_size_properties(dev):
    """BASELINE configs[1] shape (true_batch_001, batch 32 is benchmarked; batch 8 here to bound test time):
    size-independent properties instead of an oracle run."""
    from sensorium_b200 import DwiseNeuro, constants
    from sensorium_b200.losses import MicePoissonLoss
    from sensorium_b200.utils import init_weights
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW)
    init_weights(net)
    net = net.to(dev).train()
    B = 8
    x = O.synthetic_clip(B, 16, 64, seed=3).to(dev)
    tg, w = O.synthetic_targets(B, constants.num_neurons, 16, seed=4)
    tg, w = [t.to(dev) for t in tg], w.to(dev)
    outs = []
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        torch.manual_seed(1)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = net(x)
            loss = MicePoissonLoss()(pred, (tg, w))
        loss.backward()
        outs.append((float(loss), [p.detach().clone() for p in pred],
                     {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
        with torch.no_grad():  # restore running stats so both runs are identical
            pass
    assert all(p.shape == (B, n, 16) and bool(torch.isfinite(p).all()) and float(p.min()) >= 0 for p, n in
               zip(outs[0][1], constants.num_neurons))
    # determinism of the forward pass for identical inputs / masks
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)
    live = (w != 0).any(0).tolist()
    for m, lv in enumerate(live):
        has = f"readouts.{m}.layer.1.weight" in outs[0][2]
        assert has == lv
    # gradients of biases in front of batch-stat BN vanish (analytic property), everything finite
    gmax = max(float(g.abs().max()) for g in outs[0][2].values())
    assert all(bool(torch.isfinite(g).all()) for g in outs[0][2].values())
    assert float(outs[0][2]["core.blocks.5.conv_pwl.1.bn.bias"].abs().max()) < 1e-2 * gmax


def test_correlation_metric_device(dev, golden_dir):
    """streaming device accumulators == the reference's concatenate-and-corr (golden) == oracle."""
    from sensorium_b200.metrics import CorrelationMetric, corr
    from tests.shapes import corr_step_outputs
    g = torch.load(golden_dir / "corr_metric.pt", weights_only=False)
    steps = corr_step_outputs()
    metric = CorrelationMetric()
    for s in steps:
        tg, w = s["target"]
        metric.update({"prediction": [p.to(dev) for p in s["prediction"]], "target": ([t.to(dev) for t in tg], w.to(dev))})
    res = metric.compute()
    assert set(res) == {0, 2}
    ref = O.correlation_metric(steps)
    per = metric.compute_per_neuron()
    for m in res:
        assert abs(float(res[m]) - g["mice_corr"][m]) < 2e-6
        assert abs(float(res[m]) - ref["mice_corr"][m]) < 2e-6
        assert float((per[m][0].cpu().double() - g["per_neuron"][m]).abs().max()) < 5e-6
    # update() after reset() starts from zero; 2-D (B, n) predictions take the T = 1 path of the reference
    metric.reset()
    p2 = torch.rand(6, 9, device=dev)
    t2 = p2 * 2 + torch.rand(6, 9, device=dev)
    w2 = torch.ones(6, 1, device=dev)
    metric.update({"prediction": [p2], "target": ([t2], w2)})
    want = corr(p2.cpu().numpy(), t2.cpu().numpy(), axis=0).mean()
    assert abs(float(metric.compute()[0]) - float(want)) < 2e-6
    state = type("S", (), {"phase": "val", "metrics": {}})()
    metric.epoch_complete(state)
    assert set(state.metrics) == {"val_corr_mouse_0", "val_corr"}


@pytest.mark.parametrize("vdtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("Hv,Wv,fill", [(36, 64, 0.0), (21, 30, 7.5)])
def test_assemble_clips_bit_exact(dev, vdtype, Hv, Wv, fill):
    """clips built on the device from the raw trial == StackInputsProcessor (inputs.py:22-36) followed
    by the window gather of predict_trial (predictors.py:42-51), bit for bit."""
    from sensorium_b200._lib import call
    L, size, step = 53, 16, 2
    behind = (size - 1) * step
    g = torch.Generator().manual_seed(5)
    video = torch.randint(0, 256, (Hv, Wv, L), generator=g).to(vdtype)
    beh, pup = torch.rand(2, L, generator=g) * 9, torch.rand(2, L, generator=g) * 30
    stacked = O.stack_inputs(video, beh, pup, size=(64, 64), fill=fill)            # (5, L, 64, 64)
    last0, nw = behind + 3, 11
    want = torch.stack([stacked[:, O.make_window_indexes(last0 + i, size, step, "last")] for i in range(nw)])
    clips = torch.empty((nw, 5, size, 64, 64), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    call("dwn_assemble_clips", video.to(dev), 2 if vdtype == torch.uint8 else 0, beh.to(dev), pup.to(dev), clips, L, Hv, Wv,
         64, 64, fill, size, step, last0, nw, st)
    assert torch.equal(clips.cpu(), want)
    with pytest.raises(Exception):  # window range outside the trial is refused, not clamped
        call("dwn_assemble_clips", video.to(dev), 0, beh.to(dev), pup.to(dev), clips, L, Hv, Wv, 64, 64, fill, size, step,
             L - 2, nw, st)


def test_distillation_train_step_vs_oracle(dev):
    """C4 (argus_models.py:31-71): frozen teacher (wider expansion) fills the targets / weights of the mice a sample does
    not belong to, then the student trains on all mice.  Loss, the mutated targets and the student gradients must match
    the oracle (teacher forward in eval mode -> distill_fill -> student forward -> MicePoissonLoss)."""
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.utils import init_weights
    kw_s = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0, expansion_ratio=2)
    kw_t = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0, expansion_ratio=3)
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw_s}), "loss": ("mice_poisson", {}),
              "optimizer": ("AdamW", {"lr": 1e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": False, "iter_size": 1}
    torch.manual_seed(0)
    m = MouseModel(params)
    init_weights(m.nn_module)
    m.nn_module.precision = "fp32"
    torch.manual_seed(1)
    teacher = DwiseNeuro(readout_outputs=TINY_OUTS, **kw_t).to(dev)
    init_weights(teacher)
    teacher.precision = "fp32"
    teacher.eval()
    m.distill_model, m.distill_ratio = teacher, 0.36
    B = 5
    x = O.synthetic_clip(B, 16, 32, seed=3)
    tg, w = O.synthetic_targets(B, TINY_OUTS, 16, seed=4)
    sd_s = {k: v.detach().clone() for k, v in m.nn_module.state_dict().items()}
    sd_t = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    names = [k for k, _ in m.nn_module.named_parameters()]
    for k in names:
        sd_s[k].requires_grad_(True)
    # ---- oracle
    cfg_s, cfg_t = O.make_cfg(TINY_OUTS, **kw_s), O.make_cfg(TINY_OUTS, **kw_t)
    xd = x.to(dev)
    tg_o, w_o = [t.clone().to(dev) for t in tg], w.clone().to(dev)
    with torch.no_grad():
        t_out = O.dwiseneuro_forward(xd, sd_t, cfg_t, None, False)
    O.distill_fill(tg_o, w_o, t_out, 0.36)
    ref_loss = O.mice_poisson_loss(O.dwiseneuro_forward(xd, sd_s, cfg_s, None, True), tg_o, w_o)
    ref_loss.backward()
    # ---- product: the public train step on a HOST batch
    out = m.train_step((x, ([t.clone() for t in tg], w.clone())), None)
    assert abs(out["loss"] - float(ref_loss)) / abs(float(ref_loss)) < FP32_TOL
    got_t, got_w = out["target"]
    assert rel(got_w, w_o) < 1e-6
    for a, b in zip(got_t, tg_o):
        assert rel(a, b) < FP32_TOL
    gmax = max(float(sd_s[k].grad.abs().max()) for k in names)
    for k, p in m.nn_module.named_parameters():
        assert p.grad is not None, k  # with distillation every mouse is live for every sample
        err = float((p.grad - sd_s[k].grad).abs().max())
        assert err <= FP32_TOL * max(float(sd_s[k].grad.abs().max()), 3e-3 * gmax) + 1e-6 * gmax, (k, err)


def test_device_prefetcher_is_transparent(dev):
    """Batches that arrive through DevicePrefetcher (copied one ahead on a side stream) train bit-identically to host
    batches handed to train_step directly, including a batch in which a mouse has no sample."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.prefetch import DevicePrefetcher
    from sensorium_b200.utils import init_weights

    def batches():
        out = []
        for i in range(3):
            x = O.synthetic_clip(4, 16, 32, seed=10 + i)
            tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=20 + i)
            if i == 1:
                w[:, 1] = 0
                w[:, 0] = 1
            out.append((x.pin_memory(), ([t.pin_memory() for t in tg], w.pin_memory())))
        return out

    def run(prefetch):
        params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **TINY_KW}), "loss": ("mice_poisson", {}),
                  "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
                  "iter_size": 1}
        torch.manual_seed(0)
        m = MouseModel(params)
        init_weights(m.nn_module)
        torch.manual_seed(3)
        src = DevicePrefetcher(batches(), dev) if prefetch else batches()
        losses = [m.train_step(b, None)["loss"] for b in src]
        return m, losses

    m1, l1 = run(False)
    m2, l2 = run(True)
    assert l1 == l2
    for a, b in zip(m1.nn_module.state_dict().values(), m2.nn_module.state_dict().values()):
        assert torch.equal(a, b)


def test_cutmix_and_collation_on_device(dev, golden_dir):
    """box copy + target lerp on a device batch == the reference CutMix per sample (golden), and the
    compact targets scattered into per-mouse tensors == construct_mice_sample + default collate (oracle)."""
    from sensorium_b200.mixers import DeviceCutMix, collate_on_device
    g = torch.load(golden_dir / "cutmix.pt", weights_only=False)
    recs = g["records"]
    np.random.seed(g["seed"])
    mixer = DeviceCutMix(g["alpha"], g["prob"])
    x1 = torch.stack([r["s1"][0] for r in recs]).to(dev)
    x2 = torch.stack([r["s2"][0] for r in recs]).to(dev)
    t1 = torch.stack([r["s1"][1] for r in recs]).to(dev)
    t2 = torch.stack([r["s2"][1] for r in recs]).to(dev)
    x, t = mixer(x1, x2, t1, t2)
    assert sum(r["used"] for r in recs) >= 2
    for b, r in enumerate(recs):
        assert torch.equal(x[b].cpu(), r["out"][0]), b        # pure copies: bit-exact
        assert rel(t[b].cpu(), r["out"][1]) < 1e-6, b
    # collation: ragged neuron counts, a mouse without a sample
    outs = (5, 9, 4)
    ids = torch.tensor([2, 0, 0, 2, 2])
    gen = torch.Generator().manual_seed(8)
    samples = [(int(m), torch.rand(outs[int(m)], 6, generator=gen)) for m in ids]
    compact = torch.zeros(len(ids), max(outs), 6)
    for b, (m, tt) in enumerate(samples):
        compact[b, :tt.shape[0]] = tt
        compact[b, tt.shape[0]:] = 123.0  # padding rows must never leak into the batch
    want_t, want_w = O.collate_mice_batch(samples, outs)
    got_t, got_w = collate_on_device(compact.to(dev), ids.to(dev), outs)
    assert torch.equal(got_w.cpu(), want_w)
    for a, b_ in zip(got_t, want_t):
        assert torch.equal(a.cpu(), b_)


def test_ensemble_predictor_mean_of_folds(dev, tmp_path):
    """scripts/predict.py:43-49: the response of a trial is np.mean over the fold models of Predictor.predict_trial.
    EnsemblePredictor (all models resident, windows assembled once, sum over models on the device) must equal the
    oracle's per-window loop run per model and averaged — fp32 <= 1e-4; the bf16 option stays within 2e-2."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.predictors import EnsemblePredictor
    from sensorium_b200.utils import init_weights
    kw = {"readout_outputs": (9, 6), "core_features": (8, 16), "spatial_strides": (2, 2), "expansion_ratio": 2,
          "se_reduce_ratio": 4, "cortex_features": (16,), "groups": 2}
    paths, sds = [], []
    for fold in range(3):
        params = {"nn_module": ("dwiseneuro", kw), "loss": None, "optimizer": None, "device": "cuda:0",
                  "frame_stack": {"size": 16, "step": 2, "position": "last"},
                  "inputs_processor": ("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0}),
                  "responses_processor": ("identity", {}), "amp": True, "iter_size": 1}
        torch.manual_seed(10 + fold)
        m = MouseModel(params)
        init_weights(m.nn_module)
        with torch.no_grad():
            for k, v in m.nn_module.state_dict().items():
                if "running_mean" in k:
                    v.uniform_(-0.2, 0.2)
                if "running_var" in k:
                    v.uniform_(0.5, 1.5)
        path = tmp_path / f"fold_{fold}" / "model-001-0.100000.pth"
        path.parent.mkdir()
        m.save(path)
        paths.append(path)
        sds.append({k: v.detach().cpu() for k, v in m.nn_module.state_dict().items()})
    ens = EnsemblePredictor(paths, device="cuda:0", blend_weights="ones", window_batch=5)
    rs = np.random.RandomState(1)
    L = 39
    video = rs.randint(0, 256, (36, 64, L)).astype(np.uint8)
    beh, pup = rs.rand(2, L).astype(np.float32), rs.rand(2, L).astype(np.float32)
    got = ens.predict_trial(video, beh, pup, 0)
    assert got.shape == (9, L) and got.dtype == np.float32
    cfg = O.make_cfg(kw["readout_outputs"], **{k: v for k, v in kw.items() if k != "readout_outputs"})
    inputs = O.stack_inputs(torch.from_numpy(video), torch.from_numpy(beh), torch.from_numpy(pup))
    with torch.no_grad():
        per_model = [O.predict_trial(lambda c, sd=sd: O.dwiseneuro_forward(c, sd, cfg, 0, False), inputs, 9, 16, 2, "ones")
                     .numpy() for sd in sds]
    want = np.mean(per_model, axis=0)
    assert rel(torch.from_numpy(got), torch.from_numpy(want)) < FP32_TOL
    assert np.all(got[:, [1, 3]] == 0.0) == bool((want[:, [1, 3]] == 0).all())   # frames no window covers stay 0
    ens.set_precision("bf16")
    got16 = ens.predict_trial(video, beh, pup, 0)
    assert rel(torch.from_numpy(got16), torch.from_numpy(want)) < BF16_TOL
    # sharded form (single process: every trial on this rank)
    trials = [{"video": video, "behavior": beh, "pupil_center": pup, "mouse_index": 0},
              {"video": video[..., :35], "behavior": beh[:, :35], "pupil_center": pup[:, :35], "mouse_index": 1}]
    ens.set_precision("fp32")
    res = ens.predict_trials(trials)
    assert sorted(res) == [0, 1] and np.array_equal(res[0], got) and res[1].shape == (6, 35)


def test_train_step_compact_targets_equal_dense(dev):
    """train_step accepts (input, (responses (B, n_max, T), mouse_ids (B,))) and scatters it into the reference's dense
    batch form on the device (datasets.py:172-187): the step is bit-identical to the dense host batch, and the returned
    target has the dense form."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.synthetic import compact_from_dense
    from sensorium_b200.utils import init_weights

    def run(compact):
        params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **TINY_KW}), "loss": ("mice_poisson", {}),
                  "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
                  "iter_size": 1}
        torch.manual_seed(0)
        m = MouseModel(params)
        init_weights(m.nn_module)
        torch.manual_seed(4)
        losses, out = [], None
        for i in range(2):
            x = O.synthetic_clip(5, 16, 32, seed=30 + i)
            tg, w = O.synthetic_targets(5, TINY_OUTS, 16, seed=40 + i)
            if i == 1:                       # mouse 1 has no sample in the second batch
                tg[0] = tg[0] + tg[1][:, :37] * 0 + (w[:, 1] != 0).float()[:, None, None]
                w[:, 0] += w[:, 1]
                w[:, 1] = 0
                tg[1].zero_()
            target = compact_from_dense(tg, w) if compact else (tg, w)
            out = m.train_step((x, target), None)
            losses.append(out["loss"])
        return m, losses, out

    m1, l1, o1 = run(False)
    m2, l2, o2 = run(True)
    assert l1 == l2
    for a, b in zip(m1.nn_module.state_dict().values(), m2.nn_module.state_dict().values()):
        assert torch.equal(a, b)
    t1, w1 = o1["target"]
    t2, w2 = o2["target"]
    assert torch.equal(w1.to(dev), w2) and all(torch.equal(a.to(dev), b) for a, b in zip(t1, t2))


def test_cuda_graph_train_step_is_bit_identical_to_eager(dev):
    """params["cuda_graph"] = True: the train step is captured per (shapes, set of mice present) after one eager
    warm-up occurrence and replayed.  With the RNG-driven layers off (drop-path / dropout consume a different Philox
    offset under capture) every loss, weight, running statistic, optimizer moment and EMA entry must be bit-identical
    to the eager run — across batches with different mice present (several graphs), a learning-rate change between steps
    (the captured AdamW reads lr from device memory) and compact / dense batch forms."""
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.ema import ModelEma
    from sensorium_b200.synthetic import compact_from_dense
    from sensorium_b200.utils import init_weights
    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)

    def batches():
        out = []
        for i in range(8):
            x = O.synthetic_clip(4, 16, 32, seed=60 + i % 3)
            tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=70 + i % 3)
            if i % 2 == 1:                   # odd steps: mouse 2 absent -> a second graph
                keep = (w[:, 2] == 0).float()
                w = w * keep[:, None]
                w[:, 0] = torch.clamp(w[:, 0] + (1 - keep), max=1.0)
                tg[2].zero_()
            out.append((x, compact_from_dense(tg, w) if i >= 4 and bool((w.sum(1) == 1).all()) else (tg, w)))
        return out

    def run(graph, early=False):
        params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw}), "loss": ("mice_poisson", {}),
                  "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
                  "iter_size": 1, "cuda_graph": graph, "early_readout_step": early}
        torch.manual_seed(0)
        m = MouseModel(params)
        init_weights(m.nn_module)
        m.model_ema = ModelEma(m.nn_module, decay=0.9)
        losses = []
        for i, b in enumerate(batches()):
            if i == 5:
                m.set_lr(5e-4)
            losses.append(m.train_step(b, None)["loss"])
        return m, losses

    m1, l1 = run(False)
    m2, l2 = run(True)
    for graph in (False, True):   # the overlapped readout update (early AdamW + EMA on a side stream) is bit-identical too
        m4, l4 = run(graph, early=True)
        assert l4 == l1
        for (k, a), b in zip(m1.nn_module.state_dict().items(), m4.nn_module.state_dict().values()):
            assert torch.equal(a, b), k
        for (k, a), b in zip(m1.model_ema.ema.state_dict().items(), m4.model_ema.ema.state_dict().values()):
            assert torch.equal(a, b), k
    assert len(m2._graphs) >= 2 and not m1._graphs
    assert l1 == l2, (l1, l2)
    for (k, a), b in zip(m1.nn_module.state_dict().items(), m2.nn_module.state_dict().values()):
        assert torch.equal(a, b), k
    for (k, a), b in zip(m1.model_ema.ema.state_dict().items(), m2.model_ema.ema.state_dict().values()):
        assert torch.equal(a, b), k
    for p1, p2 in zip(m1.nn_module.parameters(), m2.nn_module.parameters()):
        s1, s2 = m1.optimizer.state[p1], m2.optimizer.state[p2]
        assert torch.equal(s1["exp_avg"], s2["exp_avg"]) and torch.equal(s1["exp_avg_sq"], s2["exp_avg_sq"])
        assert int(s1["step"]) == int(s2["step"])
    # gradients of the last step are visible on the parameters, absent mice keep grad=None
    assert all((p1.grad is None) == (p2.grad is None) for p1, p2 in zip(m1.nn_module.parameters(), m2.nn_module.parameters()))
    g1 = m1.nn_module.core.stem[0].weight.grad
    assert torch.equal(g1, m2.nn_module.core.stem[0].weight.grad)
    # with the stochastic layers ON the captured step still trains (fresh masks per replay: the losses move and differ)
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **TINY_KW}), "loss": ("mice_poisson", {}),
              "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
              "iter_size": 1, "cuda_graph": True}
    torch.manual_seed(0)
    m3 = MouseModel(params)
    init_weights(m3.nn_module)
    b0 = batches()[0]
    ls = [m3.train_step(b0, None)["loss"] for _ in range(6)]
    assert len(m3._graphs) == 1 and all(math.isfinite(v) for v in ls) and len(set(ls)) == 6 and ls[-1] < ls[0]


def test_train_fold_writes_a_reference_style_checkpoint(dev, tmp_path):
    """folds.train_fold (scripts/train.py:43-170 without the data pipeline): warm-up + cosine stages, validation with the
    device CorrelationMetric, one EMA checkpoint `model-{epoch:03d}-{val_corr:.6f}.pth` (max_saves=1) that Predictor /
    get_best_model_path pick up."""
    from sensorium_b200.folds import get_best_model_path, train_fold
    from sensorium_b200.predictors import Predictor
    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)
    cfg = {"batch_size": 4, "min_base_lr": 3e-6, "ema_decay": 0.9, "init_weights": True, "num_epochs": [1, 2],
           "stages": ["warmup", "train"],
           "argus_params": {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw}),
                            "loss": ("mice_poisson", {}), "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}),
                            "device": "cuda:0", "amp": True, "iter_size": 1, "cuda_graph": True,
                            "frame_stack": {"size": 16, "step": 2, "position": "last"},
                            "inputs_processor": ("stack_inputs", {"size": (32, 32), "pad_fill_value": 0.0}),
                            "responses_processor": ("identity", {})}}
    train = [(O.synthetic_clip(4, 16, 32, seed=i), O.synthetic_targets(4, TINY_OUTS, 16, seed=100 + i)) for i in range(3)]
    val = [(O.synthetic_clip(4, 16, 32, seed=50), O.synthetic_targets(4, TINY_OUTS, 16, seed=150))]
    logs = []
    torch.manual_seed(0)
    best = train_fold(cfg, tmp_path / "fold_0", train, val, log=logs.append)
    files = sorted((tmp_path / "fold_0").glob("*.pth"))
    assert len(files) == 1 and files[0] == best == get_best_model_path(tmp_path / "fold_0")
    assert best.name.startswith("model-00") and len(logs) == 3 and "val_corr" in logs[-1]
    ckpt = torch.load(best, map_location="cpu", weights_only=False)
    assert set(ckpt) == {"model_name", "params", "nn_state_dict"} and ckpt["model_name"] == "MouseModel"
    pr = Predictor(best, device="cuda:0")
    out = pr.model.predict(O.synthetic_clip(1, 16, 32, seed=3), 0)
    assert out.shape == (1, TINY_OUTS[0], 16) and bool(torch.isfinite(out).all())


def test_train_step_raw_clips_equal_dense(dev):
    """The raw batch form — (video (B, T, Hv, Wv) uint8, scalars (B, 4, T)) — is padded / stacked on the device
    (dwn_assemble_batch == StackInputsProcessor per sample, inputs.py:22-36): bit-exact clips, and a train step (eager
    and captured) that is bit-identical to the dense form."""
    from sensorium_b200._lib import call
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.synthetic import compact_from_dense, raw_from_dense
    from sensorium_b200.utils import init_weights
    x = O.synthetic_clip(3, 16, 32, seed=5)
    video, scal = raw_from_dense(x)
    assert video.dtype == torch.uint8 and video.shape == (3, 16, 18, 32) and scal.shape == (3, 4, 16)
    want = torch.stack([O.stack_inputs(video[b].permute(1, 2, 0), scal[b, :2], scal[b, 2:], size=(32, 32)) for b in range(3)])
    assert torch.equal(want, x)
    out = torch.empty((3, 5, 16, 32, 32), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    call("dwn_assemble_batch", video.to(dev), 2, scal.to(dev), out, 3, 16, 18, 32, 32, 32, 0.0, st)
    assert torch.equal(out.cpu(), x)
    vf = video.float() + 0.25                                     # fp32 video and a non-zero fill value
    call("dwn_assemble_batch", vf.to(dev), 0, scal.to(dev), out, 3, 16, 18, 32, 32, 32, 7.5, st)
    want = torch.stack([O.stack_inputs(vf[b].permute(1, 2, 0), scal[b, :2], scal[b, 2:], size=(32, 32), fill=7.5) for b in range(3)])
    assert torch.equal(out.cpu(), want)

    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)

    def run(raw, graph):
        params = {"nn_module": ("dwiseneuro", {"readout_outputs": TINY_OUTS, **kw}), "loss": ("mice_poisson", {}),
                  "optimizer": ("AdamW", {"lr": 2e-3, "weight_decay": 0.05}), "device": "cuda:0", "amp": True,
                  "iter_size": 1, "cuda_graph": graph,
                  "inputs_processor": ("stack_inputs", {"size": (32, 32), "pad_fill_value": 0.0})}
        torch.manual_seed(0)
        m = MouseModel(params)
        init_weights(m.nn_module)
        losses = []
        for i in range(3):
            xb = O.synthetic_clip(4, 16, 32, seed=80 + i)
            tg, w = O.synthetic_targets(4, TINY_OUTS, 16, seed=90)
            losses.append(m.train_step((raw_from_dense(xb) if raw else xb, compact_from_dense(tg, w)), None)["loss"])
        return m, losses

    m0, l0 = run(False, False)
    for raw, graph in ((True, False), (True, True)):
        m1, l1 = run(raw, graph)
        assert l0 == l1, (raw, graph, l0, l1)
        for a, b in zip(m0.nn_module.state_dict().values(), m1.nn_module.state_dict().values()):
            assert torch.equal(a, b)


@pytest.mark.parametrize("HW", [30, 27])
def test_ceil_shortcut_for_sizes_the_stride_does_not_divide(dev, HW):
    """interpolate_shortcut (dwiseneuro.py:125-129) resizes the shortcut to ceil(size / stride) with F.interpolate
    (mode="nearest": source index floor(dst * in / out)), and the strided depth-wise conv (k=3, pad=1) produces the same
    ceil size.  30 -> 15 -> 15 -> 8 exercises it in the last block, 27 -> 14 -> 14 -> 7 in the first and the last; train
    mode forward, loss and every gradient against the oracle (fp32 <= 1e-4; bf16 outputs <= 2e-2)."""
    x = O.synthetic_clip(3, 8, 32, seed=4)[..., :HW, :HW].contiguous().to(dev)
    tg, w = O.synthetic_targets(3, TINY_OUTS, 8, seed=6)
    tg, w = [t.to(dev) for t in tg], w.to(dev)
    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)
    cfg = O.make_cfg(TINY_OUTS, **kw)
    for mode, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        from sensorium_b200 import DwiseNeuro
        from sensorium_b200.utils import init_weights
        torch.manual_seed(1)
        net = DwiseNeuro(readout_outputs=TINY_OUTS, **kw)
        init_weights(net)
        net = net.to(dev).train()
        net.precision = mode
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        names = [k for k, _ in net.named_parameters()]
        for k in names:
            sd[k].requires_grad_(True)
        ref = O.dwiseneuro_forward(x, sd, cfg, None, True)
        ref_loss = O.mice_poisson_loss(ref, tg, w)
        ref_loss.backward()
        out = net(x)
        loss = O.mice_poisson_loss(out, tg, w)
        loss.backward()
        for a, b in zip(out, ref):
            assert a.shape == b.shape and rel(a, b) < tol, (mode, rel(a, b))
        if mode == "fp32":
            for k, v in net.state_dict().items():
                if "running_var" in k:
                    assert rel(v, sd[k]) < tol, k
                elif "running_mean" in k:  # a mean is judged on the scale of the channel's standard deviation
                    scale = max(float(sd[k].abs().max()), float(sd[k.replace("running_mean", "running_var")].sqrt().max()))
                    assert float((v - sd[k]).abs().max()) < tol * scale, k
            gmax = max(float(sd[k].grad.abs().max()) for k in names if sd[k].grad is not None)
            for k, p in net.named_parameters():
                if sd[k].grad is None:
                    assert p.grad is None
                    continue
                err = float((p.grad - sd[k].grad).abs().max())
                assert err <= tol * max(float(sd[k].grad.abs().max()), 3e-3 * gmax) + 1e-6 * gmax, (k, err)
    # eval mode too (the predictor path), one readout
    net.eval()
    net.precision = "fp32"
    with torch.no_grad():
        sd_e = {k: v.detach().clone() for k, v in net.state_dict().items()}
        assert rel(net(x, 1), O.dwiseneuro_forward(x, sd_e, cfg, 1, False)) < FP32_TOL


def test_fused_adamw_state_dict_roundtrip(dev):
    """FusedAdamW.state_dict() carries only torch's own per-parameter entries (exp_avg, exp_avg_sq, step) — no raw
    pointers, no launch caches — and an optimizer restored with load_state_dict continues bit-identically (the flat
    moment buffers are rebuilt from the loaded state)."""
    from sensorium_b200.optim import FusedAdamW
    torch.manual_seed(0)

    def params():
        g = torch.Generator().manual_seed(3)
        return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in ((33, 7), (5,), (64, 16, 1))]

    def grads(step):
        g = torch.Generator().manual_seed(100 + step)
        return [torch.randn(s, generator=g).to(dev) for s in ((33, 7), (5,), (64, 16, 1))]

    pa = params()
    oa = FusedAdamW(pa, lr=1e-2, weight_decay=0.05)
    for st in range(3):
        for p, g in zip(pa, grads(st)):
            p.grad = g if not (st == 1 and p.dim() == 1) else None   # one tensor skips a step
        oa.step()
    sd = oa.state_dict()
    assert set(sd["param_groups"][0]) == {"lr", "betas", "eps", "weight_decay", "params"} | (
        set(sd["param_groups"][0]) & {"foreach", "maximize", "capturable", "differentiable", "fused"})
    for ent in sd["state"].values():
        assert set(ent) == {"exp_avg", "exp_avg_sq", "step"}
    assert [int(sd["state"][i]["step"]) for i in range(3)] == [3, 2, 3]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ob = FusedAdamW(pb, lr=1e-2, weight_decay=0.05)
    ob.load_state_dict({"state": {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in sd["state"].items()},
                        "param_groups": sd["param_groups"]})
    for st in range(3, 5):
        for p, q, g in zip(pa, pb, grads(st)):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    for p, q in zip(pa, pb):
        assert torch.equal(p, q)
        assert torch.equal(oa.state[p]["exp_avg_sq"], ob.state[q]["exp_avg_sq"]) and int(ob.state[q]["step"]) == int(oa.state[p]["step"])
