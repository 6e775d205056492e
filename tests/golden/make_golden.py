"""Generates the golden fixtures in tests/golden/ by importing the REAL reference from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
pytorch-argus is absent here, so a stub exposing the names the reference imports is injected (it performs no
arithmetic).  Fixtures are small (< 2 MB total) and committed together with this script.
"""
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
ROOT = OUT.parent.parent
sys.path.insert(0, str(ROOT))


def install_argus_stub():
    argus = types.ModuleType("argus")

    class Model:
        def __init__(self, params):
            self.params = params
            self.device = torch.device("cpu")
            self.prediction_transform = lambda x: x

        def train(self):
            self.nn_module.train()

        def eval(self):
            self.nn_module.eval()

        def _check_predict_ready(self):
            pass

    argus.Model = Model
    argus.load_model = lambda *a, **k: None
    engine = types.ModuleType("argus.engine")
    engine.State = type("State", (), {})
    loss = types.ModuleType("argus.loss")
    loss.pytorch_losses = {}
    utils = types.ModuleType("argus.utils")

    def deep_to(x, device=None, **kw):
        if torch.is_tensor(x):
            return x.to(device)
        if isinstance(x, (list, tuple)):
            return [deep_to(v, device) for v in x]
        return x

    def deep_detach(x):
        if torch.is_tensor(x):
            return x.detach()
        if isinstance(x, (list, tuple)):
            return [deep_detach(v) for v in x]
        return x

    utils.deep_to, utils.deep_detach = deep_to, deep_detach
    utils.deep_chunk = lambda x, n: [x]
    cb = types.ModuleType("argus.callbacks")
    cb.Checkpoint = type("Checkpoint", (), {})
    metrics = types.ModuleType("argus.metrics")
    metrics.Metric = type("Metric", (), {})
    for name, mod in {"argus": argus, "argus.engine": engine, "argus.loss": loss, "argus.utils": utils,
                      "argus.callbacks": cb, "argus.metrics": metrics}.items():
        sys.modules[name] = mod
    argus.engine, argus.loss, argus.utils, argus.callbacks, argus.metrics = engine, loss, utils, cb, metrics


def gen_corr_metric():
    """metrics.py:34-74 run by the real reference class on the batches above."""
    from tests.shapes import corr_step_outputs
    RM = importlib.import_module("src.metrics")
    steps = corr_step_outputs()
    metric = RM.CorrelationMetric()
    metric.reset()
    for s in steps:
        metric.update(s)
    res = metric.compute()
    per_neuron = {}
    for m in metric.predictions:
        t = np.concatenate(metric.targets[m], axis=0)
        p = np.concatenate(metric.predictions[m], axis=0)
        per_neuron[int(m)] = torch.from_numpy(RM.corr(p, t, axis=0).astype(np.float64))
    torch.save({"mice_corr": {int(k): float(v) for k, v in res.items()}, "per_neuron": per_neuron},
               OUT / "corr_metric.pt")
    print("corr_metric.pt:", {int(k): float(v) for k, v in res.items()})


def gen_cutmix():
    """mixers.py:10-67 run by the real reference CutMix (numpy RNG seeded) on six small sample pairs."""
    RX = importlib.import_module("src.mixers")
    g = torch.Generator().manual_seed(21)
    mixer = RX.CutMix(alpha=1.0, prob=0.5)
    np.random.seed(1234)
    recs = []
    for i in range(6):
        s1 = (torch.rand(2, 3, 16, 16, generator=g), torch.rand(7, 3, generator=g))
        s2 = (torch.rand(2, 3, 16, 16, generator=g), torch.rand(7, 3, generator=g))
        used = mixer.use()
        if used:
            out = mixer(s1, s2)
        else:
            out = s1
        recs.append({"s1": s1, "s2": s2, "used": bool(used), "out": (out[0].clone(), out[1].clone())})
    torch.save({"seed": 1234, "alpha": 1.0, "prob": 0.5, "records": recs}, OUT / "cutmix.pt")
    print("cutmix.pt: used =", [r["used"] for r in recs])


def main():
    install_argus_stub()
    sys.path.insert(0, str(REF))
    if "--only-corr" in sys.argv:
        gen_corr_metric()
        return
    if "--only-cutmix" in sys.argv:
        gen_cutmix()
        return
    R = importlib.import_module("src.models.dwiseneuro")
    RL = importlib.import_module("src.losses")
    RU = importlib.import_module("src.utils")
    RI = importlib.import_module("src.indexes")
    RE = importlib.import_module("src.ema")
    RA = importlib.import_module("src.argus_models")
    RIN = importlib.import_module("src.inputs")
    from oracle import dwiseneuro_oracle as O
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.utils import init_weights

    torch.set_num_threads(8)
    # ------------------------------------------------------------------ tiny config: fwd (eval/train), loss, grads
    outs = (37, 64, 129)
    kw = dict(core_features=(16, 16, 32), spatial_strides=(2, 1, 2), expansion_ratio=4, se_reduce_ratio=8,
              cortex_features=(64, 128), groups=2, drop_path_rate=0.3)
    torch.manual_seed(0)
    ref = R.DwiseNeuro(readout_outputs=outs, **kw)
    RU.init_weights(ref)
    torch.manual_seed(0)
    mine = DwiseNeuro(readout_outputs=outs, **kw)
    init_weights(mine)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), f"parameter tree / init RNG stream mismatch at {ka}"
    x = O.synthetic_clip(4, 16, 32, seed=0)
    tg, w = O.synthetic_targets(4, outs, 16, seed=1)
    ref.eval()
    with torch.no_grad():
        ev = ref(x)
    ref.train()
    torch.manual_seed(5)
    tr = ref(x)
    loss = RL.MicePoissonLoss()(tr, (tg, w))
    loss.backward()
    grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    none_grads = [k for k, p in ref.named_parameters() if p.grad is None]
    torch.save({"cfg": dict(readout_outputs=outs, **kw), "eval_out": [t.clone() for t in ev],
                "train_out": [t.detach().clone() for t in tr], "loss": loss.detach().clone(), "grads": grads,
                "none_grads": none_grads,
                "running": {k: v.clone() for k, v in ref.state_dict().items() if "running" in k or "tracked" in k}},
               OUT / "tiny_forward_backward.pt")

    # ------------------------------------------------------------------ C1: full architecture, batch 1, index 0, eval
    from sensorium_b200 import constants
    big = dict(in_channels=5, core_features=(64, 64, 64, 64, 128, 128, 128, 256, 256),
               spatial_strides=(2, 1, 1, 1, 2, 1, 1, 2, 1), spatial_kernel=3, temporal_kernel=5, expansion_ratio=7,
               se_reduce_ratio=32, cortex_features=(1024, 2048, 4096), groups=2, softplus_beta=0.07, drop_rate=0.4,
               drop_path_rate=0.1)
    torch.manual_seed(0)
    refb = R.DwiseNeuro(readout_outputs=constants.num_neurons, **big)
    RU.init_weights(refb)
    torch.manual_seed(0)
    mineb = DwiseNeuro(readout_outputs=constants.num_neurons, **big)
    init_weights(mineb)
    n_entries = 0
    for (ka, va), (kb, vb) in zip(refb.state_dict().items(), mineb.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
        n_entries += 1
    assert n_entries == 365 and sum(p.numel() for p in refb.parameters()) == 170656070
    x1 = O.synthetic_clip(1, 16, 64, seed=0)
    refb.eval()
    with torch.no_grad():
        y1 = refb(x1, 0)
    torch.save({"out_index0": y1.clone(), "n_state_entries": n_entries,
                "n_params": sum(p.numel() for p in refb.parameters())}, OUT / "c1_forward_index0.pt")
    del refb, mineb

    # ------------------------------------------------------------------ integer index work / PE / softplus
    facts = {}
    sl = R.ShuffleLayer(8, 8, groups=2)
    t = torch.arange(8.0).view(1, 8, 1)
    facts["shuffle_8_g2"] = sl.shuffle_channels(t).flatten().int().tolist()
    sl2 = R.ShuffleLayer(3, 8, groups=1)
    sl2.bn_sc = torch.nn.Identity()
    facts["tile_3_to_8"] = sl2.tile_shortcut(torch.arange(3.0).view(1, 3, 1)).flatten().int().tolist()
    blk = R.InvertedResidual3d(4, 10, spatial_stride=2)
    blk.bn_sc = torch.nn.Identity()
    src = torch.arange(5.0).view(1, 1, 1, 5, 1).expand(1, 4, 1, 5, 5)
    facts["nearest_5_to_3"] = blk.interpolate_shortcut(src)[0, 0, 0, :, 0].int().tolist()
    cidx = torch.arange(4.0).view(1, 4, 1, 1, 1).expand(1, 4, 1, 2, 2)
    facts["tile_4_to_10"] = blk.interpolate_shortcut(cidx)[0, :, 0, 0, 0].int().tolist()
    g = RI.IndexesGenerator(16, 2, "last")
    facts["indexes_16_2_last"] = dict(behind=g.behind, ahead=g.ahead, width=g.width, at_40=g.make_indexes(40))
    g2 = RI.IndexesGenerator(7, 3, "middle")
    facts["indexes_7_3_middle"] = dict(behind=g2.behind, ahead=g2.ahead, width=g2.width, at_40=g2.make_indexes(40),
                                       clip=[g2.clip_index(i, 50, 2) for i in (0, 20, 49)])
    facts["softplus_beta007_at_0"] = float(torch.nn.Softplus(beta=0.07)(torch.zeros(1)))
    facts["drop_path_rates_core"] = [0.1 * i / 9 for i in range(9)]
    (OUT / "index_facts.json").write_text(json.dumps(facts, indent=1))
    pes = {}
    for c, shape in ((64, (16, 64, 64)), (128, (16, 32, 32)), (256, (16, 8, 8)), (16, (8, 4, 6))):
        pe = R.PositionalEncoding3d(c)
        enc = pe.create_cached_encoding(torch.zeros(1, c, *shape))
        # store the three separable axis profiles (exactly what the full tensor is made of)
        pes[str(c)] = dict(shape=shape, t=enc[0, :, :, 0, 0].clone(), h=enc[0, :, 0, :, 0].clone(),
                           w=enc[0, :, 0, 0, :].clone(), checksum=float(enc.double().sum()))
    torch.save(pes, OUT / "positional_encoding.pt")

    # ------------------------------------------------------------------ EMA (incl. int64 truncation), AdamW, distill fill
    torch.manual_seed(1)
    small = R.DwiseNeuro(readout_outputs=(5, 4), core_features=(8,), spatial_strides=(1,), expansion_ratio=2,
                         se_reduce_ratio=4, cortex_features=(8,), groups=2)
    ema = RE.ModelEma(small, decay=0.9)
    before = {k: v.clone() for k, v in small.state_dict().items()}
    with torch.no_grad():
        for k, v in small.state_dict().items():
            if v.dtype == torch.int64:
                v.fill_(25)
            else:
                v.add_(torch.randn_like(v))
    ema.update(small)
    torch.save({"before": before, "model": {k: v.clone() for k, v in small.state_dict().items()},
                "after": {k: v.clone() for k, v in ema.ema.state_dict().items()}, "decay": 0.9}, OUT / "ema_update.pt")

    torch.manual_seed(2)
    p0 = torch.randn(300)
    gs = [torch.randn(300) for _ in range(3)]
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p], lr=2.4e-3, weight_decay=0.05)
    traj = []
    for gi in gs:
        p.grad = gi.clone()
        opt.step()
        traj.append(p.detach().clone())
    torch.save({"p0": p0, "grads": gs, "traj": traj, "lr": 2.4e-3, "wd": 0.05}, OUT / "adamw_steps.pt")

    mm = RA.MouseModel.__new__(RA.MouseModel)
    mm.distill_ratio = 0.36
    torch.manual_seed(3)
    teacher = [torch.rand(6, n, 4) for n in (5, 4, 3)]
    mm.distill_model = lambda inp: teacher
    tgt = [torch.rand(6, n, 4) for n in (5, 4, 3)]
    mw = torch.nn.functional.one_hot(torch.tensor([0, 2, 1, 1, 0, 2]), 3).float()
    tgt_in = [t.clone() for t in tgt]
    mw_in = mw.clone()
    mm.add_distill_predictions(None, (tgt, mw))
    torch.save({"teacher": teacher, "targets_in": tgt_in, "weights_in": mw_in, "targets_out": tgt, "weights_out": mw,
                "ratio": 0.36}, OUT / "distill_fill.pt")

    # ------------------------------------------------------------------ predictor overlap-add + input stacking
    RP = importlib.import_module("src.predictors")
    torch.manual_seed(4)
    L, n_out = 47, 6
    video = np.random.RandomState(0).randint(0, 256, (36, 64, L)).astype(np.uint8)
    beh = np.random.RandomState(1).rand(2, L).astype(np.float32)
    pup = np.random.RandomState(2).rand(2, L).astype(np.float32)
    proc = RIN.StackInputsProcessor(size=(64, 64), pad_fill_value=0.)
    stacked = proc(video, beh, pup)
    res = {}
    for blend in ("ones", "linear"):
        pr = RP.Predictor.__new__(RP.Predictor)
        pr.inputs_processor = proc
        pr.frame_stack_size, pr.frame_stack_step = 16, 2
        pr.indexes_generator = RI.IndexesGenerator(16, 2)
        pr.blend_weights = RP.get_blend_weights(blend, 16)

        class FakeModel:
            device = "cpu"

            def predict(self, inp, mouse_index):
                # deterministic "network": depends on the window content only
                feat = inp[0, :, :, 20:24, 30:34].mean((0, 2, 3))                    # (16,)
                return (feat[None, None, :] * torch.arange(1, n_out + 1)[None, :, None]).float()

        pr.model = FakeModel()
        RP.constants.num_neurons = [n_out] * 10
        res[blend] = torch.from_numpy(pr.predict_trial(video, beh, pup, 0).copy())
    torch.save({"video": torch.from_numpy(video), "behavior": torch.from_numpy(beh), "pupil": torch.from_numpy(pup),
                "stacked_checksum": float(stacked.double().sum()), "stacked_slice": stacked[:, 5, 10:54:7, ::9].clone(),
                "responses": res, "n_out": n_out}, OUT / "predictor_blend.pt")
    gen_corr_metric()
    gen_cutmix()
    print("golden fixtures written to", OUT)
    for f in sorted(OUT.glob("*")):
        print(f"  {f.name:32s} {f.stat().st_size / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
