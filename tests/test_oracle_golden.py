"""CPU: the oracle (oracle/dwiseneuro_oracle.py) against the golden fixtures generated from the real reference
(tests/golden/make_golden.py).  Bit-exact where the arithmetic is identical, 1e-6 otherwise."""
import json

import pytest
import torch

from oracle import dwiseneuro_oracle as O
from sensorium_b200 import DwiseNeuro, constants
from sensorium_b200.utils import init_weights


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def test_tiny_forward_loss_grads(golden_dir):
    g = torch.load(golden_dir / "tiny_forward_backward.pt", weights_only=False)
    cfgkw = dict(g["cfg"])
    outs = cfgkw.pop("readout_outputs")
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=outs, **cfgkw)
    init_weights(net)
    cfg = O.make_cfg(outs, **cfgkw)
    x = O.synthetic_clip(4, 16, 32, seed=0)
    tg, w = O.synthetic_targets(4, outs, 16, seed=1)
    with torch.no_grad():
        ev = O.dwiseneuro_forward(x, _sd(net), cfg, None, False)
    for a, b in zip(ev, g["eval_out"]):
        assert torch.equal(a, b)
    sd = _sd(net)
    names = [k for k, _ in net.named_parameters()]
    for k in names:
        sd[k].requires_grad_(True)
    torch.manual_seed(5)
    tr = O.dwiseneuro_forward(x, sd, cfg, None, True)
    for a, b in zip(tr, g["train_out"]):
        assert torch.equal(a.detach(), b)
    loss = O.mice_poisson_loss(tr, tg, w)
    assert torch.equal(loss.detach(), g["loss"])
    loss.backward()
    for k in names:
        if k in g["none_grads"]:
            assert sd[k].grad is None
        else:
            torch.testing.assert_close(sd[k].grad, g["grads"][k], rtol=1e-5, atol=1e-7)
    for k, v in g["running"].items():
        assert torch.equal(sd[k], v), k


def test_c1_full_architecture_forward(golden_dir):
    g = torch.load(golden_dir / "c1_forward_index0.pt", weights_only=False)
    from tests.shapes import TRUE_BATCH_KW
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW)
    init_weights(net)
    assert len(net.state_dict()) == g["n_state_entries"] == 365
    assert sum(p.numel() for p in net.parameters()) == g["n_params"] == 170656070
    cfg = O.make_cfg(constants.num_neurons, **TRUE_BATCH_KW)
    x = O.synthetic_clip(1, 16, 64, seed=0)
    with torch.no_grad():
        y = O.dwiseneuro_forward(x, _sd(net), cfg, 0, False)
    assert y.shape == (1, 7863, 16)
    assert torch.equal(y, g["out_index0"])


def test_index_facts(golden_dir):
    f = json.loads((golden_dir / "index_facts.json").read_text())
    assert O.shuffle_source_index(8, 2) == f["shuffle_8_g2"] == [0, 4, 1, 5, 2, 6, 3, 7]
    assert O.tile_source_index(3, 8) == f["tile_3_to_8"]
    assert O.tile_source_index(4, 10) == f["tile_4_to_10"]
    assert O.nearest_source_index(5, 3) == f["nearest_5_to_3"] == [0, 1, 3]
    assert O.nearest_source_index(64, 32) == list(range(0, 64, 2))
    assert O.make_window_indexes(40, 16, 2, "last") == f["indexes_16_2_last"]["at_40"]
    assert O.make_window_indexes(40, 7, 3, "middle") == f["indexes_7_3_middle"]["at_40"]
    assert abs(float(torch.nn.functional.softplus(torch.zeros(1), beta=0.07)) - f["softplus_beta007_at_0"]) < 1e-6


def test_positional_encoding(golden_dir):
    pes = torch.load(golden_dir / "positional_encoding.pt", weights_only=False)
    for c, ent in pes.items():
        c = int(c)
        ch = -(-c // 6) * 2
        inv = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
        enc = O.positional_encoding(c, tuple(ent["shape"]), inv, torch.float32)[0]
        assert torch.equal(enc[:, :, 0, 0], ent["t"]) and torch.equal(enc[:, 0, :, 0], ent["h"])
        assert torch.equal(enc[:, 0, 0, :], ent["w"])
        assert abs(float(enc.double().sum()) - ent["checksum"]) < 1e-9 * max(1.0, abs(ent["checksum"]))


def test_ema_adamw_distill(golden_dir):
    g = torch.load(golden_dir / "ema_update.pt", weights_only=False)
    ema = {k: v.clone() for k, v in g["before"].items()}
    O.ema_update(ema, g["model"], g["decay"])
    for k in ema:
        assert torch.equal(ema[k], g["after"][k]), k
    a = torch.load(golden_dir / "adamw_steps.pt", weights_only=False)
    p, m, v = a["p0"].clone(), torch.zeros(300), torch.zeros(300)
    for i, gr in enumerate(a["grads"]):
        O.adamw_step(p, gr, m, v, i + 1, a["lr"], a["wd"])
        torch.testing.assert_close(p, a["traj"][i], rtol=1e-6, atol=1e-7)
    d = torch.load(golden_dir / "distill_fill.pt", weights_only=False)
    tg = [t.clone() for t in d["targets_in"]]
    w = d["weights_in"].clone()
    O.distill_fill(tg, w, d["teacher"], d["ratio"])
    assert torch.equal(w, d["weights_out"])
    for x, y in zip(tg, d["targets_out"]):
        assert torch.equal(x, y)


def test_predictor_blend_and_inputs(golden_dir):
    g = torch.load(golden_dir / "predictor_blend.pt", weights_only=False)
    stacked = O.stack_inputs(g["video"], g["behavior"], g["pupil"])
    assert abs(float(stacked.double().sum()) - g["stacked_checksum"]) < 1e-6
    assert torch.equal(stacked[:, 5, 10:54:7, ::9], g["stacked_slice"])
    n_out = g["n_out"]

    def fake(inp):
        feat = inp[0, :, :, 20:24, 30:34].mean((0, 2, 3))
        return (feat[None, None, :] * torch.arange(1, n_out + 1)[None, :, None]).float()

    for blend in ("ones", "linear"):
        r = O.predict_trial(fake, stacked, n_out, 16, 2, blend)
        torch.testing.assert_close(r, g["responses"][blend], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        O.predict_trial(fake, stacked, n_out, 16, 2, "cosine")


def test_correlation_metric(golden_dir):
    """oracle.correlation_metric == the reference's CorrelationMetric (metrics.py:34-74) on the shared batches."""
    from tests.shapes import corr_step_outputs
    g = torch.load(golden_dir / "corr_metric.pt", weights_only=False)
    r = O.correlation_metric(corr_step_outputs())
    assert set(r["mice_corr"]) == set(g["mice_corr"]) == {0, 2}  # mouse 1 never had a sample
    for m, v in g["mice_corr"].items():
        assert abs(r["mice_corr"][m] - v) < 1e-6
        assert float((torch.from_numpy(r["per_neuron"][m]) - g["per_neuron"][m]).abs().max()) < 2e-6


def test_cutmix_host_decisions_and_oracle(golden_dir):
    """The host side of the device CutMix draws the reference's random numbers in the reference's order
    (mixers.py:13-14, 36-49, 57-66): same seed -> same samples mixed, same boxes, and oracle.cutmix reproduces the
    reference's mixed sample bit for bit."""
    import numpy as np
    from sensorium_b200.mixers import DeviceCutMix
    g = torch.load(golden_dir / "cutmix.pt", weights_only=False)
    recs = g["records"]
    np.random.seed(g["seed"])
    boxes, lams = DeviceCutMix(g["alpha"], g["prob"]).sample(len(recs), 16, 16)
    for b, r in enumerate(recs):
        used = bool(boxes[b].any()) or float(lams[b]) != 0.0
        if not r["used"]:
            assert not used and torch.equal(r["out"][0], r["s1"][0])
            continue
        x, t = O.cutmix(r["s1"][0], r["s1"][1], r["s2"][0], r["s2"][1], boxes[b])
        assert torch.equal(x, r["out"][0]) and torch.equal(t, r["out"][1])
        area = (boxes[b][2] - boxes[b][0]) * (boxes[b][3] - boxes[b][1]) / 256.0
        assert abs(float(lams[b]) - area) < 1e-7
