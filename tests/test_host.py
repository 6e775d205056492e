"""CPU: host-side logic, the drop-in parameter tree, and that libdwn_b200.so loads and exports every symbol that
include/dwn_b200.h declares (no compute calls without a GPU)."""
import copy
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from sensorium_b200 import DwiseNeuro, _lib, constants
from sensorium_b200.indexes import IndexesGenerator
from sensorium_b200.inputs import StackInputsProcessor, get_inputs_processor
from sensorium_b200.predictors import get_blend_weights
from sensorium_b200.utils import get_lr, init_weights
from tests.shapes import TINY_KW, TINY_OUTS

ROOT = Path(__file__).resolve().parent.parent


def test_cabi_exports_every_declared_symbol():
    header = (ROOT / "include" / "dwn_b200.h").read_text()
    declared = set(re.findall(r"\b(dwn_[a-z0-9_]+)\s*\(", header))
    declared.discard("dwn_gemm_desc")
    assert len(declared) >= 45
    handle = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert handle.dwn_abi_version() == 1
    # every bound symbol is declared in the header
    assert set(_lib.exported_symbols()) <= declared | {"dwn_last_error"}


def test_parameter_tree_matches_reference_names(golden_dir):
    g = torch.load(golden_dir / "tiny_forward_backward.pt", weights_only=False)
    net = DwiseNeuro(readout_outputs=TINY_OUTS, **TINY_KW)
    names = [k for k, _ in net.named_parameters()]
    assert set(names) == set(g["grads"]) | set(g["none_grads"])
    sd = net.state_dict()
    assert "core.blocks.1.spat_covn_dw.0.weight" in sd and "core.blocks.1.temp_covn_dw.0.weight" in sd
    assert "core.blocks.0.inv_freq" in sd and not any("cached_encoding" in k for k in sd)
    assert set(g["running"]) <= set(sd)
    # ModelEma relies on deepcopy + order-stable state_dict (ema.py:40,49)
    twin = copy.deepcopy(net)
    assert list(twin.state_dict()) == list(sd)
    twin.load_state_dict(sd)


def test_no_cpu_fallback():
    net = DwiseNeuro(readout_outputs=TINY_OUTS, **TINY_KW)
    with pytest.raises(RuntimeError, match="5D"):
        net(torch.zeros(5, 16, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 5, 16, 32, 32))
    with pytest.raises(AssertionError):
        DwiseNeuro(readout_outputs=(3,), core_features=(8, 8), spatial_strides=(1,))


def test_indexes_generator(golden_dir):
    import json
    f = json.loads((golden_dir / "index_facts.json").read_text())
    g = IndexesGenerator(16, 2, "last")
    assert (g.behind, g.ahead, g.width) == (30, 0, 31)
    assert g.make_indexes(40) == f["indexes_16_2_last"]["at_40"]
    g2 = IndexesGenerator(7, 3, "middle")
    m = f["indexes_7_3_middle"]
    assert (g2.behind, g2.ahead, g2.width) == (m["behind"], m["ahead"], m["width"])
    assert g2.make_indexes(40) == m["at_40"]
    assert [g2.clip_index(i, 50, 2) for i in (0, 20, 49)] == m["clip"]
    with pytest.raises(ValueError):
        IndexesGenerator(4, 1, "center")
    assert len(range(g.behind, 300 - g.ahead)) == 270  # windows of a 300-frame trial (predictors.py:46-49)


def test_inputs_processor_and_misc(golden_dir):
    g = torch.load(golden_dir / "predictor_blend.pt", weights_only=False)
    proc = get_inputs_processor("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0})
    assert isinstance(proc, StackInputsProcessor)
    out = proc(g["video"].numpy(), g["behavior"].numpy(), g["pupil"].numpy())
    assert out.shape == (5, 47, 64, 64) and out.dtype == torch.float32
    assert abs(float(out.double().sum()) - g["stacked_checksum"]) < 1e-6
    assert torch.equal(out[:, 5, 10:54:7, ::9], g["stacked_slice"])
    assert float(out[0, :, :14].abs().sum()) == 0.0 and float(out[0, :, 50:].abs().sum()) == 0.0
    assert get_lr(3e-4, 32) == pytest.approx(2.4e-3)
    assert np.array_equal(get_blend_weights("ones", 16), np.ones(16, np.float32))
    assert np.allclose(get_blend_weights("linear", 16), np.linspace(0, 1, 16))
    with pytest.raises(ValueError):
        get_blend_weights("cosine", 16)
    assert constants.num_neurons == [7863, 7908, 8202, 7939, 8122, 7440, 7928, 8285, 7671, 7495]


def test_init_weights_statistics():
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=(64,), core_features=(16,), spatial_strides=(1,), expansion_ratio=4,
                     se_reduce_ratio=8, cortex_features=(64,), groups=2)
    init_weights(net)
    w = net.core.blocks[1].conv_pw[0].weight
    assert abs(float(w.std()) - (2.0 / 64) ** 0.5) < 0.02  # fan_out = 1*64
    assert float(net.core.stem[1].bn.weight.min()) == 1.0 and float(net.readouts[0].layer[1].bias.abs().max()) == 0.0


def test_argus_shim_checkpoint_roundtrip(tmp_path):
    from sensorium_b200 import argus_shim
    from sensorium_b200.argus_models import MouseModel
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": (5, 4), "core_features": (8,), "spatial_strides": (1,),
                                           "expansion_ratio": 2, "se_reduce_ratio": 4, "cortex_features": (8,)}),
              "loss": ("mice_poisson", {}), "optimizer": ("AdamW", {"lr": 1e-3}), "device": "cpu",
              "frame_stack": {"size": 16, "step": 2, "position": "last"},
              "inputs_processor": ("stack_inputs", {"size": (64, 64), "pad_fill_value": 0.0}),
              "responses_processor": ("identity", {}), "amp": True, "iter_size": 1}
    m = MouseModel(params)
    assert m.iter_size == 1 and m.amp is True and m.model_ema is None and m.distill_ratio == 0.0
    path = tmp_path / "model-000-0.123456.pth"
    m.save(path)
    state = torch.load(path, weights_only=False)
    assert set(state) == {"model_name", "params", "nn_state_dict"} and state["model_name"] == "MouseModel"
    m2 = argus_shim.load_model(path, device="cpu", optimizer=None, loss=None)
    for a, b in zip(m.nn_module.state_dict().values(), m2.nn_module.state_dict().values()):
        assert torch.equal(a, b)
    parts = argus_shim.deep_chunk((torch.arange(8), ([torch.arange(8), torch.arange(8)], torch.arange(8))), 2)
    assert len(parts) == 2 and parts[1][1][0][1].tolist() == [4, 5, 6, 7]


def test_ctypes_signatures_match_the_header():
    """Every ctypes signature string in sensorium_b200/_lib.py has the arity and the scalar/pointer kinds of the C
    prototype in include/dwn_b200.h (a mismatch would silently corrupt the stack of a kernel launch)."""
    import re
    from pathlib import Path
    header = (Path(__file__).resolve().parents[1] / "include" / "dwn_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(dwn_\w+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    kinds = {"p": "pointer", "i": "int", "l": "long", "f": "float", "d": "double"}
    for name, sig in _lib.SIGNATURES.items():
        assert name in protos, name
        params = [a.strip() for a in protos[name].split(",") if a.strip() and a.strip() != "void"]
        assert len(params) == len(sig), (name, len(params), len(sig))
        for ch, par in zip(sig, params):
            if "*" in par:
                want = "pointer"
            else:
                ty = par.rsplit(" ", 1)[0].replace("const", "").strip()
                want = {"int": "int", "long": "long", "float": "float", "double": "double"}.get(ty)
            assert want == kinds[ch], (name, par, ch)


def test_ema_checkpoint_layout(tmp_path):
    """EmaCheckpoint.save_model writes the EMA weights in the argus layout of the reference (ema.py:61-73) and the file
    loads back through load_model (what Predictor does, predictors.py:25)."""
    from copy import deepcopy
    from types import SimpleNamespace
    from sensorium_b200 import argus_shim
    from sensorium_b200.argus_models import MouseModel
    from sensorium_b200.ema import EmaCheckpoint
    params = {"nn_module": ("dwiseneuro", {"readout_outputs": (5, 4), "core_features": (8,), "spatial_strides": (1,),
                                           "expansion_ratio": 2, "se_reduce_ratio": 4, "cortex_features": (8,)}),
              "loss": ("mice_poisson", {}), "optimizer": ("AdamW", {"lr": 1e-3}), "device": "cpu", "amp": True,
              "iter_size": 1}
    m = MouseModel(params)
    ema_module = deepcopy(m.nn_module)
    with torch.no_grad():
        for p in ema_module.parameters():
            p.add_(1.0)  # the EMA copy differs from the raw weights
    m.model_ema = SimpleNamespace(ema=ema_module)
    path = tmp_path / "model-001-0.250000.pth"
    EmaCheckpoint().save_model(SimpleNamespace(model=m), path)
    state = torch.load(path, weights_only=False)
    assert set(state) == {"model_name", "params", "nn_state_dict"} and state["model_name"] == "MouseModel"
    assert all(v.device.type == "cpu" for v in state["nn_state_dict"].values())
    loaded = argus_shim.load_model(path, device="cpu", optimizer=None, loss=None)
    for (k, a), b in zip(ema_module.state_dict().items(), loaded.nn_module.state_dict().values()):
        assert torch.equal(a, b), k
    assert not torch.equal(loaded.nn_module.core.stem[0].weight, m.nn_module.core.stem[0].weight)


def test_synthetic_generators_match_the_oracle():
    """bench.py's GPU arm draws its inputs from sensorium_b200.synthetic (no oracle import); the tests use the oracle's
    generators: both must produce identical tensors."""
    import torch
    from oracle import dwiseneuro_oracle as O
    from sensorium_b200 import synthetic as S
    assert torch.equal(S.synthetic_clip(3, 16, 64, seed=5), O.synthetic_clip(3, 16, 64, seed=5))
    a, wa = S.synthetic_targets(5, (7, 9, 4), 16, seed=2)
    b, wb = O.synthetic_targets(5, (7, 9, 4), 16, seed=2)
    assert torch.equal(wa, wb) and all(torch.equal(x, y) for x, y in zip(a, b))
    comp, ids = S.compact_from_dense(a, wa)
    assert comp.shape == (5, 9, 16) and torch.equal(ids, wa.argmax(1))
    for i in range(5):
        m = int(ids[i])
        assert torch.equal(comp[i, :a[m].shape[1]], a[m][i]) and float(comp[i, a[m].shape[1]:].abs().sum()) == 0.0
    v, bh, pc = S.synthetic_trial(50, seed=1)
    assert v.shape == (36, 64, 50) and v.dtype.name == "uint8" and bh.shape == (2, 50) and pc.shape == (2, 50)


def test_comm_entry_points_fail_loudly_before_init():
    """The NCCL entry points behind the C ABI (include/dwn_b200.h) refuse to run without a communicator -
    no silent no-op (checked without a GPU: the call fails before anything touches the device)."""
    import ctypes
    with pytest.raises(_lib.DwnError, match="dwn_comm_init has not been called"):
        _lib.call("dwn_allreduce_bucket", ctypes.c_void_p(0), 16, 0, 1, 0, ctypes.c_void_p(0))
    with pytest.raises(_lib.DwnError, match="dwn_comm_init has not been called"):
        _lib.call("dwn_comm_group_begin")
    _lib.call("dwn_comm_destroy")   # nothing to destroy: succeeds
