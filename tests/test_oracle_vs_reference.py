"""CPU, build container only: the oracle against the LIVE reference modules (skipped where /root/reference is absent,
e.g. on the GPU box — the committed golden fixtures cover that case)."""
import importlib.util
from pathlib import Path

import pytest
import torch

from oracle import dwiseneuro_oracle as O

REF = Path("/root/reference/src")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference sources not present")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("kw,outs,B,T,HW", [
    (dict(core_features=(8, 8, 16), spatial_strides=(2, 1, 2), expansion_ratio=2, se_reduce_ratio=4,
          cortex_features=(16, 32), groups=2), (7, 6), 3, 8, 16),
    (dict(core_features=(8, 16), spatial_strides=(1, 2), expansion_ratio=3, se_reduce_ratio=2,
          cortex_features=(24,), groups=2, drop_path_rate=0.5, drop_rate=0.2), (5, 9, 4), 2, 6, 12),
])
def test_live_reference_bit_exact(kw, outs, B, T, HW):
    R = _load(REF / "models" / "dwiseneuro.py", "ref_dw")
    L = _load(REF / "losses.py", "ref_loss")
    torch.manual_seed(0)
    net = R.DwiseNeuro(readout_outputs=outs, **kw)
    for p in net.parameters():
        torch.nn.init.normal_(p, 0, 0.3)
    cfg = O.make_cfg(outs, **kw)
    x = O.synthetic_clip(B, T, HW, seed=0)
    tg, w = O.synthetic_targets(B, outs, T, seed=3)
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        for a, b in zip(net(x), O.dwiseneuro_forward(x, sd, cfg, None, False)):
            assert torch.equal(a, b)
        assert torch.equal(net(x, 1), O.dwiseneuro_forward(x, sd, cfg, 1, False))
    net.train()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    torch.manual_seed(5)
    a = net(x)
    torch.manual_seed(5)
    b = O.dwiseneuro_forward(x, sd, cfg, None, True)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    for k, v in net.state_dict().items():
        assert torch.equal(sd[k], v), k
    assert torch.equal(L.MicePoissonLoss()(a, (tg, w)), O.mice_poisson_loss(b, tg, w))
