"""Per-kernel table (serialized CUDA-event intervals) of the eval forward, batch N windows, one readout.
    python tests/gpu_checks/prof_eval.py [fp32|bf16] [batch]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from sensorium_b200 import DwiseNeuro, _lib, constants, engine  # noqa: E402
from sensorium_b200.synthetic import synthetic_clip  # noqa: E402
from sensorium_b200.utils import init_weights  # noqa: E402
from tests.shapes import TRUE_BATCH_KW  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW).to(dev)
    init_weights(net)
    net.eval()
    net.precision = mode
    x = synthetic_clip(B, 16, 64, seed=0).to(dev)
    with torch.no_grad():
        for _ in range(2):
            net(x, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            net(x, 0)
        e1.record()
        torch.cuda.synchronize()
        print(f"eval fwd {mode} batch {B}: {e0.elapsed_time(e1) / 5:.2f} ms")
        engine.SERIALIZE = True
        _lib.PROF = []
        for _ in range(3):
            net(x, 0)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0, 0])
    for name, tag, nbytes, flops, a, b in _lib.PROF:
        k = tag or name
        agg[k][0] += a.elapsed_time(b) / 3
        agg[k][1] += 1
        agg[k][2] += nbytes / 3
    tot = sum(v[0] for v in agg.values())
    print(f"serialized sum {tot:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda r: -r[1][0]):
        print(f"{k:24s} {v[0]:8.3f} ms  {v[1] // 3:4d} launches  {v[2] / max(v[0], 1e-9) * 1e-6:8.1f} GB/s")


if __name__ == "__main__":
    main()
