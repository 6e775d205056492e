"""GPU micro-benchmark (not a pytest): FusedAdamW step of the true_batch_001 parameter set (170.7 M parameters, 30 B each),
L2 flushed between iterations.  Usage: python tests/gpu_checks/bench_adamw.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
from sensorium_b200 import DwiseNeuro  # noqa: E402
from sensorium_b200.constants import mice, num_neurons  # noqa: E402
from sensorium_b200.optim import FusedAdamW  # noqa: E402
from tests.shapes import TRUE_BATCH_KW  # noqa: E402

dev = torch.device("cuda:0")
net = DwiseNeuro(readout_outputs=tuple(num_neurons), **TRUE_BATCH_KW).to(dev)
opt = FusedAdamW(net.parameters(), lr=2.4e-3, weight_decay=0.05)
for p in net.parameters():
    p.grad = torch.randn_like(p) * 1e-3
n = sum(p.numel() for p in net.parameters())
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
opt.step()
ts = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); opt.step(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"adamw: {n / 1e6:.1f} M params  {ms:.3f} ms  {n * 30 / ms * 1e-6:.0f} GB/s  "
      f"{n * 30 / ms * 1e-6 / 6539.5 * 100:.1f}% of HBM peak", flush=True)
