"""2+ GPU check of the data-parallel gradient exchange through the real backward (torchrun, NCCL).

Every rank trains on its own shard; the set of mice present differs per rank (rank 1 lacks mouse 1, mouse 2 is absent
everywhere).  The exchanged gradients must equal the mean over ranks of the single-GPU gradients of every shard, and
the optimizer's active flags must skip only the mouse that is absent everywhere.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_checks/check_dp.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import dwiseneuro_oracle as O  # noqa: E402
from tests.shapes import TINY_KW, TINY_OUTS  # noqa: E402


def shard(r, dev):
    x = O.synthetic_clip(3, 8, 16, seed=50 + r).to(dev)
    tg, w = O.synthetic_targets(3, TINY_OUTS, 8, seed=70 + r)
    w[:, 2] = 0
    w[:, 0] = 1.0
    w[:, 1] = 0.0 if r % 2 == 1 else 1.0
    return x, [t.to(dev) for t in tg], w.to(dev)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.losses import MicePoissonLoss
    from sensorium_b200.parallel import DataParallelGrads
    from sensorium_b200.utils import init_weights
    kw = dict(TINY_KW, drop_path_rate=0.0, drop_rate=0.0)
    ok = True
    for mode in ("fp32", "bf16"):
        torch.manual_seed(0)
        net = DwiseNeuro(readout_outputs=TINY_OUTS, **kw).to(dev)
        init_weights(net)
        net.train()
        net.precision = mode
        loss_fn = MicePoissonLoss()
        params = list(net.parameters())
        mean = [torch.zeros_like(p) for p in params]
        for r in range(world):  # single-GPU gradients of every shard
            x, tg, w = shard(r, dev)
            for p in params:
                p.grad = None
            loss_fn(net(x), (tg, w)).backward()
            for m, p in zip(mean, params):
                if p.grad is not None:
                    m += p.grad / world
        dp = DataParallelGrads.attach(net, False)
        for it in range(3):
            x, tg, w = shard(rank, dev)
            for p in params:
                p.grad = None
            loss_fn(net(x), (tg, w)).backward()
        torch.cuda.synchronize()
        gmax = max(float(m.abs().max()) for m in mean)
        worst = 0.0
        for (n, p), m in zip(net.named_parameters(), mean):
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            worst = max(worst, float((g - m).abs().max()) / gmax)
        names = [n for n, _ in net.named_parameters()]
        act = dp.active.tolist()
        flags_ok = all((a == 0) == n.startswith("readouts.2.") for a, n in zip(act, names))
        tol = 1e-5 if mode == "fp32" else 2e-2  # bf16: batch statistics are per-shard either way, only rounding noise
        good = worst < tol and flags_ok
        ok &= good
        print(f"[rank {rank}] {mode}: max |dp grad - mean of shard grads| / gmax = {worst:.2e}  flags_ok={flags_ok} "
              f"bytes={dp.bytes_reduced}  {'ok' if good else 'FAIL'}", flush=True)
        net._dp = None
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("check_dp:", "PASS" if int(t) else "FAIL", flush=True)
    sys.exit(0 if int(t) else 1)


if __name__ == "__main__":
    main()
