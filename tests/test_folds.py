"""Fold-parallel driver: host logic on CPU — checkpoint selection (utils.py:22-43), the per-iteration
learning-rate schedule of both stages against torch's own schedulers (scripts/train.py:122-136), fold sharding, and the
2-rank gloo run of the fold loop."""
import math
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_get_best_model_path(tmp_path):
    from sensorium_b200.folds import get_best_model_path
    assert get_best_model_path(tmp_path) is None
    assert get_best_model_path(tmp_path, return_score=True) == (None, -math.inf)
    for name in ("model-003-0.251000.pth", "model-017-0.290500.pth", "model-009-0.270000.pth", "notes.txt", "raw.pth"):
        (tmp_path / name).write_bytes(b"x")
    assert get_best_model_path(tmp_path).name == "model-017-0.290500.pth"
    p, s = get_best_model_path(tmp_path, return_score=True, more_better=False)
    assert p.name == "model-003-0.251000.pth" and abs(s - 0.251) < 1e-9


def test_stage_lr_matches_torch_schedulers():
    from sensorium_b200.folds import stage_lr
    base, eta_min, n = 2.4e-3, 2.4e-5, 37
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda x: x / n)          # warm-up stage, stepped per iteration
    for it in range(n):
        assert abs(opt.param_groups[0]["lr"] - stage_lr("warmup", base, eta_min, it, n)) < 1e-12
        opt.step()
        sch.step()
    opt = torch.optim.SGD([p], lr=base)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=n, eta_min=eta_min)
    for it in range(n):
        assert abs(opt.param_groups[0]["lr"] - stage_lr("train", base, eta_min, it, n)) < 1e-9
        opt.step()
        sch.step()


def test_shard_folds():
    from sensorium_b200.folds import folds_splits, shard_folds
    assert folds_splits == [f"fold_{i}" for i in range(7)]
    got = [shard_folds(folds_splits, r, 8) for r in range(8)]
    assert got[:7] == [[f] for f in folds_splits] and got[7] == []
    assert shard_folds(folds_splits, 1, 3) == ["fold_1", "fold_4"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pathlib import Path
    from sensorium_b200 import folds
    seen = []

    def fake_train_fold(config, save_dir, train_loader, val_loader, log=None, distill_model_path=None):
        save_dir = Path(save_dir)
        save_dir.mkdir(parents=True, exist_ok=True)
        seen.append((save_dir.name, train_loader, val_loader))
        path = save_dir / f"model-002-0.{int(save_dir.name[-1]) + 1}00000.pth"
        path.write_bytes(b"x")
        return path

    folds.train_fold = fake_train_fold
    cfg = {"argus_params": {"device": "cpu"}}
    res = folds.run_folds(cfg, Path(tmp) / "exp", lambda tr, va: (tuple(tr), tuple(va)), folds="0,1,2", log=None)
    ok = list(res) == ["fold_0", "fold_1", "fold_2"] and all(p.exists() for p in res.values())
    ok &= [s[0] for s in seen] == (["fold_0", "fold_2"] if rank == 0 else ["fold_1"])
    # the validation fold is held out of the training splits (scripts/train.py:180-181)
    ok &= all(s[2] == (s[0],) and s[0] not in s[1] and len(s[1]) == 6 for s in seen)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_fold_loop_sharded_over_two_ranks(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
