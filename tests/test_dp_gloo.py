"""CPU, world_size 2 over gloo: the bucketed gradient exchange of sensorium_b200.parallel (host logic of the N>1 path):
big tensors reduced individually, small ones coalesced into one flat bucket, result == mean over ranks, and the
per-mouse has-grad flags are MAX-reduced so a mouse absent everywhere stays inactive."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.parallel import DataParallelGrads
    torch.manual_seed(rank)  # different weights per rank before attach
    net = DwiseNeuro(readout_outputs=(5, 4, 3), core_features=(8,), spatial_strides=(1,), expansion_ratio=2,
                     se_reduce_ratio=4, cortex_features=(8,), groups=2)
    dp = DataParallelGrads.attach(net, False)
    w0 = net.core.stem[0].weight.detach().clone()
    params = list(net.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    grads = {p: torch.randn(p.shape, generator=g) for p in params}
    big = torch.randn(1 << 20, generator=g)
    grads["big"] = big.clone()
    local = {k: v.clone() for k, v in grads.items()}
    live = [rank == 0, False, True]  # mouse 0 only on rank 0, mouse 1 nowhere, mouse 2 everywhere
    dp.begin(live, torch.device("cpu"))
    dp.reduce(grads, list(grads.keys()))
    dp.finish(torch.device("cpu"))
    # reference: gather every rank's local grads
    ok = True
    for k in local:
        bucket = [torch.zeros_like(local[k]) for _ in range(world)]
        dist.all_gather(bucket, local[k])
        mean = sum(bucket) / world
        ok &= torch.allclose(grads[k], mean, atol=1e-6)
    idx = {id(p): i for i, p in enumerate(params)}
    act = dp.active.tolist()
    m0 = idx[id(net.readouts[0].layer[1].weight)]
    m1 = idx[id(net.readouts[1].layer[1].weight)]
    m2 = idx[id(net.readouts[2].layer[1].bias)]
    ok &= act[m0] == 1 and act[m1] == 0 and act[m2] == 1 and act[idx[id(net.core.stem[0].weight)]] == 1
    q.put((rank, bool(ok), w0))
    dist.destroy_process_group()


def test_bucketed_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert torch.equal(res[0][2], res[1][2])  # attach() broadcast rank 0's weights


def _worker_readouts(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.parallel import DataParallelGrads
    net = DwiseNeuro(readout_outputs=(5, 4, 3), core_features=(8,), spatial_strides=(1,), expansion_ratio=2,
                     se_reduce_ratio=4, cortex_features=(8,), groups=2)
    dp = DataParallelGrads.attach(net, False)
    # the set of mice with a local sample differs per rank (the usual case at N=8 with 10 mice in a batch of 32):
    # rank 0 sees mice {0, 2}, rank 1 sees {1, 2}; the readout shapes differ, so any rank-dependent collective order
    # would pair tensors of different sizes
    live = [rank == 0, rank == 1, True]
    g = torch.Generator().manual_seed(7 + rank)
    grads, local = {}, {}
    for m in reversed(range(3)):  # insertion order must not matter either
        for p in net.readouts[m].parameters():
            local[p] = torch.randn(p.shape, generator=g) if live[m] else torch.zeros_like(p)
            if live[m]:
                grads[p] = local[p].clone()
    dp.begin(live, torch.device("cpu"))
    dp.reduce_readouts(grads)
    dp.finish(torch.device("cpu"))
    ok = True
    for p, v in local.items():
        bucket = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(bucket, v)
        ok &= torch.allclose(grads[p], sum(bucket) / world, atol=1e-6)
    params = list(net.parameters())
    idx = {id(p): i for i, p in enumerate(params)}
    act = dp.active.tolist()
    ok &= all(act[idx[id(p)]] == 1 for m in range(3) for p in net.readouts[m].parameters())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_readout_exchange_with_rank_dependent_mice():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_readouts, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def _worker_trials(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sensorium_b200.parallel import predict_trials_sharded, shard_trials
    trials = [torch.full((3,), float(i)) for i in range(7)]  # 7 trials do not divide 2 ranks
    seen = []

    def predict(t):
        seen.append(int(t[0]))
        return t * 2 + rank * 0  # result must not depend on which rank computed it

    res = predict_trials_sharded(predict, trials)
    ok = sorted(res) == list(range(7)) and all(torch.equal(res[i], trials[i] * 2) for i in range(7))
    ok &= seen == shard_trials(7, rank, world) == list(range(rank, 7, world))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_inference_trials_are_sharded_and_gathered():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_trials, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
    from sensorium_b200.parallel import shard_trials
    assert shard_trials(5, 0, 1) == [0, 1, 2, 3, 4] and shard_trials(3, 5, 8) == []


def _worker_flags(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sensorium_b200 import DwiseNeuro
    from sensorium_b200.parallel import DataParallelGrads
    net = DwiseNeuro(readout_outputs=(5, 4, 3), core_features=(8,), spatial_strides=(1,), expansion_ratio=2,
                     se_reduce_ratio=4, cortex_features=(8,), groups=2)

    class Opt:  # stand-in for FusedAdamW: attach() must wire the provider, step() calls consumed()
        active_provider = None

    opt = Opt()
    dp = DataParallelGrads.attach(net, opt)
    ok = opt.active_provider is dp
    params = list(net.parameters())
    idx = {id(p): i for i, p in enumerate(params)}
    w = [idx[id(net.readouts[m].layer[1].weight)] for m in range(3)]
    cpu = torch.device("cpu")
    # iter_size 2: micro-batch 1 has mouse 0 (rank 0 only), micro-batch 2 has mouse 2 only; mouse 1 never
    dp.begin([rank == 0, False, False], cpu)
    dp.finish(cpu)
    dp.begin([False, False, True], cpu)
    dp.finish(cpu)
    act = dp.active.tolist()
    ok &= [act[i] for i in w] == [1, 0, 1]
    dp.consumed()                                   # optimizer stepped: a new accumulation window starts
    ok &= dp.active is None
    # forward(x, index=1): one output whose mouse is 1 -> the flags still cover all three mice
    dp.begin([True], cpu, mice=[1])
    dp.finish(cpu)
    act = dp.active.tolist()
    ok &= [act[i] for i in w] == [0, 1, 0] and act[idx[id(net.core.stem[0].weight)]] == 1
    try:
        DataParallelGrads.attach(net, None)
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_has_grad_flags_accumulate_over_micro_batches_and_cover_all_mice():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_flags, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
