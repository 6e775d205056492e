"""GPU check (torchrun, >= 2 GPUs; also wrapped as a -m gpu test): the NCCL entry points behind the C ABI
(dwn_comm_unique_id / dwn_comm_init / dwn_allreduce_bucket / dwn_comm_group_* / dwn_comm_destroy, include/dwn_b200.h).
torch.distributed (gloo, CPU) is used ONLY to hand rank 0's rendezvous token to the other ranks - the exchange itself goes
through the library.  Checks: fp32 mean (the DDP gradient average), bf16 sum, int32 max (has-grad flags), several buckets
in one group, against closed-form expectations and against torch.distributed's own NCCL all-reduce.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_checks/check_comm_cabi.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from sensorium_b200._lib import call  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("gloo")
token = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    call("dwn_comm_unique_id", token)
dist.broadcast(token, src=0)
call("dwn_comm_init", rank, world, token)
comm = torch.cuda.Stream(device=dev)
ok = True
n = 3_000_001
g = torch.Generator().manual_seed(5)
base = torch.randn(n, generator=g)
with torch.cuda.stream(comm):
    st = comm.cuda_stream
    x = (base * (rank + 1)).to(dev)                      # fp32 mean over ranks = base * (world + 1) / 2
    call("dwn_allreduce_bucket", x, n, 0, 1, 0, st)
    y = torch.full((1025,), float(rank + 1), device=dev, dtype=torch.bfloat16)   # bf16 sum = world (world + 1) / 2
    flags = torch.tensor([1 if rank == 0 else 0, 0, 1 if rank == world - 1 else 0, 0], dtype=torch.int32, device=dev)
    call("dwn_comm_group_begin")
    call("dwn_allreduce_bucket", y, y.numel(), 1, 0, 0, st)
    call("dwn_allreduce_bucket", flags, flags.numel(), 2, 0, 1, st)
    call("dwn_comm_group_end")
comm.synchronize()
want = base.double() * (world + 1) / 2
err = float((x.cpu().double() - want).abs().max() / want.abs().max())
ok = ok and err < 1e-6
ok = ok and bool((y.float() == world * (world + 1) / 2).all())
ok = ok and flags.tolist() == [1, 0, 1, 0]
print(f"[rank {rank}] fp32 mean rel err {err:.2e}  bf16 sum {float(y[0])}  flags {flags.tolist()}  {'ok' if ok else 'FAIL'}", flush=True)
call("dwn_comm_destroy")
res = torch.tensor([1 if ok else 0])
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("check_comm_cabi:", "PASS" if int(res) == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(res) == 1 else 1)
