"""GPU micro-benchmark (not a pytest) of the bandwidth-bound kernels at the C2 (batch 32, expansion 7) shapes.
CUDA-event timing, L2 flushed between iterations.  Usage: python tests/gpu_checks/kbench.py [names...] [--ncu]"""
import sys

import torch

sys.path.insert(0, ".")
from sensorium_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
B, T = 32, 16
HBM = 6539.5
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
NCU = "--ncu" in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith("--")]


def timeit(tag, fn, nbytes, iters=5):
    if names and not any(n in tag for n in names):
        return
    if NCU:
        fn()
        torch.cuda.synchronize()
        return
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{tag:42s} {ms:8.3f} ms  {nbytes / ms * 1e-6:8.1f} GB/s  {nbytes / ms * 1e-6 / HBM * 100:5.1f}% of HBM peak", flush=True)


def coef(C):
    c = torch.empty(4, C, device=dev)
    c[0].uniform_(0.5, 1.5); c[1].uniform_(-0.3, 0.3); c[2].uniform_(-0.3, 0.3); c[3].uniform_(0.5, 1.5)
    return c


def bcoef(C):
    return torch.randn(2, C, device=dev) * 0.01


shapes = [  # (tag, ci, H, W, stride)
    ("blk0", 64, 64, 64, 2), ("blk1", 64, 32, 32, 1), ("blk4", 128, 32, 32, 2), ("blk5", 128, 16, 16, 1),
    ("blk7", 256, 16, 16, 2), ("blk8", 256, 8, 8, 1)]
P, PS = 592, 148
for tag, ci, H, W, s in shapes:
    mid = ci * 7
    Ho, Wo = H // s, W // s
    Mi, Mo = B * T * H * W, B * T * Ho * Wo
    E = torch.randn(Mi, mid, device=dev).to(torch.bfloat16)
    S = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
    Tm = torch.empty_like(S)
    A = torch.empty_like(S)
    c1, c2, c3 = coef(mid), coef(mid), coef(mid)
    b1, b2, b3 = bcoef(mid), bcoef(mid), bcoef(mid)
    ws = torch.randn(mid, 9, device=dev) * 0.3
    wt = torch.randn(mid, 5, device=dev) * 0.4
    part = torch.empty(P, 11, mid, device=dev)
    es = 2
    timeit(f"sdw_fwd {tag}", lambda: call("dwn_sdw_fwd", E, c1, ws, S, part, PS, B * T, H, W, mid, s, 1, st), (Mi + Mo) * mid * es)
    timeit(f"tdw_fwd {tag}", lambda: call("dwn_tdw_fwd", S, c2, wt, Tm, part, P, B, T, Ho * Wo, mid, 1, st), 2 * Mo * mid * es)
    pp = torch.empty(B, 16, mid, device=dev)
    timeit(f"se_pool {tag}", lambda: call("dwn_se_pool", Tm, c3, A, pp, 16, B, T * Ho * Wo, mid, 1, st), 2 * Mo * mid * es)
    da = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
    dmean = torch.randn(B, mid, device=dev) * 0.01
    timeit(f"tdw_bwd_reduce {tag}", lambda: call("dwn_tdw_bwd_reduce", da, Tm, c3, dmean, T * Ho * Wo, part, 37, B, mid, 1, st), 3 * Mo * mid * es)
    timeit(f"tdw_bwd {tag}", lambda: call("dwn_tdw_bwd", da, Tm, S, c3, b3, c2, wt, part, P, B, T, Ho * Wo, mid, 1, st), 4 * Mo * mid * es)
    dE = torch.empty_like(E)
    timeit(f"sdw_bwd {tag}", lambda: call("dwn_sdw_bwd", da, S, E, c2, b2, c1, ws, dE, part, PS, B * T, H, W, mid, s, 1, st), (2 * Mo + 2 * Mi) * mid * es)
    timeit(f"bn_bwd_apply {tag}", lambda: call("dwn_bn_bwd_apply", dE, E, c1, b1, Mi, mid, 1, st), 3 * Mi * mid * es)
    timeit(f"colstats {tag}", lambda: call("dwn_colstats", E, Mi, mid, mid, part, P, 1, st), Mi * mid * es)
    del E, S, Tm, A, da, dE
print("done")
