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
import os
P, PS = int(os.environ.get("KB_P", 592)), int(os.environ.get("KB_PS", 148))
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
    JSE = int(os.environ.get("KB_JSE", 16))
    pp = torch.empty(B, JSE, mid, device=dev)
    timeit(f"se_pool {tag}", lambda: call("dwn_se_pool", Tm, c3, A, pp, JSE, B, T * Ho * Wo, mid, 1, st), 2 * Mo * mid * es)
    da = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
    dmean = torch.randn(B, mid, device=dev) * 0.01
    JT = int(os.environ.get("KB_JTDW", 37))
    timeit(f"tdw_bwd_reduce {tag}", lambda: call("dwn_tdw_bwd_reduce", da, Tm, c3, dmean, T * Ho * Wo, part, JT, B, mid, 1, st), 2 * Mo * mid * es)
    timeit(f"tdw_bwd {tag}", lambda: call("dwn_tdw_bwd", da, Tm, S, c3, b3, c2, wt, dmean, part, P, B, T, Ho * Wo, mid, 1, st), 4 * Mo * mid * es)
    dE = torch.empty_like(E)
    timeit(f"sdw_bwd {tag}", lambda: call("dwn_sdw_bwd", da, S, E, c2, b2, c1, ws, dE, part, PS, B * T, H, W, mid, s, 1, st), (2 * Mo + 2 * Mi) * mid * es)
    timeit(f"bn_bwd_apply {tag}", lambda: call("dwn_bn_bwd_apply", dE, E, c1, b1, Mi, mid, 1, st), 3 * Mi * mid * es)
    timeit(f"colstats {tag}", lambda: call("dwn_colstats", E, Mi, mid, mid, part, P, 1, st), Mi * mid * es)
    del E, S, Tm, A, da, dE

# ---- tensor-core GEMMs at the readout / cortex shapes (one mouse: n = 8122 -> half = 4061), timed alone
from sensorium_b200._lib import gemm  # noqa: E402


def timeg(tag, fn, flops, nbytes, iters=5):
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
    print(f"{tag:42s} {ms * 1e3:8.1f} us  {flops / ms * 1e-9:7.1f} TFLOP/s  {nbytes / ms * 1e-6:8.1f} GB/s", flush=True)


G, K, Mbt, half, n_out = 2, 4096, B * T, 4061, 8122
Kg = K // G
half_pad = ((half + 63) // 64) * 64
bf = torch.bfloat16
wq = (torch.randn(G * half, Kg, device=dev) * 0.02).to(bf)
xm = torch.randn(Mbt, K, device=dev).to(bf)
xt = torch.randn(K, Mbt, device=dev).to(bf)
pred = torch.empty(B, n_out, T, device=dev)
bias = torch.zeros(G * half, device=dev)
timeg("gemm readout_fwd", lambda: gemm(st, dtype=1, A=wq, B=xm, lda=Kg, ldb=K, a_zstride=half * Kg, b_zstride=Kg, a_zmode=1,
                                       b_zmode=1, M=half, N=Mbt, K=Kg, Z=G, epi=1, D=pred, bias=bias, beta=0.07, Tn=T,
                                       n_out_total=n_out, row_offset_per_z=half, n_limit=Mbt),
      2 * G * half * Mbt * Kg, G * half * Kg * 2 + Mbt * K * 2 + B * n_out * T * 4)
dz_nm = torch.randn(G * half, Mbt, device=dev).to(bf)
dz_mn = torch.randn(Mbt, G * half_pad, device=dev).to(bf)
dW = torch.empty(G * half, Kg, device=dev)
dxm = torch.empty(Mbt, K, device=dev)
timeg("gemm readout_wgrad", lambda: gemm(st, dtype=1, A=dz_nm, B=xt, lda=Mbt, ldb=Mbt, a_zstride=half * Mbt, b_zstride=Kg * Mbt,
                                         a_zmode=1, b_zmode=1, M=half, N=Kg, K=Mbt, Z=G, D=dW, d_dtype=0, ldd=Kg,
                                         d_zstride=half * Kg),
      2 * G * half * Mbt * Kg, G * half * Mbt * 2 + K * Mbt * 2 + G * half * Kg * 4)
timeg("gemm readout_dgrad", lambda: gemm(st, dtype=1, A=dz_mn, B=wq, b_mn=1, lda=G * half_pad, ldb=Kg, a_zstride=half_pad,
                                         b_zstride=half * Kg, a_zmode=1, b_zmode=1, M=Mbt, N=Kg, K=half, Z=G, D=dxm, d_dtype=0,
                                         ldd=K, d_zstride=Kg),
      2 * G * half * Mbt * Kg, Mbt * G * half_pad * 2 + G * half * Kg * 2 + Mbt * K * 4)
# point-wise expansion / projection GEMMs at the block shapes (HBM-bound: report GB/s next to TFLOP/s)
for tag, ci, H, W, s in shapes:
    mid, Mi = ci * 7, B * T * H * W
    Mo = Mi // (s * s)
    Xb = torch.randn(Mi, ci, device=dev).to(bf)
    Wp = (torch.randn(mid, ci, device=dev) * 0.1).to(bf)
    Eo = torch.empty(Mi, mid, device=dev, dtype=bf)
    timeg(f"gemm pw_fwd {tag}", lambda: gemm(st, dtype=1, A=Xb, B=Wp, lda=ci, ldb=ci, M=Mi, N=mid, K=ci, Z=1, D=Eo, d_dtype=1,
                                             ldd=mid), 2 * Mi * mid * ci, (Mi * ci + mid * ci + Mi * mid) * 2)
    dXo = torch.empty(Mi, ci, device=dev)
    timeg(f"gemm pw_dgrad {tag}", lambda: gemm(st, dtype=1, A=Eo, B=Wp, b_mn=1, lda=mid, ldb=ci, M=Mi, N=ci, K=mid, Z=1, D=dXo,
                                               d_dtype=0, ldd=ci), 2 * Mi * mid * ci, Mi * mid * 2 + mid * ci * 2 + Mi * ci * 4)
    Nsp = Mo // B
    Aa = torch.randn(Mo, mid, device=dev).to(bf)
    Wb = (torch.randn(B, ci, mid, device=dev) * 0.1).to(bf)
    Yo = torch.empty(Mo, ci, device=dev, dtype=bf)
    timeg(f"gemm pwl_fwd {tag}", lambda: gemm(st, dtype=1, A=Aa, B=Wb, lda=mid, ldb=mid, a_zstride=Nsp * mid, b_zstride=ci * mid,
                                              a_zmode=1, b_zmode=1, M=Nsp, N=ci, K=mid, Z=B, D=Yo, d_dtype=1, ldd=ci,
                                              d_zstride=Nsp * ci), 2 * Mo * mid * ci, (Mo * mid + B * ci * mid + Mo * ci) * 2)
    del Xb, Eo, dXo, Aa, Yo
for ci_, co_ in ((256, 1024), (1024, 2048), (2048, 4096)):
    xa = torch.randn(Mbt, ci_, device=dev).to(bf)
    wc = (torch.randn(co_, ci_ // G, device=dev) * 0.05).to(bf)
    yo = torch.empty(Mbt, co_, device=dev).to(bf)
    timeg(f"gemm cortex_fwd {ci_}->{co_}", lambda: gemm(st, dtype=1, A=xa, B=wc, lda=ci_, ldb=ci_ // G, a_zstride=ci_ // G,
                                                         b_zstride=(co_ // G) * (ci_ // G), a_zmode=1, b_zmode=1, M=Mbt,
                                                         N=co_ // G, K=ci_ // G, Z=G, D=yo, d_dtype=1, ldd=co_,
                                                         d_zstride=co_ // G),
          2 * Mbt * co_ * (ci_ // G), (Mbt * ci_ + co_ * ci_ // G + Mbt * co_) * 2)
print("done")
