"""GPU diagnostic (not a pytest): dwn_gemm (tcgen05 bf16 + SIMT fp32) against torch.matmul for every operand
major / batching / epilogue the engine uses.  Prints one line per case; exit code 1 on any failure."""
import itertools
import math
import sys

import torch

sys.path.insert(0, ".")
from sensorium_b200 import _lib  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
fails = 0


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def make_operand(rows, K, mn, Z, dtype, pad=0):
    """returns (tensor as stored, logical [Z][rows][K] fp32 view, ld, zstride)"""
    pad = (-(rows if mn else K)) % 8
    if not mn:
        t = torch.randn(Z, rows, K + pad, device=dev).to(dtype)
        return t, t[:, :, :K].float(), K + pad, rows * (K + pad)
    t = torch.randn(Z, K, rows + pad, device=dev).to(dtype)
    return t, t[:, :, :rows].float().transpose(1, 2), rows + pad, K * (rows + pad)


def run_case(name, M, N, K, Z, a_mn, b_mn, dt, d_dt, tol, block_n=0, dbg=None):
    global fails
    dtype = torch.bfloat16 if dt else torch.float32
    A, Al, lda, azs = make_operand(M, K, a_mn, Z, dtype)
    Bm, Bl, ldb, bzs = make_operand(N, K, b_mn, Z, dtype)
    ref = torch.matmul(Al.double(), Bl.double().transpose(1, 2)).float()
    D = torch.full((Z, M, N), float("nan"), device=dev, dtype=torch.bfloat16 if d_dt else torch.float32)
    kw = dict(dtype=dt, A=A, B=Bm, a_mn=a_mn, b_mn=b_mn, lda=lda, ldb=ldb, a_zstride=azs, b_zstride=bzs, a_zmode=1,
              b_zmode=1, M=M, N=N, K=K, Z=Z, D=D, d_dtype=d_dt, ldd=N, d_zstride=M * N, block_n=block_n)
    if dbg:
        kw.update(dbg)
    try:
        _lib.gemm(st, **kw)
        torch.cuda.synchronize()
        e = rel(D.float(), ref)
        nan = int(torch.isnan(D.float()).sum())
    except Exception as ex:  # noqa: BLE001
        print(f"FAIL {name}: exception {ex}")
        fails += 1
        return False
    ok = e < tol and nan == 0
    print(f"{'ok  ' if ok else 'FAIL'} {name}: M={M} N={N} K={K} Z={Z} a_mn={a_mn} b_mn={b_mn} dt={dt} d_dt={d_dt} "
          f"bn={block_n} rel={e:.3e} nan={nan} {dbg or ''}")
    if not ok:
        fails += 1
    return ok


for dt, tol in ((0, 1e-5), (1, 2e-2)):
    for a_mn, b_mn in itertools.product((0, 1), (0, 1)):
        run_case("basic", 256, 128, 128, 1, a_mn, b_mn, dt, 0, tol)
        run_case("ragged", 300, 192, 200, 2, a_mn, b_mn, dt, 0, tol)
    run_case("pw-like", 4096, 448, 64, 1, 0, 0, dt, dt, tol if not dt else 2e-2)
    run_case("pwl-like", 1024, 64, 448, 3, 0, 0, dt, dt, tol if not dt else 2e-2)
    run_case("dgrad-like", 2048, 448, 64, 2, 0, 1, dt, dt, tol if not dt else 2e-2)
    run_case("wgrad-like", 448, 64, 8192, 2, 1, 1, dt, 0, tol if not dt else 2e-2)
    run_case("small-n", 512, 16, 256, 1, 0, 0, dt, 0, tol)
    run_case("many-tiles", 128 * 300, 256, 64, 1, 0, 0, dt, dt, tol if not dt else 2e-2)

if fails:
    print("---- MN-major descriptor sweep (debug) ----")
    for lbo, sbo in ((8192, 1024), (1024, 8192), (128, 1024), (1024, 128), (2048, 1024), (1024, 2048)):
        run_case("sweepA", 256, 128, 128, 1, 1, 0, 1, 0, 2e-2, dbg=dict(dbg_lbo_a=lbo, dbg_sbo_a=sbo))
        run_case("sweepB", 256, 128, 128, 1, 0, 1, 1, 0, 2e-2, dbg=dict(dbg_lbo_b=lbo, dbg_sbo_b=sbo))

# ---- readout epilogue (bias + softplus, transposed store) ----
for dt, tol in ((0, 1e-5), (1, 2e-2)):
    dtype = torch.bfloat16 if dt else torch.float32
    Bsz, T, K, G, n_out = 4, 16, 256, 2, 301
    half = math.ceil(n_out / G)
    Kg = K // G
    W = (torch.randn(2 * half, Kg, device=dev) * 0.5).to(dtype)
    bias = torch.randn(2 * half, device=dev)
    X = torch.randn(Bsz * T, K, device=dev).to(dtype)
    pred = torch.full((Bsz, n_out, T), float("nan"), device=dev)
    _lib.gemm(st, dtype=dt, A=W, B=X, lda=Kg, ldb=K, a_zstride=half * Kg, b_zstride=Kg, a_zmode=1, b_zmode=1, M=half,
              N=Bsz * T, K=Kg, Z=G, epi=1, D=pred, bias=bias, beta=0.07, Tn=T, n_out_total=n_out, row_offset_per_z=half,
              n_limit=Bsz * T)
    torch.cuda.synchronize()
    xin = X.float().view(Bsz, T, K).permute(0, 2, 1)
    ref = torch.nn.functional.softplus(
        torch.nn.functional.conv1d(xin, W.float()[:, :, None], bias, groups=G)[:, :n_out], beta=0.07)
    e = rel(pred, ref)
    ok = e < tol and not torch.isnan(pred).any()
    print(f"{'ok  ' if ok else 'FAIL'} readout dt={dt} rel={e:.3e}")
    fails += 0 if ok else 1

print("GEMM CHECK", "FAILED" if fails else "PASSED", fails)
sys.exit(1 if fails else 0)
