"""GPU check (not a pytest): the TMA-staged spatial depth-wise forward (sdw_fwd_v6) against the cp.async kernels
(DWN_SDW_FWD_TMA=0) on the six C2 block shapes and two small ones - same arithmetic in the same order, so S_raw must be
bit-identical and the partial sums equal after the sum over workers - followed by an A/B timing at batch 32.
Usage: python tests/gpu_checks/check_sdw_fwd_tma.py [--time]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from sensorium_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
shapes = [("blk0", 448, 64, 64, 2), ("blk1", 448, 32, 32, 1), ("blk4", 896, 32, 32, 2), ("blk5", 896, 16, 16, 1),
          ("blk7", 1792, 16, 16, 2), ("blk8", 1792, 8, 8, 1), ("tiny0", 64, 32, 32, 2), ("tiny1", 64, 16, 16, 1)]


def coef(C):
    c = torch.empty(4, C, device=dev)
    c[0].uniform_(0.5, 1.5); c[1].uniform_(-0.3, 0.3); c[2].uniform_(-0.3, 0.3); c[3].uniform_(0.5, 1.5)
    return c


def run(mode, tho, args):
    os.environ["DWN_SDW_FWD_TMA"] = str(mode)
    os.environ["DWN_SDW_FWD_THO"] = str(tho)
    (E, c1, ws, NP, H, W, mid, s, PS) = args
    S = torch.full((NP * (H // s) * (W // s), mid), float("nan"), device=dev).to(torch.bfloat16)
    part = torch.full((PS, 2, mid), float("nan"), device=dev)
    call("dwn_sdw_fwd", E, c1, ws, S, part, PS, NP, H, W, mid, s, 1, st)
    torch.cuda.synchronize()
    return S, part


ok = True
torch.manual_seed(0)
SMALL = "--small" in sys.argv  # compute-sanitizer runs: every kernel instantiation once, few planes
for tag, mid, H, W, s in shapes:
    for NP, PS in (((5, 3),) if SMALL else ((24, 37), (64, 42), (100, 7))):
        E = torch.randn(NP * H * W, mid, device=dev).to(torch.bfloat16)
        c1 = coef(mid)
        ws = torch.randn(mid, 9, device=dev) * 0.3
        args = (E, c1, ws, NP, H, W, mid, s, PS)
        ref_S, ref_part = run(0, 0, args)
        for tho in ((0, 8, 4, 16) if s == 1 else (0, 2, 4)):
            S, part = run(1, tho, args)
            same = torch.equal(S.view(torch.int16), ref_S.view(torch.int16))
            ps, rs = part.double().sum(0), ref_part.double().sum(0)
            perr = ((ps - rs).abs().max() / rs.abs().max()).item()
            print(f"{tag} NP={NP} P={PS} tho={tho}: S bit-identical={same} sum-of-partials rel={perr:.3e}", flush=True)
            ok = ok and same and perr < 2e-6
print("CHECK", "PASS" if ok else "FAIL")

if "--time" in sys.argv:
    B, T = 32, 16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for tag, mid, H, W, s in shapes[:6]:
        NP = B * T
        E = torch.randn(NP * H * W, mid, device=dev).to(torch.bfloat16)
        S = torch.empty(NP * (H // s) * (W // s), mid, device=dev, dtype=torch.bfloat16)
        c1 = coef(mid)
        ws = torch.randn(mid, 9, device=dev) * 0.3
        nbytes = (E.numel() + S.numel()) * 2
        for PS in (42, 148):
            part = torch.empty(PS, 2, mid, device=dev)
            for mode, tho in ((0, 0),) + (((1, 8), (1, 16), (1, 4)) if s == 1 else ((1, 4), (1, 2))):
                os.environ["DWN_SDW_FWD_TMA"] = str(mode)
                os.environ["DWN_SDW_FWD_THO"] = str(tho)
                fn = lambda: call("dwn_sdw_fwd", E, c1, ws, S, part, PS, NP, H, W, mid, s, 1, st)
                fn()
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = sorted(ts)[2]
                print(f"sdw_fwd {tag} P={PS} tma={mode} tho={tho}: {ms:.3f} ms  {nbytes / ms * 1e-6:7.1f} GB/s  {nbytes / ms * 1e-6 / 6539.5 * 100:5.1f}% of HBM peak", flush=True)
        del E, S
