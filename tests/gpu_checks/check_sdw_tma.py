"""GPU check (not a pytest): the TMA-staged spatial depth-wise backward (sdw_bwd_v6) against the cp.async kernels
(DWN_SDW_TMA=0) on the six C2 block shapes - same arithmetic in the same order, so dE and the partial sums must be
bit-identical - followed by an A/B timing at batch 32.  Usage: python tests/gpu_checks/check_sdw_tma.py [--time]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from sensorium_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
shapes = [("blk0", 64, 64, 64, 2), ("blk1", 64, 32, 32, 1), ("blk4", 128, 32, 32, 2), ("blk5", 128, 16, 16, 1),
          ("blk7", 256, 16, 16, 2), ("blk8", 256, 8, 8, 1)]


def coef(C):
    c = torch.empty(4, C, device=dev)
    c[0].uniform_(0.5, 1.5); c[1].uniform_(-0.3, 0.3); c[2].uniform_(-0.3, 0.3); c[3].uniform_(0.5, 1.5)
    return c


def run(mode, thi, args):
    os.environ["DWN_SDW_TMA"] = str(mode)
    os.environ["DWN_SDW_THI"] = str(thi)
    (da, S, E, c2, b2, c1, ws, NP, H, W, mid, s, PS) = args
    dE = torch.full_like(E, float("nan"))
    part = torch.full((PS, 11, mid), float("nan"), device=dev)
    call("dwn_sdw_bwd", da, S, E, c2, b2, c1, ws, dE, part, PS, NP, H, W, mid, s, 1, st)
    torch.cuda.synchronize()
    return dE, part


ok = True
torch.manual_seed(0)
SMALL = "--small" in sys.argv  # compute-sanitizer runs: every kernel instantiation once, few planes
for tag, ci, H, W, s in shapes:
    for NP, PS in (((5, 3),) if SMALL else ((24, 37), (64, 148))):
        mid = ci * 7
        Ho, Wo = H // s, W // s
        E = torch.randn(NP * H * W, mid, device=dev).to(torch.bfloat16)
        S = torch.randn(NP * Ho * Wo, mid, device=dev).to(torch.bfloat16)
        da = torch.randn(NP * Ho * Wo, mid, device=dev).to(torch.bfloat16)
        c1, c2 = coef(mid), coef(mid)
        b2 = torch.randn(2, mid, device=dev) * 0.01
        ws = torch.randn(mid, 9, device=dev) * 0.3
        args = (da, S, E, c2, b2, c1, ws, NP, H, W, mid, s, PS)
        ref_dE, ref_part = run(0, 0, args)
        # mode 1: stride 1 -> one-pass v7, stride 2 -> v6; mode 2: two-pass v6 for both.  dE must be bit-identical; the
        # per-worker partial sums are only comparable after the sum over workers (tile -> worker maps differ)
        for mode, thi in (((1, 0), (1, 8), (1, 4)) if s == 2 else ((1, 0), (1, 8), (2, 0))):
            dE, part = run(mode, thi, args)
            same = torch.equal(dE.view(torch.int16), ref_dE.view(torch.int16))
            err = (dE.float() - ref_dE.float()).abs().max().item()
            ps, rs = part.double().sum(0), ref_part.double().sum(0)
            perr = ((ps - rs).abs().max() / rs.abs().max()).item()
            print(f"{tag} NP={NP} P={PS} mode={mode} thi={thi}: dE bit-identical={same} max|ddE|={err:.3e} sum-of-partials rel={perr:.3e}", flush=True)
            ok = ok and same and perr < 2e-6
print("CHECK", "PASS" if ok else "FAIL")

if "--time" in sys.argv:
    B, T = 32, 16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for tag, ci, H, W, s in shapes:
        mid = ci * 7
        Ho, Wo = H // s, W // s
        Mi, Mo = B * T * H * W, B * T * Ho * Wo
        E = torch.randn(Mi, mid, device=dev).to(torch.bfloat16)
        S = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
        da = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
        dE = torch.empty_like(E)
        c1, c2 = coef(mid), coef(mid)
        b2 = torch.randn(2, mid, device=dev) * 0.01
        ws = torch.randn(mid, 9, device=dev) * 0.3
        part = torch.empty(148, 11, mid, device=dev)
        nbytes = (2 * Mo + 2 * Mi) * mid * 2
        for mode, thi in ((0, 0), (1, 0)) + (((1, 8),) if s == 2 else ((1, 8), (2, 0))):
            os.environ["DWN_SDW_TMA"] = str(mode)
            os.environ["DWN_SDW_THI"] = str(thi)
            fn = lambda: call("dwn_sdw_bwd", da, S, E, c2, b2, c1, ws, dE, part, 148, B * T, H, W, mid, s, 1, st)
            fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            print(f"sdw_bwd {tag} tma={mode} thi={thi}: {ms:.3f} ms  {nbytes / ms * 1e-6:7.1f} GB/s  {nbytes / ms * 1e-6 / 6539.5 * 100:5.1f}% of HBM peak", flush=True)
        del E, S, da, dE

if "--sweep" in sys.argv:  # workers per channel chunk (the engine's _p_sdw)
    B, T = 32, 16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    os.environ["DWN_SDW_TMA"] = "1"
    os.environ["DWN_SDW_THI"] = "0"
    for tag, ci, H, W, s in shapes:
        mid = ci * 7
        Ho, Wo = H // s, W // s
        Mi, Mo = B * T * H * W, B * T * Ho * Wo
        E = torch.randn(Mi, mid, device=dev).to(torch.bfloat16)
        S = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
        da = torch.randn(Mo, mid, device=dev).to(torch.bfloat16)
        dE = torch.empty_like(E)
        c1, c2 = coef(mid), coef(mid)
        b2 = torch.randn(2, mid, device=dev) * 0.01
        ws = torch.randn(mid, 9, device=dev) * 0.3
        nbytes = (2 * Mo + 2 * Mi) * mid * 2
        for PS in (21, 37, 42, 74, 111, 148, 296):
            part = torch.empty(PS, 11, mid, device=dev)
            fn = lambda: call("dwn_sdw_bwd", da, S, E, c2, b2, c1, ws, dE, part, PS, B * T, H, W, mid, s, 1, st)
            fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            print(f"sweep sdw_bwd {tag} P={PS}: {ms:.3f} ms  {nbytes / ms * 1e-6 / 6539.5 * 100:5.1f}% of HBM peak", flush=True)
        del E, S, da, dE
