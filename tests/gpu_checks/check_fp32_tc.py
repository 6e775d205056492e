"""fp32-accurate tensor-core GEMM (3-way bf16 split, dwn_gemm split=3) vs torch fp64 matmul, and eval-forward speed /
accuracy with it on and off.   python tests/gpu_checks/check_fp32_tc.py"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from sensorium_b200 import engine  # noqa: E402
from sensorium_b200._lib import call, gemm  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(0)
    for (M, N, K) in [(4096, 448, 64), (8192, 64, 448), (512, 4000, 2048), (1000, 256, 1792)]:
        A = torch.randn(M, K, device=dev, generator=g) * 3 + 1.0
        B = torch.randn(N, K, device=dev, generator=g)
        ref = (A.double() @ B.double().t())
        D = torch.empty(M, N, device=dev)
        engine.gemm_f32(st, False, dtype=0, A=A, B=B, lda=K, ldb=K, M=M, N=N, K=K, Z=1, D=D, d_dtype=0, ldd=N)
        D2 = torch.empty(M, N, device=dev)
        gemm(st, dtype=0, A=A, B=B, lda=K, ldb=K, M=M, N=N, K=K, Z=1, D=D2, d_dtype=0, ldd=N)
        torch.cuda.synchronize()
        scale = float(ref.abs().max())
        print(f"M={M} N={N} K={K}: split-TC max err/max {float((D.double() - ref).abs().max()) / scale:.2e}  "
              f"FFMA {float((D2.double() - ref).abs().max()) / scale:.2e}  torch fp32 "
              f"{float(((A @ B.t()).double() - ref).abs().max()) / scale:.2e}")
    # eval forward of the real architecture
    from oracle import dwiseneuro_oracle as O
    from sensorium_b200 import DwiseNeuro, constants
    from sensorium_b200.utils import init_weights
    from tests.shapes import TRUE_BATCH_KW
    torch.manual_seed(0)
    net = DwiseNeuro(readout_outputs=constants.num_neurons, **TRUE_BATCH_KW).to(dev)
    init_weights(net)
    net.eval()
    x = O.synthetic_clip(32, 16, 64, seed=0).to(dev)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = O.dwiseneuro_forward(x[:4], sd, O.make_cfg(constants.num_neurons, **TRUE_BATCH_KW), 0, False)
        for flag in (False, True):
            engine.FP32_TENSOR_CORES = flag
            y = net(x, 0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                y = net(x, 0)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            err = float((y[:4] - ref).abs().max() / ref.abs().max())
            print(f"eval fwd fp32, batch 32, one readout, tensor cores {flag}: {dt * 1e3:.1f} ms, rel err vs oracle {err:.2e}")


if __name__ == "__main__":
    main()
