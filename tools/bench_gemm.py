"""Timing of the dense products of one pass at the bench shape: split-precision tcgen05 GEMM vs cuBLAS fp32 (GPU box only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200._lib import check, lib, ptr  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
shapes = [("gx enc", 6400, 3072, 486, 0, 1), ("gx dec", 6400, 3072, 306, 0, 1), ("dW_hh rz", 2048, 1024, 6400, 1, 0),
          ("dW_hh n", 1024, 1024, 6400, 1, 0), ("dW_x enc", 3072, 486, 6400, 1, 0), ("dW_y", 3072, 64, 6400, 1, 0),
          ("dW_o", 64, 1024, 6400, 1, 0), ("dxc enc", 6400, 486, 3072, 0, 0), ("conv1 tap", 7040, 486, 162, 0, 1),
          ("conv0 tap", 7040, 162, 54, 0, 1), ("dconv1 w", 486, 162, 7040, 1, 0), ("dconv1 in", 7040, 162, 486, 0, 0)]
only = sys.argv[1] if len(sys.argv) > 1 else ""
shapes = [s_ for s_ in shapes if only in s_[0]]
print(f"{'product':12s} {'M':>5s} {'N':>5s} {'K':>5s}  tc_us  cublas_us  tc TFLOP/s(alg)")
for name, M, N, K, ta, tb in shapes:
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((N, K) if tb else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    res = []
    for mode in ("tc", "cublas"):
        def run():
            if mode == "tc":
                check(lib.cvb_gemm_tc(ta, tb, M, N, K, ptr(A), A.shape[1], ptr(B), B.shape[1], 0, None, ptr(C), N, 0, st))
            else:
                os.environ["CVB_GEMM"] = "cublas"
                check(lib.cvb_gemm(ta, tb, M, N, K, 1.0, ptr(A), A.shape[1], ptr(B), B.shape[1], 0.0, ptr(C), N, st))
                os.environ.pop("CVB_GEMM")
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 100.0)
    print(f"{name:12s} {M:5d} {N:5d} {K:5d} {res[0]:7.1f} {res[1]:9.1f} {2.0 * M * N * K / res[0] / 1e6:9.1f}")
