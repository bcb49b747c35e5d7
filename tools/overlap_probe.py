"""Experiment: does a capped-grid GEMM on a side stream overlap with the persistent recurrence kernels (128 of 148 SMs), and
what does it cost them?   python tools/overlap_probe.py [cap]"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cap = sys.argv[1] if len(sys.argv) > 1 else "16"
import cyclevae_vc_b200 as cvb  # noqa: E402
from cyclevae_vc_b200._lib import check, lib, ptr  # noqa: E402

B, T = 80, 80
enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda().train()
enc.apply(cvb.initialize)
x = torch.randn(B, T, 54, device="cuda", requires_grad=True)
y0 = torch.zeros(B, 1, 64, device="cuda")
M, N, K = 3072, 1024, 6400
A = torch.randn(K, M, device="cuda")
Bm = torch.randn(K, N, device="cuda")
C = torch.zeros(M, N, device="cuda")
side = torch.cuda.Stream()


def rec_pass():
    o, y, h = enc(x, y0, do=True, clamp_vae=True, lat_dim=32)
    o.square().sum().backward()


def timed(with_side):
    for _ in range(2):
        rec_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_side = 0
    if with_side:
        os.environ["CVB_GEMM_MAX_CTAS"] = cap
        with torch.cuda.stream(side):
            s0.record()
            for _ in range(6):
                check(lib.cvb_gemm_tc(1, 0, M, N, K, ptr(A), M, ptr(Bm), N, 1, None, ptr(C), N, 0, side.cuda_stream))
                n_side += 1
            s1.record()
        os.environ.pop("CVB_GEMM_MAX_CTAS", None)
    e0.record()
    for _ in range(4):
        rec_pass()
    e1.record()
    torch.cuda.synchronize()
    ms_side = s0.elapsed_time(s1) if with_side else 0.0
    return e0.elapsed_time(e1) / 4, ms_side, n_side


os.environ["CVB_GEMM_MAX_CTAS"] = cap
check(lib.cvb_gemm_tc(1, 0, M, N, K, ptr(A), M, ptr(Bm), N, 1, None, ptr(C), N, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
check(lib.cvb_gemm_tc(1, 0, M, N, K, ptr(A), M, ptr(Bm), N, 1, None, ptr(C), N, 0, torch.cuda.current_stream().cuda_stream))
e1.record()
torch.cuda.synchronize()
print(f"GEMM {M}x{N}x{K} alone with {cap} CTAs: {e0.elapsed_time(e1):.3f} ms")
os.environ.pop("CVB_GEMM_MAX_CTAS", None)
e0.record()
check(lib.cvb_gemm_tc(1, 0, M, N, K, ptr(A), M, ptr(Bm), N, 1, None, ptr(C), N, 0, torch.cuda.current_stream().cuda_stream))
e1.record()
torch.cuda.synchronize()
print(f"GEMM alone, full grid: {e0.elapsed_time(e1):.3f} ms")
a, _, _ = timed(False)
b, ms_side, n = timed(True)
print(f"ENC pass (fwd + BPTT + GEMMs) alone: {a:.3f} ms;  with {n} capped GEMMs on a side stream: {b:.3f} ms (side stream busy {ms_side:.3f} ms)")
