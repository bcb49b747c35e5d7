"""Operand ingest of a GEMM-like tile loop by thread-block clusters: unicast bulk copies vs. multicast of the block the
cluster shares (profiling hook, GPU box only).  Question answered: is the ~6.3 KB/clk L2 -> SM cap on DELIVERED bytes
or on L2 reads (one multicast read serving several SMs)?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200._lib import check  # noqa: E402
from tests.native.hooks import load  # noqa: E402
import ctypes as C  # noqa: E402

lib = load()
lib.cvb_bench_mcast.restype = C.c_int
lib.cvb_bench_mcast.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p]
src = torch.randint(0, 255, (64 << 20,), dtype=torch.uint8, device="cuda")
out = torch.zeros(160, dtype=torch.int64, device="cuda")
iters = 200
print("grid cluster mode  a_KB  b_KB b_all  cyc/round  delivered B/clk/SM  delivered KB/clk chip")
for grid in (148, 128):
    for cluster in (1, 2, 4):
        if grid % cluster:
            continue
        for mode in ((0,) if cluster == 1 else (0, 1)):
            for a_b, b_b in ((32768, 32768), (16384, 49152)):
                for b_all in (1, 0):
                    out.zero_()
                    check(lib.cvb_bench_mcast(grid, cluster, mode, a_b, b_b, b_all, iters, src.data_ptr(), out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
                    torch.cuda.synchronize()
                    cyc = out[:grid].max().item() / iters
                    tot = a_b + b_b
                    print(f"{grid:4d} {cluster:7d} {mode:4d} {a_b >> 10:5d} {b_b >> 10:5d} {b_all:5d} {cyc:10.0f} {tot / cyc:14.1f} {grid * tot / cyc / 1024:14.2f}")
