"""L2 -> shared memory ingest rates (profiling hook, GPU box only): cp.async.bulk vs ld.global.v4."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200._lib import check  # noqa: E402
from tests.native.hooks import load  # noqa: E402

lib = load()

MHZ = 1965.0
src = torch.randint(0, 255, (64 << 20,), dtype=torch.uint8, device="cuda")
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 50
print("grid mode bytes inflight shared  cyc/round  B/clk/SM  us/round")
for grid in ((1, 16, 128) if "--all" in sys.argv else ()):
    for mode in (0, 1):
        for bytes_, inflight in ((10240, 2), (10240, 6), (2560, 8), (2560, 24), (20480, 6), (1024, 20)):
            for shared in (1, 0):
                out.zero_()
                check(lib.cvb_bench_ingest(grid, mode, bytes_, inflight, shared, iters, src.data_ptr(), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                cyc = out[:grid].max().item() / iters
                tot = bytes_ * inflight
                print(f"{grid:4d} {mode:4d} {bytes_:6d} {inflight:4d} {shared:6d} {cyc:10.0f} {tot / cyc:9.1f} {cyc / MHZ:8.2f}")

print("\nall-gather pattern (slices rewritten every round by all CTAs)")
print("grid mode wmode bytes inflight  ingest cyc/round  B/clk/SM   sync cyc/round")
buf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
ctr = torch.zeros(64, dtype=torch.int32, device="cuda")
out2 = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
for grid in (16, 128):
    for mode in (0, 1):
        for wmode in (0, 1, 2):
            for bytes_, inflight in ((10240, 2), (20480, 6), (20480, 8)):
                if (bytes_ * inflight) % (grid * 16):
                    continue
                check(lib.cvb_bench_allgather(grid, mode, wmode, bytes_, inflight, iters, buf.data_ptr(), ctr.data_ptr(), out2.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                o = out2[:2 * grid].view(grid, 2).double()
                cin, csy = o[:, 0].max().item() / iters, o[:, 1].max().item() / iters
                print(f"{grid:4d} {mode:4d} {wmode:5d} {bytes_:6d} {inflight:4d} {cin:14.0f} {bytes_ * inflight / cin:12.1f} {csy:14.0f}")
