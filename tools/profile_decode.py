"""Kernel-level breakdown of one stage-6 conversion run (tools/bench_decode.py's workload) with torch.profiler (CUPTI):
top kernels by device time, device-busy time vs wall time.  GPU box only; not a benchmark (profiler overhead)."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200 import cycle, synth  # noqa: E402

N_UTT = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
dev = torch.device("cuda", 0)
enc, dec, y0d1 = synth.build_models(1024, 32, 2, 50, 4, seed=1, device=dev)
enc.eval(); dec.eval()
x, cv, sc, tc = (t.to(dev) for t in synth.make_batch(N_UTT, T, 7, 2, 50))
y0e = torch.zeros(N_UTT, 1, 64, device=dev)
y0d = y0d1.to(dev).repeat(N_UTT, 1, 1).contiguous()
for _ in range(2):
    cycle.convert(enc, dec, x, tc, lat_dim=32, y0_enc=y0e, y0_dec=y0d)
torch.cuda.synchronize()
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    cycle.convert(enc, dec, x, tc, lat_dim=32, y0_enc=y0e, y0_dec=y0d)
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in ev:
    d = tot.setdefault(e.name[:90], [0.0, 0])
    d[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    d[1] += 1
busy = sum(v[0] for v in tot.values()) / 1e3
print(f"wall {wall:.1f} ms, device busy {busy:.1f} ms, {len(ev)} device activities")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {us / 1e3:9.3f} ms  x{n:<4d} {name}")
