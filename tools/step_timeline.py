"""GPU busy time vs wall time of one benchmark step (torch.profiler, CUDA activities): are the gaps between kernels
(host launch overhead of the Python/autograd driver) or the kernels themselves the bound?  GPU box only."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200 import cycle, synth  # noqa: E402

HIDDEN, LAT, NSPK, NMCEP, STDIM, NCYC, T = 1024, 32, 2, 50, 4, 2, 80
B = int(sys.argv[1]) if len(sys.argv) > 1 else 80
dev = torch.device("cuda", 0)
enc, dec, y0d1 = synth.build_models(HIDDEN, LAT, NSPK, NMCEP, STDIM, seed=1, device=dev)
enc.train(); dec.train()
opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-4)
x, cv, sc, tc = (t.to(dev) for t in synth.make_batch(B, T, 100, NSPK, NMCEP))
y0e = torch.zeros(B, 1, 2 * LAT, device=dev)
y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()
flens = torch.full((B,), T, dtype=torch.int32, device=dev)


def step():
    opt.zero_grad()
    out, _ = cycle.cyc_forward(enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, y0_enc=y0e,
                               y0_dec=y0d, do=True)
    loss, _ = cycle.cyc_loss(out, x, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, flen_acc=None, select_utt_idx=list(range(B)), flens_dev=flens)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        step()
    e1.record()
    torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / N
agg = defaultdict(lambda: [0, 0.0])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
for e in evs:
    a = agg[e.name[:70]]
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
busy = sum(a[1] for a in agg.values()) / N / 1e3
print(f"wall {wall:.2f} ms/step (under the profiler), GPU kernel time {busy:.2f} ms/step, idle {wall - busy:.2f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{a[1] / N / 1e3:8.3f} ms {a[0] // N:5d}x {a[1] / a[0]:8.1f} us  {k}")
