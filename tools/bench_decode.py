"""Stage-6 conversion throughput (BASELINE.json configs[2]: 512 utterances x 800 frames x 50 mcep, eval mode, 1 ENC + 1 DEC
pass per utterance, decode_*.py:303-305,318) through cycle.convert on one GPU.  Prints one JSON line (GPU box only)."""
import json
import os
import sys

import torch

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


def run():
    return cycle.convert(enc, dec, x, tc, lat_dim=32, y0_enc=y0e, y0_dec=y0d)


for _ in range(2):
    out = run()
torch.cuda.synchronize()
import ctypes as C  # noqa: E402
from cyclevae_vc_b200._lib import lib  # noqa: E402
lib.cvb_profile_reset()
lib.cvb_profile_enable(2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
lib.cvb_profile_enable(0)
prof = {}
for kind, name in ((0, "recurrence"), (2, "gemm")):
    tot, n = C.c_float(0), C.c_int(0)
    lib.cvb_profile_summary(kind, C.byref(tot), C.byref(n))
    prof[name] = {"ms_per_run": tot.value / 3, "launches_per_run": n.value // 3}
print(json.dumps({"workload": f"stage-6 conversion, {N_UTT} utterances x {T} frames, hu1024 ld32, 1 ENC + 1 DEC pass, eval",
                  "ms": ms, "frames_per_s": N_UTT * T / (ms * 1e-3), "finite": bool(torch.isfinite(out).all()), "profile": prof}))
