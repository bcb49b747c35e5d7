"""Phase timeline of the persistent tcgen05 recurrence kernels (profiling hook, GPU box only).

CVB_TRACE_FILE makes gru_ar_{fwd,bwd}_tc record clock64 stamps of CTA 0 per step; this script runs one
ENC forward+backward at the bench shape and prints, per event, the mean offset (in SM cycles and us at the
sampled clock) from the finaliser's start-of-step stamp, plus the mean step period.
    python tools/trace_recurrence.py [B] [T]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cyclevae_vc_b200 as cvb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 80
T = int(sys.argv[2]) if len(sys.argv) > 2 else 80
MHZ = 1965.0

NAMES_BWD = {0: "fin: step start", 1: "fin: accum_full seen", 2: "fin: exchange copies issued", 3: "fin: inbox complete",
             6: "fin: gate math + A2 staged", 7: "fin: dgh published", 8: "fin: fences done", 9: "fin: ctrA arrive",
             13: "prod: ctrB seen", 14: "prod: ctrA seen", 20: "aux: ctrA seen", 27: "aux: partials staged",
             21: "aux: dy summed + published", 22: "aux: ctrB arrive", 25: "aux: part_full seen", 26: "aux: D3 drained"}
NAMES_BWD.update({32 + i: f"prod: chunk {i} slot free, issuing" for i in range(8)})
NAMES_BWD.update({40 + i: f"mma: chunk {i} full" for i in range(8)})
NAMES_BWD.update({48 + i: f"mma: chunk {i} issued+committed" for i in range(8)})


NAMES_FWD = {0: "fin: step start", 1: "fin: accum_full seen", 2: "fin: exchange copies issued", 3: "fin: inbox complete",
             4: "fin: accum_full (y part) seen", 5: "fin: partial sums added", 10: "fin: gates done", 11: "fin: splits done", 6: "fin: gates + A2 staged", 7: "fin: h published", 8: "fin: proxy fence done", 9: "fin: ctrA arrive",
             13: "prod: ctrB seen", 14: "prod: ctrA seen", 20: "aux: ctrA seen", 27: "aux: partials staged",
             21: "aux: y summed + published", 22: "aux: ctrB arrive", 26: "aux: D3 drained"}
NAMES_FWD.update({32 + i: f"prod: chunk {i} slot free, issuing" for i in range(8)})
NAMES_FWD.update({40 + i: f"mma: chunk {i} full" for i in range(8)})
NAMES_FWD.update({48 + i: f"mma: chunk {i} issued+committed" for i in range(8)})


def run(tag):
    enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda().train()
    enc.apply(cvb.initialize)
    x = torch.randn(B, T, 54, device="cuda", requires_grad=True)
    y0 = torch.zeros(B, 1, 64, device="cuda")
    for i in range(3):
        path = os.path.join(ROOT, "gpurun_out", f"trace_{tag}_{i}.bin")
        os.environ["CVB_TRACE_FILE"] = path
        os.environ["CVB_TRACE_FILE_FWD"] = path.replace("trace_bwd", "trace_fwd")
        o, y, h = enc(x, y0, do=True, clamp_vae=True, lat_dim=32)
        o.square().sum().backward()
        torch.cuda.synchronize()
    os.environ.pop("CVB_TRACE_FILE", None)
    os.environ.pop("CVB_TRACE_FILE_FWD", None)
    return path


def report(path, names):
    raw = np.fromfile(path, dtype=np.int64)
    n_rows = (raw.size - (8 * 256 if raw.size % 64 == 0 and (raw.size - 8 * 256) % 64 == 0 and "fwd" in path else 0)) // 64
    tr = raw[:n_rows * 64].reshape(-1, 64)
    if "fwd" in path and raw.size > n_rows * 64:
        sk = raw[n_rows * 64:].reshape(8, 256)[:, :128].astype(np.float64)
        sk_names = ["fin: before ctrA release", "fin: after ctrA release", "aux: y published", "aux: after ctrB release", "prod: ctrA seen (t=40)",
                 "prod: ctrB seen (t=40)"]
        ref = sk[0].min()
        print("  cross-CTA skew at step 40 (globaltimer ns relative to the earliest CTA reaching its ctrA release):")
        for i, nm in enumerate(sk_names):
            v = sk[i][sk[i] > 0]
            if v.size:
                print(f"    {nm:28s} min {v.min() - ref:8.0f}  median {np.median(v) - ref:8.0f}  max {v.max() - ref:8.0f} ns")
    n_it = tr.shape[0]
    base = tr[:, 0].astype(np.float64)
    sel = slice(5, n_it - 2)
    period = np.diff(base[sel]).mean()
    print(f"{path}: {n_it} iterations, mean step period {period:.0f} cycles = {period / MHZ:.2f} us")
    if tr[0, 60] and tr[0, 62]:
        setup, first, total = tr[0, 61] - tr[0, 60], tr[0, 0] - tr[0, 60], tr[0, 62] - tr[0, 60]
        print(f"  kernel: setup {setup / MHZ:.1f} us, entry -> first step {first / MHZ:.1f} us, whole kernel {total / MHZ:.1f} us, "
              f"steps {period * (n_it - 1) / MHZ:.1f} us")
    rows = []
    for ev in range(60):
        v = tr[sel, ev].astype(np.float64)
        if (v == 0).all():
            continue
        off = (v - base[sel])
        rows.append((off.mean(), ev))
    for off, ev in sorted(rows):
        print(f"  ev{ev:2d} {names.get(ev, '?'):40s} {off:9.0f} cyc  {off / MHZ:7.2f} us")


NAMES_EVAL = {0: "fin: step start", 1: "fin: d1_full seen", 2: "fin: exchange copies issued", 3: "fin: inbox complete", 5: "fin: partial sums added",
              10: "fin: gates done", 7: "fin: h published", 8: "fin: proxy fence done", 9: "fin: counter arrive", 14: "prod: counter seen"}
NAMES_EVAL.update({40 + i: f"mma: chunk {i} full" for i in range(8)})


def run_eval():
    enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda().eval()
    enc.apply(cvb.initialize)
    x = torch.randn(B, T, 54, device="cuda")
    y0 = torch.zeros(B, 1, 64, device="cuda")
    path = os.path.join(ROOT, "gpurun_out", "trace_eval.bin")
    with torch.no_grad():
        for i in range(3):
            os.environ["CVB_TRACE_FILE_EVAL"] = path
            enc(x, y0, clamp_vae=True, lat_dim=32)
            torch.cuda.synchronize()
    os.environ.pop("CVB_TRACE_FILE_EVAL", None)
    return path


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if len(sys.argv) > 3 and sys.argv[3] == "eval":
        report(run_eval(), NAMES_EVAL)
        sys.exit(0)
    p = run("bwd")
    report(p.replace("trace_bwd", "trace_fwd"), NAMES_FWD)
    report(p, NAMES_BWD)
