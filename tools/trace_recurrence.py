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


# event names of the one-exchange kernels (gru_tc2.cu, gru_tc2_bwd.cu): same slots, other roles
NAMES_ONE = {0: "fin: step start", 1: "fin: D1 complete seen", 2: "fin: exchange copies issued", 3: "fin: inbox complete", 4: "fin: W_y y / q accumulator seen",
             10: "fin: gates done", 7: "fin: h / dgh published, counter-H barrier arrived", 6: "fin: o / dgi staged, swap copies issued",
             25: "fin: partial (D3) complete seen", 26: "fin: partial added to the fixed-point totals", 9: "fin: counter Y released",
             14: "prod: counter H complete seen", 20: "aux: counter Y complete seen", 27: "aux: totals read", 21: "aux: own quarter staged, swap copies issued",
             22: "aux (w8): counter H released", 12: "mma: y / dy operand complete, chain issued", 13: "mma: o / dgi operand complete"}
NAMES_ONE.update({32 + i: f"prod: chunk {i} slot free, issuing" for i in range(8)})
NAMES_ONE.update({40 + i: f"mma: chunk {i} full" for i in range(8)})
NAMES_ONE.update({48 + i: f"mma: chunk {i} issued+committed" for i in range(8)})


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


PH_FWD = [("fin  step top -> D1 (W_hh h) complete", 0, 0, 1, 0), ("fin  D1 drained, exchange copies issued", 1, 0, 2, 0),
          ("fin  -> W_y y accumulator complete", 2, 0, 3, 0), ("fin  -> inbox complete", 3, 0, 4, 0), ("fin  partial sums + gates", 4, 0, 5, 0),
          ("fin  splits, o staged, h published, proxy fence", 5, 0, 6, 0), ("fin  -> D3 (partial y) complete", 6, 0, 7, 0),
          ("fin  drain of D3 -> part", 7, 0, 8, 0), ("fin  barrier with the aux warps", 8, 0, 9, 0), ("fin  red.release (counter A)", 9, 0, 10, 0),
          ("     own release A -> aux sees all arrived", 10, 0, 13, 1), ("     own release A -> prod sees all arrived", 10, 0, 11, 1),
          ("fin    drain: TMEM loads + stores", 7, 0, 19, 0), ("fin    drain: proxy fence", 19, 0, 8, 0),
          ("aux  partials pulled (bulk copy)", 13, 1, 14, 1), ("aux  sums + barrier", 14, 1, 16, 1), ("aux  publication (sum of subsets, split, stores)", 16, 1, 17, 1),
          ("aux  proxy fence + barrier", 17, 1, 18, 1), ("aux  red.release (counter B)", 18, 1, 15, 1),
          ("     own release B -> prod sees all arrived", 15, 1, 12, 1), ("     prod saw B -> W_y y accumulator complete", 12, 1, 3, 1),
          ("     whole step (fin step top to step top)", 0, 0, 0, 1)]


def report_phases(ph, table):
    """ph [2 steps][16 slots][256 CTAs] clock64 stamps; differences are taken inside one CTA (one SM clock)."""
    print("  per-CTA phase durations at steps 40/41, distribution over the CTAs (cycles): min / median / p90 / max  [slowest CTAs]")
    for name, s0, t0, s1, t1 in table:
        a, b = ph[t0, s0, :128].astype(np.float64), ph[t1, s1, :128].astype(np.float64)
        ok = (a > 0) & (b > 0)
        if not ok.any():
            continue
        d = (b - a)[ok]
        idx = np.nonzero(ok)[0][np.argsort(-d)[:4]]
        print(f"    {name:52s} {d.min():7.0f} {np.median(d):7.0f} {np.percentile(d, 90):7.0f} {d.max():7.0f}   {list(idx)}")


def report(path, names):
    raw = np.fromfile(path, dtype=np.int64)
    n_rows = T + 1 if "eval" not in path else raw.size // 64
    tr = raw[:n_rows * 64].reshape(-1, 64)
    if "fwd" in path and raw.size >= n_rows * 64 + 56 * 256:
        report_phases(raw[n_rows * 64 + 8 * 256:n_rows * 64 + 56 * 256].reshape(2, 24, 256), PH_FWD)
    if "fwd" in path and raw.size > n_rows * 64:
        sk = raw[n_rows * 64:n_rows * 64 + 8 * 256].reshape(8, 256)[:, :128].astype(np.float64)
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
    from cyclevae_vc_b200._lib import lib
    report(p.replace("trace_bwd", "trace_fwd"), NAMES_ONE if lib.cvb_last_recurrence_hops(0) == 1 else NAMES_FWD)
    report(p, NAMES_ONE if lib.cvb_last_recurrence_hops(1) == 1 else NAMES_BWD)
