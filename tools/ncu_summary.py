"""Summaries of ncu reports for profiles/:
    python tools/ncu_summary.py full  gpurun_out/r02_k_*.ncu-rep        -> CSV rows (one per captured launch)
    python tools/ncu_summary.py list  gpurun_out/r02_launches.csv       -> per-kernel launch counts / total time / share
"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL = [("gpu__time_duration.sum", "duration_us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__cluster_dim_x", "cluster", 1) if False else ("launch__grid_size", "grid", 1),
        ("launch__registers_per_thread", "regs", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
        ("dram__bytes_read.sum", "dram_read_MB", 1e-6), ("dram__bytes_write.sum", "dram_write_MB", 1e-6),
        ("lts__t_sectors.sum", "l2_MB", 32e-6), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", 1),
        ("smsp__inst_executed.sum", "warp_insts", 1)]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def scale(unit):
    u = unit.strip().lower()
    return {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)


if sys.argv[1] == "full":
    seen = []
    cols = []
    for m, name, _ in FULL:
        if name not in seen:
            seen.append(name)
            cols.append((m, name, _))
    print("kernel," + ",".join(n for _, n, _ in cols))
    for path in sys.argv[2:]:
        hdr, units, rows = raw(path)
        for r in rows:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("cvb::", "")
            vals = []
            for m, n, k in cols:
                if m in hdr:
                    i = hdr.index(m)
                    v = to_float(r[i])
                    v = None if v is None else v * scale(units[i]) * k
                    vals.append("" if v is None else f"{v:.4g}")
                else:
                    vals.append("")
            print(name + "," + ",".join(vals))
else:
    rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = to_float(r[vi])
        if v is None:
            continue
        a = agg[r[ki].split("(")[0]]
        a[0] += 1
        a[1] += v * scale(r[ui]) * 1e-3
    tot = sum(a[1] for a in agg.values())
    print(f"# total device time in captured launches: {tot / 1e3:.2f} ms")
    print("kernel,launches,total_us,share,avg_us")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[1] / a[0]:.1f}")
