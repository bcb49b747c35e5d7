"""Top stalled SASS instructions of an ncu report's source page:  ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_hot.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
sb = hdr.index("stall_long_sb")
data = []
for k, r in enumerate(rows[h + 1:]):
    try:
        data.append((float(r[idx]), k, r[src], r[ex], r[sb]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, reverse=True)[:n]:
    print(f"{d[0]:7.0f} {100 * d[0] / tot:5.1f}%  #{d[1]:5d} exec {d[3]:>9s} long_sb {d[4]:>6s}  {d[2][:100]}")
