"""Top stall-sample SASS instructions of an `ncu --page source --csv` export (first kernel in the file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and len(r) == len(hdr): data.append(r)
si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed")
tot = sum(int(r[si] or 0) for r in data); tex = sum(int(r[ei] or 0) for r in data)
print("instructions %d, samples %d, warp-instr executed %d" % (len(data), tot, tex))
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
for i in sorted(order):
    r = data[i]
    print("%5d %6.2f%% smp  exec %9s  %s" % (i, 100.0 * int(r[si] or 0) / max(tot, 1), r[ei], r[1].strip()[:110]))
