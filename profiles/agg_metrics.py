"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv`
pass over a few eager training steps: launches, time share, DRAM traffic per launch and the DRAM GB/s that traffic means
at the (cold-cache, serialised) ncu duration.  Usage: python profiles/agg_metrics.py metrics.csv [steps]"""
import collections
import csv
import sys

UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "": 1.0}


def main(path, steps=1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_launch = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        key = (row["ID"], row["Kernel Name"])
        per_launch.setdefault(key, {})[row["Metric Name"]] = v * UNIT.get(row["Metric Unit"], 1.0)
    agg = collections.OrderedDict()
    for (_, name), m in per_launch.items():
        a = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "pct": 0.0, "regs": 0})
        a["n"] += 1
        a["t"] += m.get("gpu__time_duration.sum", 0.0)
        a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
        a["pct"] += m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["regs"] = max(a["regs"], int(m.get("launch__registers_per_thread", 0)))
    tot = sum(a["t"] for a in agg.values()) or 1.0
    print("%d launches, %.1f us of kernel time over %d step(s)" % (sum(a["n"] for a in agg.values()), tot * 1e6, steps))
    print("%6s %9s %6s %9s %10s %10s %8s %6s %5s  %s" % ("share", "us/step", "n/step", "us/launch", "rdMB/launch", "wrMB/launch",
                                                        "GB/s", "dram%", "regs", "kernel"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = a["n"]
        print("%5.1f%% %9.1f %6.1f %9.1f %10.2f %10.2f %8.0f %6.1f %5d  %s" % (
            100 * a["t"] / tot, a["t"] * 1e6 / steps, n / steps, a["t"] * 1e6 / n, a["rd"] / n / 1e6, a["wr"] / n / 1e6,
            (a["rd"] + a["wr"]) / a["t"] / 1e9 if a["t"] else 0.0, a["pct"] / n, a["regs"], name[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
