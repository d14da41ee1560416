"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=45):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("total %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%6.2f%% %9.1f us %5d x %8.1f us  %s" % (100 * v / tot, v / 1e3, n, v / 1e3 / n, k[:120]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
