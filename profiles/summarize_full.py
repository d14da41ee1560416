"""One block per captured launch from `ncu -i X.ncu-rep --page raw --csv` files (the --set full captures of
profiles/r01_capture.sh): duration, DRAM bytes and throughput, tensor-pipe and issue activity, occupancy limits and the
largest warp-stall reasons.  Usage: python profiles/summarize_full.py a.raw.csv [b.raw.csv ...] > summary.md"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (achieved occupancy)"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_registers", "CTAs/SM limit: registers"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM limit: shared memory"), ("launch__occupancy_limit_warps", "CTAs/SM limit: warps"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %")]


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
            col.setdefault(h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h, i)
        print("## %s\n" % path.split("/")[-1].replace(".raw.csv", ""))
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            print("### `%s`  grid %s block %s\n" % (name[:100], r[col["Grid Size"]], r[col["Block Size"]]))
            for key, label in WANT:
                i = col.get(key)
                if i is None:
                    i = next((j for j, h in enumerate(hdr) if h.endswith(key)), None)
                if i is not None and r[i] not in ("", "n/a"):
                    print("- %s: %s %s" % (label, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("- warp stalls per issue (top 4): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4]))
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
