"""Per-kernel SASS census of libpn2b200.so: how many tcgen05 / TMA / mma.sync / cluster instructions each kernel holds.

    python profiles/sass_counts.py > profiles/r02_sass_tc.txt

Runs `cuobjdump -sass` on the shipped library (no GPU needed) and counts, per `Function :`, the mnemonics that prove
which hardware path a kernel uses (B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld/st,
tensor memory), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (cp.async.bulk), HMMA (mma.sync), SYNCS
(mbarrier), UCGABAR / CGA (cluster barriers), REDUX (warp reductions), RED / ATOM (global reductions).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "khairil_tum-facade_semantic_segmentation_b200", "libpn2b200.so")
PATTERNS = collections.OrderedDict([
    ("UTCHMMA", r"\bUTC[HQI]MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
    ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("SYNCS", r"\bSYNCS"),
    ("UCGABAR", r"\bUCGABAR"), ("REDUX", r"\bREDUX"), ("RED", r"\bRED\."), ("ATOM", r"\bATOMG?\b|\bATOM\."),
    ("FFMA", r"\bFFMA"), ("total", r"^\s+/\*[0-9a-f]{4}\*/"),
])


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    regs = {k: re.compile(v) for k, v in PATTERNS.items()}
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        for k, rx in regs.items():
            if rx.search(line):
                counts[cur][k] += 1
    names = demangle(order)
    cols = list(PATTERNS)
    print("# cuobjdump -sass %s : instruction counts per kernel (%d kernels, arch sm_100a)" % (os.path.relpath(SO, ROOT), len(order)))
    print("%-78s " % "kernel" + " ".join("%7s" % c for c in cols))
    key = lambda f: (-(counts[f]["UTCHMMA"] + counts[f]["LDTM"] + counts[f]["UTMALDG"] + counts[f]["UBLKCP"] + counts[f]["HMMA"]), names[f])
    for f in sorted(order, key=key):
        full = names[f].replace("(anonymous namespace)::", "").replace("pn2::", "")
        cut = full.rfind(">(") + 1 if ">(" in full else full.find("(")
        short = re.sub(r"^void ", "", full[:cut] if cut > 0 else full).replace("(int)", "").replace("(bool)", "")
        print("%-78s " % short[:78] + " ".join("%7d" % counts[f][c] for c in cols))
    tc = [f for f in order if counts[f]["UTCHMMA"]]
    print("# %d kernels issue tcgen05.mma (UTC*MMA), %d read tensor memory (LDTM), %d use TMA tensor copies, %d use mma.sync (HMMA)" % (
        len(tc), sum(1 for f in order if counts[f]["LDTM"]), sum(1 for f in order if counts[f]["UTMALDG"] or counts[f]["UTMASTG"]),
        sum(1 for f in order if counts[f]["HMMA"])))


if __name__ == "__main__":
    sys.exit(main())
