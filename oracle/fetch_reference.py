"""Recipe: stage the UNMODIFIED reference files of the hot path (and its two callers) under oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (reference sources never enter this repository's history)
but NOT gpurun-ignored, so the staged files travel to the GPU box with the snapshot -- /root/reference does not
exist there.  __graft_entry__.build() runs this wherever /root/reference is mounted; on the GPU box the staged
copy is used as it arrived.  Consumers: tests/ (drop-in tests of the unchanged get_model / modelTraining /
modelTesting on this repo's operators) and bench.py's reference legs (the real reference on the host cores and,
eagerly, on the GPU).  Nothing in the product package may import from here (tests/test_abi.py enforces it).

    python oracle/fetch_reference.py            # copies, prints what it staged
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PN2_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the hot path, its caller, and the two loops / datasets either side of it (f)
FILES = (
    "models/pointnet2_utils.py",
    "models/pointnet2_sem_seg.py",
    "localfunctions.py",
    "provider.py",
    "sem_seg_training.py",
    "sem_seg_testing.py",
    "geofunction.py",
)


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose=True):
    """Copy FILES from the mounted reference tree; returns the number of files staged (0 if the tree is absent)."""
    if not os.path.isdir(SRC):
        if verbose:
            print("fetch_reference: %s is not mounted; keeping %s as it is (%s)" % (
                SRC, DST, "present" if os.path.isdir(DST) else "absent"))
        return 0
    lines = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
        lines.append("%s  %s" % (_sha(dst), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print("fetch_reference: staged %d unmodified reference files under %s" % (len(FILES), DST))
    return len(FILES)


if __name__ == "__main__":
    sys.exit(0 if stage() or os.path.isdir(DST) else 1)
