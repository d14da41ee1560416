"""CPU model of the planned cell-grid ball query: bins a cloud into cells of edge
h = r * (1 + MARGIN), evaluates the reference's fp32 expanded-form distance only for the points of the 27 cells around a
query, and selects "the first nsample in-radius indices in index order, padded with the first" through a per-query bitmap.
Checks that the result is IDENTICAL to the brute-force C oracle (oracle/pn2_oracle.c, itself pinned to the reference's
fixtures) and reports the work reduction.  A study script, not product code: python profiles/cellgrid_model.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
from oracle import c_oracle as C

MARGIN = 1e-2      # the acceptance test runs in fp32 expanded form: a point up to ~1e-5 beyond r can still pass at unit scale


def expanded_d(q, p):
    """square_distance of pointnet2_utils.py:37-39 in fp32 operation order for one query against points p [n, 3]"""
    f = np.float32
    mm = f(q[0]) * p[:, 0]
    mm = (f(q[1]) * p[:, 1].astype(np.float64) + mm.astype(np.float64)).astype(np.float32)      # fma
    mm = (f(q[2]) * p[:, 2].astype(np.float64) + mm.astype(np.float64)).astype(np.float32)      # fma
    qn = f(f(f(q[0] * q[0]) + f(q[1] * q[1])) + f(q[2] * q[2]))
    pn = ((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]).astype(np.float32) + p[:, 2] * p[:, 2]).astype(np.float32)
    return ((f(-2.0) * mm + qn).astype(np.float32) + pn).astype(np.float32)


def grid_ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    r2 = np.float32(np.float64(radius) ** 2)
    h = np.float32(radius * (1.0 + MARGIN))
    out = np.empty((B, S, nsample), np.int64)
    evals = 0
    for b in range(B):
        lo = xyz[b].min(0)
        cell = np.floor((xyz[b] - lo) / h).astype(np.int64)
        dims = cell.max(0) + 1
        key = (cell[:, 0] * dims[1] + cell[:, 1]) * dims[2] + cell[:, 2]
        order = np.argsort(key, kind="stable")                    # counting sort on the device; stable keeps index order per cell
        starts = np.searchsorted(key[order], np.arange(dims.prod() + 1))
        qcell = np.floor((new_xyz[b] - lo) / h).astype(np.int64)
        for s in range(S):
            bitmap = np.zeros(N, bool)
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        c = qcell[s] + (dx, dy, dz)
                        if (c < 0).any() or (c >= dims).any():
                            continue
                        k = (c[0] * dims[1] + c[1]) * dims[2] + c[2]
                        members = order[starts[k]:starts[k + 1]]
                        if members.size:
                            d = expanded_d(new_xyz[b, s], xyz[b, members])
                            evals += members.size
                            bitmap[members[~(d > r2)]] = True          # NaN counts as inside, as `sqrdists > r**2` does
            hit = np.flatnonzero(bitmap)[:nsample]
            if hit.size == 0:
                out[b, s] = N
            else:
                out[b, s, :hit.size] = hit
                out[b, s, hit.size:] = hit[0]
    return out, evals


def main():
    rng = np.random.RandomState(0)
    cases = [("facade sa1", I.facade_batch(2, 4096, 9, 11)[:, :, :3].contiguous().numpy(), 1024, 0.1),
             ("facade sa2-like", I.facade_batch(2, 1024, 9, 12)[:, :, :3].contiguous().numpy(), 256, 0.2),
             ("unit cube", rng.rand(2, 2048, 3).astype(np.float32), 512, 0.1),
             ("duplicates + lattice", np.round(rng.rand(2, 1500, 3) * 10).astype(np.float32) / 10, 300, 0.1)]
    for name, xyz, S, r in cases:
        B, N, _ = xyz.shape
        idx = C.fps(xyz, S, rng.randint(0, N, size=(B,)).astype(np.int64))
        new = np.stack([xyz[b, idx[b]] for b in range(B)])
        want = C.ball_query(r, 32, xyz, new)
        got, evals = grid_ball_query(r, 32, xyz, new)
        same = np.array_equal(got, want)
        print("%-22s N=%d S=%d r=%.1f: identical to the brute-force oracle: %s; distance evaluations %.1f per query (brute force %d)"
              % (name, N, S, r, same, evals / (B * S), N))
        assert same, name


if __name__ == "__main__":
    main()
