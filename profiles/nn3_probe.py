"""three_nn through the cell grid vs the index-order scan at the network's feature-propagation shapes (32 facade clouds):
CUDA events around the eager call, median of 9.  Usage: python profiles/nn3_probe.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
ops = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.ops")


def timeit(fn, reps=9, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


B = 32
levels = [I.facade_batch(B, 4096, 9, 11).cuda()[:, :, :3].contiguous()]
for S in (1024, 256):
    src = levels[-1]
    idx = pn2.farthest_point_sample(src, S, start=I.start_indices(B, src.shape[1], 1).cuda())
    levels.append(torch.gather(src, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
for fine, coarse in ((levels[0], levels[1]), (levels[1], levels[2])):
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.THREE_NN_GRID = True
    gi, gw = ops.three_nn(fine, coarse, fallback_count=cnt)
    t_grid = timeit(lambda: ops.three_nn(fine, coarse))
    ops.THREE_NN_GRID = False
    si, sw = ops.three_nn(fine, coarse)
    t_scan = timeit(lambda: ops.three_nn(fine, coarse))
    ops.THREE_NN_GRID = False
    print("N %5d S %5d: grid %.1f us (%d of %d queries fell back), scan %.1f us, identical %s" % (
        fine.shape[1], coarse.shape[1], t_grid * 1e3, int(cnt), B * fine.shape[1], t_scan * 1e3,
        bool(torch.equal(gi, si) and torch.equal(gw, sw))))
