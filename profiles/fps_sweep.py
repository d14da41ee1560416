"""FPS timing sweep over the points-per-thread knob (PN2_FPS_P) at the network's sa1/sa2 shapes and one
config-3-shaped batch.  Usage: python profiles/fps_sweep.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


out = {}
for B, N, S in ((32, 4096, 1024), (32, 1024, 256), (128, 4096, 1024)):
    xyz = I.facade_batch(B, N, 9, 11).cuda()[:, :, :3]
    start = I.start_indices(B, N, 1).cuda()
    for P in (0, 8, 16, 32):
        os.environ["PN2_FPS_P"] = str(P)
        ms = timeit(lambda: pn2.farthest_point_sample(xyz, S, start=start))
        out["fps_B%d_%d_%d_P%d" % (B, N, S, P)] = {"ms": round(ms, 4), "us_per_iter": round(ms * 1e3 / S, 4)}
cube = I.cube_xyz(18, 65536, 0).cuda()
start3 = I.start_indices(18, 65536, 2).cuda()
for P in (0, 8, 16, 32):
    os.environ["PN2_FPS_P"] = str(P)
    ms = timeit(lambda: pn2.farthest_point_sample(cube, 2048, start=start3), reps=3, warm=1)
    out["fps_cluster8_B18_65536_2048_P%d" % P] = {"ms": round(ms, 4), "us_per_iter": round(ms * 1e3 / 2048, 4)}
os.environ["PN2_FPS_P"] = "0"
big = I.facade_batch(32, 8192, 9, 5).cuda()[:, :, :3]
startb = I.start_indices(32, 8192, 3).cuda()
ms = timeit(lambda: pn2.farthest_point_sample(big, 2048, start=startb))
out["fps_B32_8192_2048_P0"] = {"ms": round(ms, 4), "us_per_iter": round(ms * 1e3 / 2048, 4)}
print(json.dumps(out, indent=1))
