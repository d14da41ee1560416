"""Config-3 FPS (64 clouds x 65536 points -> npoint samples) under the PN2_FPS_P / PN2_FPS_CL tuning knobs (set in the
environment before the library is loaded).  Usage: PN2_FPS_P=16 PN2_FPS_CL=8 python profiles/fps_cfg3_sweep.py [npoint] [B]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
npoint = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cube = I.cube_xyz(B, 65536, 0).cuda()
start = I.start_indices(B, 65536, 2).cuda()
pn2.farthest_point_sample(cube, 64, start=start)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
idx = pn2.farthest_point_sample(cube, npoint, start=start)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("P=%s CL=%s  B=%d npoint=%d: %.2f ms, %.3f us/iteration, checksum %d" % (
    os.environ.get("PN2_FPS_P", "-"), os.environ.get("PN2_FPS_CL", "-"), B, npoint, ms, ms * 1e3 / npoint, int(idx.sum())))
