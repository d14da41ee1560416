"""Runs only the FPS kernels (sa1 shape, and one config-3 cloud batch) -- the command profiled by ncu --set full."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
xyz = I.facade_batch(32, 4096, 9, 11).cuda()[:, :, :3]
start = I.start_indices(32, 4096, 1).cuda()
for _ in range(3):
    pn2.farthest_point_sample(xyz, 1024, start=start)
cube = I.cube_xyz(16, 65536, 0).cuda()
start3 = I.start_indices(16, 65536, 2).cuda()
for _ in range(2):
    pn2.farthest_point_sample(cube, 2048, start=start3)
torch.cuda.synchronize()
print("ok")
