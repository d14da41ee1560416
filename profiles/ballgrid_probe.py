import sys, importlib, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
x = I.facade_batch(32,4096,9,11).cuda()
xyz = x[:,:,:3]
torch.manual_seed(0)
_, nx = pn2.farthest_point_sample(xyz, 1024, return_xyz=True)
for _ in range(3):
    idx = pn2.query_ball_point(0.1, 32, xyz, nx)
_, nx2 = pn2.farthest_point_sample(nx, 256, return_xyz=True)
for _ in range(3):
    idx = pn2.query_ball_point(0.2, 32, nx, nx2)
torch.cuda.synchronize()
