"""Times the fused inference kernel of each set-abstraction level (csrc/sa_fused.cu) at the SSG network's shapes,
batch B (default 32): CUDA events on the launching stream, warm, median of 9; algorithmic bytes = indices + gathered
source rows (fp32, each read once per use) + pooled output.  Usage: python profiles/sa_fused_bench.py [B] [--once]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
mods = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules")
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
once = "--once" in sys.argv
pn2.set_precision("bf16")
torch.manual_seed(0)
net = pn2.get_model(18, 3).cuda().eval()
x = I.facade_batch(B, 4096, 9, 3).cuda()
xyz = x[:, :, :3]
feats = x
out = {}
with torch.no_grad():
    for name, sa in (("sa1", net.sa1), ("sa2", net.sa2), ("sa3", net.sa3), ("sa4", net.sa4)):
        N = xyz.shape[1]
        start = I.start_indices(B, N, 1).cuda()
        _, new_xyz = pn2.farthest_point_sample(xyz, sa.npoint, start=start, return_xyz=True)
        idx = pn2.query_ball_point(sa.radius, 32, xyz, new_xyz)
        fn = lambda: mods.sa_fused_eval(idx, sa.mlp_convs, sa.mlp_bns, new_xyz, xyz, feats)
        res = fn()
        if not once:
            lib = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200._lib")
            for _ in range(2):
                fn()
            lib.time_entry_point("pn2_sa_fused_eval")
            for _ in range(9):
                fn()
            ts = sorted(ms for _, _, ms in lib.timed_calls())
            lib.time_entry_point(None)
            D = feats.shape[2]
            M = B * sa.npoint * 32
            widths = [c.out_channels for c in sa.mlp_convs]
            K = [D + 3] + widths[:-1]
            flop = 2.0 * M * sum(k * n for k, n in zip(K, widths))
            alg = M * (8 + (D + 3) * 4) + B * sa.npoint * widths[-1] * 4
            ms = ts[len(ts) // 2]
            out[name] = {"rows": M, "mlp": [D + 3] + widths, "ms": round(ms, 4), "TFLOPs": round(flop / ms / 1e9, 2),
                         "alg_GBps": round(alg / ms / 1e6, 1)}
        xyz, feats = new_xyz, res
torch.cuda.synchronize()
print(json.dumps(out, indent=1))
