"""Kernel microbenchmarks (CUDA events on the launching stream, warm, median of N):
FPS / ball query / 3-NN at the network's level shapes and at BASELINE.json config 3
(FPS 65536 -> 16384, ball query r=0.1 k=32, batch 64).  Usage: python profiles/microbench.py [--config3]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I  # noqa: E402

pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
DEV = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {}
    B = 32
    xyz = I.facade_batch(B, 4096, 9, 11).to(DEV)[:, :, :3]
    lv = [xyz]
    for (N, S, r) in ((4096, 1024, 0.1), (1024, 256, 0.2), (256, 64, 0.4), (64, 16, 0.8)):
        start = I.start_indices(B, N, 1).to(DEV)
        src = lv[-1]
        ms = timeit(lambda: pn2.farthest_point_sample(src, S, start=start))
        _, new_xyz = pn2.farthest_point_sample(src, S, start=start, return_xyz=True)
        out["fps_%d_%d_B%d" % (N, S, B)] = {"ms": ms, "us_per_iter": ms * 1e3 / S,
                                              "gflops_nonfma": 9.0 * B * N * S / (ms * 1e-3) / 1e9}
        msb = timeit(lambda: pn2.query_ball_point(r, 32, src, new_xyz))
        out["ball_%d_%d_r%g_B%d" % (N, S, r, B)] = {"ms": msb, "gpairs_per_s_bruteforce_basis": B * S * N / (msb * 1e-3) / 1e9}
        msn = timeit(lambda: pn2.three_nn(src, new_xyz))
        out["three_nn_%d_%d_B%d" % (N, S, B)] = {"ms": msn, "gpairs_per_s": B * S * N / (msn * 1e-3) / 1e9}
        lv.append(new_xyz)
    if "--config3" in sys.argv:
        B3 = 64
        cube = I.cube_xyz(B3, 65536, 0).to(DEV)
        start = I.start_indices(B3, 65536, 2).to(DEV)
        ms = timeit(lambda: pn2.farthest_point_sample(cube, 16384, start=start), reps=3, warm=1)
        _, new_xyz = pn2.farthest_point_sample(cube, 16384, start=start, return_xyz=True)
        flop = 9.0 * B3 * 65536 * 16384
        out["config3_fps_65536_16384_B64"] = {"ms": ms, "us_per_iter": ms * 1e3 / 16384, "tflops_nonfma": flop / (ms * 1e-3) / 1e12,
                                               "clouds_per_s": B3 / (ms * 1e-3)}
        msb = timeit(lambda: pn2.query_ball_point(0.1, 32, cube, new_xyz), reps=3, warm=1)
        out["config3_ball_r0.1_k32_B64"] = {"ms": msb, "gpairs_per_s_bruteforce_basis": B3 * 16384 * 65536 / (msb * 1e-3) / 1e9,
                                             "clouds_per_s": B3 / (msb * 1e-3)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
