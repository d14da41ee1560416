"""MLP-kernel microbenchmarks at the layer shapes of BASELINE.json config 2 (32 x 4096 points):
pn2_linear_fwd (+statistics), pn2_linear_bwd_data, pn2_linear_bwd_weight, pn2_bn_relu_bwd_reduce,
pn2_bn_relu_bwd_dz, through the C ABI, bf16 rows.  CUDA events on the launching stream, L2 flushed
between repetitions, median; GB/s on ALGORITHMIC bytes (each operand read/written once).

  python profiles/microbench_mlp.py                 # table + JSON
  python profiles/microbench_mlp.py --only wgrad --level sa1 --layer 3 --reps 1    # one launch (for ncu)
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
L = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200._lib")
DEV = "cuda"
B = 32
LEVELS = {  # name: (rows, [K0, widths...])
    "sa1": (B * 1024 * 32, [12, 32, 32, 64]), "sa2": (B * 256 * 32, [67, 64, 64, 128]),
    "sa3": (B * 64 * 32, [131, 128, 128, 256]), "sa4": (B * 16 * 32, [259, 256, 256, 512]),
    "fp4": (B * 64, [768, 256, 256]), "fp3": (B * 256, [384, 256, 256]), "fp2": (B * 1024, [320, 256, 128]),
    "fp1": (B * 4096, [128, 128, 128, 128]),
}


def ld(c):
    return (c + 7) // 8 * 8


def rows(M, C, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.zeros(M, ld(C), device=DEV, dtype=torch.bfloat16)
    x[:, :C] = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    return x


def timeit(fn, reps, flush):
    for _ in range(2 if reps > 1 else 0):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="fwd | dgrad | wgrad | bnred | bndz")
    ap.add_argument("--level", default=None)
    ap.add_argument("--layer", type=int, default=None)
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    lib = L.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV) if args.reps > 1 else None
    st = L.stream
    out = {}
    for name, (M, widths) in LEVELS.items():
        if args.level and name != args.level:
            continue
        for li in range(1, len(widths)):
            if args.layer and li != args.layer:
                continue
            K, N = widths[li - 1], widths[li]
            x, dz = rows(M, K, 1), rows(M, N, 2)
            z = torch.empty(M, ld(N), device=DEV, dtype=torch.bfloat16)
            da = torch.empty(M, ld(K), device=DEV, dtype=torch.bfloat16)
            W = (torch.randn(N, K, device=DEV) / K ** 0.5)
            sc, sh = torch.rand(K, device=DEV) + 0.5, torch.randn(K, device=DEV) * 0.1
            scn, shn = torch.rand(N, device=DEV) + 0.5, torch.randn(N, device=DEV) * 0.1
            mean, invstd = torch.randn(N, device=DEV) * 0.1, torch.rand(N, device=DEV) + 0.5
            dgb = torch.randn(2, N, device=DEV)
            accum = torch.zeros(4 * 2 * 4096, device=DEV, dtype=torch.float64)
            wp1 = torch.empty(lib.pn2_linear_wpack_bytes(K, N), device=DEV, dtype=torch.uint8)
            wp2 = torch.empty(lib.pn2_linear_wpack_bytes(N, K), device=DEV, dtype=torch.uint8)
            scratch = torch.empty(lib.pn2_linear_wgrad_scratch_bytes(M, K, N), device=DEV, dtype=torch.uint8)
            dW = torch.empty(N, K, device=DEV)
            ins, insh = (None, None) if li == 1 else (sc, sh)
            p = L.ptr
            ops = {
                "fwd": (lambda: L.call("pn2_linear_fwd", p(x), x.shape[1], 1, p(ins), p(insh), p(W), None, M, K, N, p(z),
                                       z.shape[1], 1, p(accum), p(wp1), st()), 2 * M * (ld(K) + ld(N))),
                "dgrad": (lambda: L.call("pn2_linear_bwd_data", p(dz), dz.shape[1], 1, p(W), M, K, N, p(da), da.shape[1], 1,
                                         p(wp2), st()), 2 * M * (ld(K) + ld(N))),
                "wgrad": (lambda: L.call("pn2_linear_bwd_weight", p(dz), dz.shape[1], 1, p(x), x.shape[1], 1, p(ins), p(insh),
                                         M, K, N, p(dW), p(scratch), st()), 2 * M * (ld(K) + ld(N))),
                "wgrad_acc": (lambda: L.call("pn2_linear_bwd_weight_accum", p(dz), dz.shape[1], 1, p(x), x.shape[1], 1, p(ins),
                                             p(insh), M, K, N, p(dW), st()), 2 * M * (ld(K) + ld(N))),
                "bnred": (lambda: L.call("pn2_bn_relu_bwd_reduce", p(dz), dz.shape[1], 1, p(z), z.shape[1], 1, p(scn), p(shn),
                                         p(mean), p(invstd), M, N, p(accum), st()), 2 * M * 2 * ld(N)),
                "bndz": (lambda: L.call("pn2_bn_relu_bwd_dz", p(dz), dz.shape[1], 1, p(z), z.shape[1], 1, p(scn), p(shn),
                                        p(mean), p(invstd), p(dgb[0]), p(dgb[1]), M, N, p(dz), dz.shape[1], 1, st()),
                         2 * M * 3 * ld(N)),
            }
            ops["fwd"][0]()        # z must hold real values for the BN kernels
            torch.cuda.synchronize()
            for op, (fn, nbytes) in ops.items():
                if args.only and op not in args.only.split(","):
                    continue
                if op == "wgrad_acc" and K % 4:
                    continue
                ms = timeit(fn, args.reps, flush)
                key = "%s.%d %s" % (name, li, op)
                out[key] = {"M": M, "K": K, "N": N, "us": round(ms * 1e3, 1), "GBps": round(nbytes / (ms * 1e-3) / 1e9, 1)}
                print("%-14s M=%8d K=%4d N=%4d  %8.1f us  %7.1f GB/s" % (key, M, K, N, ms * 1e3, nbytes / (ms * 1e-3) / 1e9),
                      flush=True)
            accum.zero_()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not args.only:
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench_mlp.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
