"""pn2_mlp_bwd_layer (csrc/bwd_fused.cu) at the layer shapes of BASELINE.json config 2 (32 x 4096 points) that take the
fused kernel, through the C ABI, bf16 rows.  CUDA events on the launching stream, L2 flushed between repetitions, median;
GB/s on ALGORITHMIC bytes: dA_l + Z_l + X read, dA_{l-1} written (bf16), each once.

  python profiles/microbench_bwd_fused.py                      # table + JSON line
  python profiles/microbench_bwd_fused.py --only sa1.3 --reps 1   # one launch (for ncu)
"""
import argparse
import ctypes
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
# name: (rows, K, N, has_prev, want_dx, da_mode)
SHAPES = {
    "sa1.3": (B * 1024 * 32, 32, 64, 1, 1, 3), "sa1.2": (B * 1024 * 32, 32, 32, 1, 1, 1), "sa1.1": (B * 1024 * 32, 12, 32, 0, 0, 1),
    "sa2.3": (B * 256 * 32, 64, 128, 1, 1, 3), "sa2.2": (B * 256 * 32, 64, 64, 1, 1, 1), "sa2.1": (B * 256 * 32, 67, 64, 0, 1, 1),
    "sa3.2": (B * 64 * 32, 128, 128, 1, 1, 1),
    "fp1.3": (B * 4096, 128, 128, 1, 1, 1), "fp1.1": (B * 4096, 128, 128, 0, 1, 1), "head.conv1": (B * 4096, 128, 128, 1, 1, 0),
}


def ld(c):
    return (c + 7) // 8 * 8


def rows(M, C, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.zeros(M, ld(C), device=DEV, dtype=torch.bfloat16)
    x[:, :C] = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    lib = L.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV) if args.reps > 1 else None
    out = {}
    for name, (M, K, N, has_prev, want_dx, da_mode) in SHAPES.items():
        if args.only and name != args.only:
            continue
        assert lib.pn2_mlp_bwd_layer_supported(M, K, N, ld(K), ld(K) if want_dx else 0, da_mode, has_prev, want_dx, 1), name
        dA, Z, X = rows(M, N, 1), rows(M, N, 2), rows(M, K, 3)
        W = torch.randn(N, K, device=DEV) / K ** 0.5
        img = torch.empty(lib.pn2_linear_wpack_bytes(N, K), device=DEV, dtype=torch.uint8)
        vp, ci = ctypes.c_void_p * 1, ctypes.c_int * 1
        L.call("pn2_pack_weights", 1, vp(W.data_ptr()), ci(K), ci(N), ci(1), vp(img.data_ptr()), L.stream())
        coef = torch.rand(10, 128, device=DEV) + 0.5
        dX = torch.empty(M, ld(K), device=DEV, dtype=torch.bfloat16) if want_dx else None
        dW = torch.zeros(N, K, device=DEV)
        nbytes = lib.pn2_mlp_bwd_layer_scratch_bytes(M, K, N)
        dbg = int(os.environ.get("PN2_BWD_DBG", "0"))
        scratch = torch.zeros(max(nbytes, 24 * 16 * 8), device=DEV, dtype=torch.uint8)
        accum = torch.zeros(4 * 2 * 4096, device=DEV, dtype=torch.float64)
        ticket = torch.zeros(4, device=DEV, dtype=torch.int32)
        dgb = torch.zeros(2, 128, device=DEV)
        a = L.BwdLayer()
        a.dA, a.ldda, a.da_mode = dA.data_ptr(), dA.shape[1], da_mode
        a.Z, a.ldz = Z.data_ptr(), Z.shape[1]
        a.scale, a.shift, a.mean, a.invstd, a.dgamma, a.dbeta = (coef[i].data_ptr() for i in range(6))
        a.wpack_t = img.data_ptr() if want_dx else None
        a.X, a.ldx = X.data_ptr(), X.shape[1]
        if has_prev:
            a.prev_scale, a.prev_shift, a.prev_mean, a.prev_invstd = (coef[6 + i].data_ptr() for i in range(4))
            a.stat_accum, a.ticket, a.dgamma_prev, a.dbeta_prev = accum.data_ptr(), ticket.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr()
        a.dX, a.lddx = (dX.data_ptr(), dX.shape[1]) if want_dx else (None, 0)
        a.dW, a.scratch = dW.data_ptr(), (scratch.data_ptr() if (nbytes or dbg & 32) else None)
        a.M, a.K, a.N = M, K, N

        def fn():
            L.call("pn2_mlp_bwd_layer", ctypes.addressof(a), L.stream())

        for _ in range(2 if args.reps > 1 else 0):
            fn()
        ts = []
        for _ in range(args.reps):
            if flush is not None:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        ms = ts[len(ts) // 2]
        nbytes_alg = 2 * M * (ld(N) * (1 if da_mode == 3 else 2) + ld(K) + (ld(K) if want_dx else 0))
        out[name] = {"M": M, "K": K, "N": N, "us": round(ms * 1e3, 1), "GBps": round(nbytes_alg / ms / 1e6, 1)}
        print("%-11s M %8d K %4d N %4d prev %d dx %d mode %d: %7.1f us  %7.1f GB/s" % (name, M, K, N, has_prev, want_dx, da_mode, ms * 1e3, nbytes_alg / ms / 1e6))
    if int(os.environ.get("PN2_BWD_DBG", "0")) & 32:
        st = scratch[:24 * 16 * 8].view(torch.int64).view(24, 16).cpu()
        base = int(st[st > 0].min())
        names = ["P:top", "P:emptyOK", "T:top", "T:fullA", "T:T1done", "T:T2done", "T:bar", "T:acc_emptyOK", "T:issued", "E:top", "E:acc_full", "E:epi_done", "E:bar", "E:stored+stats", "E:store_read"]
        print("clock64 stamps of CTA 0 (cycles since the first stamp), one row per tile:")
        print("tile " + " ".join("%13s" % n for n in names))
        for t in range(24):
            print("%4d " % t + " ".join("%13d" % (int(st[t, i]) - base if int(st[t, i]) else -1) for i in range(15)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
