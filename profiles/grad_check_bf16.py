"""How far can a bf16 train step's gradients be from the fp32 reference's?  Per-parameter cosine of one train step's
gradients against the torch-CPU oracle port in fp32, for

  ours-bf16   SemSegTrainer (bf16 rows, fused head/loss, flat gradients) -- the benched configuration
  ours-fp32   the same trainer with fp32 rows (FMA-pipe kernels)
  torch-bf16  the oracle port itself on the GPU under torch.autocast(bfloat16) (cuDNN / cuBLAS bf16, fp32 BatchNorm) --
              what a user of the reference gets from PyTorch's own mixed precision

with random labels (the gradient is what is left of 131 072 nearly cancelling per-point terms) and with labels that
depend on the input (height bands: a gradient with signal).  Usage: python profiles/grad_check_bf16.py [B N]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I                                   # noqa: E402
from oracle import pn2_oracle as O                    # noqa: E402

pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
DEV = "cuda"
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 4096)
NC, SEED = 18, 4321
torch.set_num_threads(os.cpu_count() or 8)


_sqdist = O.pairwise_sqdist


def _sqdist_fp32(src, dst):
    """geometry stays fp32 under autocast (indices are not what is being compared here)"""
    tf32, torch.backends.cuda.matmul.allow_tf32 = torch.backends.cuda.matmul.allow_tf32, False
    try:
        with torch.autocast("cuda", enabled=False):
            return _sqdist(src.float(), dst.float())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


O.pairwise_sqdist = _sqdist_fp32


def oracle_grads(net, x, y, dev="cpu", autocast=False):
    net.train()
    net.zero_grad()
    torch.manual_seed(SEED)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        pred, _ = net(x.to(dev).transpose(2, 1))
    loss = O.nll(pred.float().contiguous().view(-1, NC), y.to(dev), torch.ones(NC, device=dev))
    loss.backward()
    return {n: p.grad.detach().double().cpu().flatten() for n, p in net.named_parameters()}, float(loss)


def ours_grads(precision, state, x, y):
    pn2.set_precision(precision)
    tr = pn2.SemSegTrainer(NC, 3, device=DEV)
    tr.model.load_state_dict(state)
    tr.model.drop1.p = 0.0
    torch.manual_seed(SEED)
    loss = float(tr.step_device(x.to(DEV), y.to(DEV)))
    views = {id(p): v for p, v in zip(tr.grads.params, tr.grads.views)}
    return {n: views[id(p)].detach().double().cpu().flatten() for n, p in tr.model.named_parameters()}, loss


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a, b, dim=0))


def main():
    torch.manual_seed(1234)
    ref = O.OracleSemSeg(NC, 3)
    ref.drop1.p = 0.0
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    x = I.facade_batch(B, N, 9, 11)
    labels = {"random labels": I.labels(B, N, NC, 111),
              "height-band labels": (x[:, :, 2] / 3.0 * NC).long().clamp(0, NC - 1).reshape(-1)}
    print("# python profiles/grad_check_bf16.py %d %d  -- per-parameter gradient cosine vs the fp32 torch-CPU oracle port, one train step" % (B, N))
    for tag, y in labels.items():
        ref.load_state_dict(state)
        g32, l32 = oracle_grads(ref, x, y)
        rows = {}
        gpu_ref = O.OracleSemSeg(NC, 3).to(DEV)
        gpu_ref.load_state_dict(state)
        gpu_ref.drop1.p = 0.0
        rows["torch-bf16"], l_ac = oracle_grads(gpu_ref, x, y, DEV, autocast=True)
        rows["ours-bf16"], l_b = ours_grads("bf16", state, x, y)
        rows["ours-fp32"], l_f = ours_grads("fp32", state, x, y)
        print("\n== %s: loss oracle-fp32 %.6f | torch-bf16 %.6f | ours-bf16 %.6f | ours-fp32 %.6f" % (tag, l32, l_ac, l_b, l_f))
        print("%-28s %10s %12s %12s %12s" % ("parameter", "|g_ref|", "torch-bf16", "ours-bf16", "ours-fp32"))
        worst = {k: (2.0, "") for k in rows}
        for n in g32:
            if n.endswith("bias") and ("mlp_convs" in n or n == "conv1.bias"):
                continue
            c = {k: cos(v[n], g32[n]) for k, v in rows.items()}
            for k in rows:
                if c[k] < worst[k][0]:
                    worst[k] = (c[k], n)
            print("%-28s %10.3e %12.5f %12.5f %12.5f" % (n, float(g32[n].norm()), c["torch-bf16"], c["ours-bf16"], c["ours-fp32"]))
        print("worst: " + "; ".join("%s %.4f (%s)" % (k, v[0], v[1]) for k, v in worst.items()))
    pn2.set_precision("fp32")


if __name__ == "__main__":
    main()
