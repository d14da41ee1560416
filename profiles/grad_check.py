"""Per-parameter gradient comparison of one fp32 train step: ours vs the torch-CPU oracle in fp32 and in fp64 (evaluated on
the fp32 sampling / grouping indices) -- tells rounding noise (ours-vs-64 of the size of oracle32-vs-64) from real
differences.  Usage: python profiles/grad_check.py"""
import sys, os, importlib, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
from oracle import pn2_oracle as O
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
DEV = 'cuda'
x_h = I.facade_batch(2, 2048, 9, 3).transpose(2, 1); x = x_h.to(DEV)
target = I.labels(2, 2048, 18, 7); weights = torch.ones(18)
def grads(model, inp, tgt, w):
    model.train(); model.zero_grad()
    torch.manual_seed(72); pred, _ = model(inp)
    loss = torch.nn.functional.nll_loss(pred.contiguous().view(-1, 18), tgt, weight=w)
    loss.backward()
    return {n: p.grad.detach().double().cpu() for n, p in model.named_parameters()}, loss.item()
ref = I.randomize_module_(O.OracleSemSeg(18, 3), 61); ref.drop1.p = 0.0
g32, l32 = grads(ref, x_h, target, weights)
O.GEOMETRY_DTYPE = torch.float32
ref64 = I.randomize_module_(O.OracleSemSeg(18, 3), 61).double(); ref64.drop1.p = 0.0
g64, l64 = grads(ref64, x_h.double(), target, weights.double())
O.GEOMETRY_DTYPE = None
pn2.set_precision('fp32')
net = I.randomize_module_(pn2.get_model(18, 3), 61); net.drop1.p = 0.0; net = net.to(DEV)
gm, lm = grads(net, x, target.to(DEV), weights.to(DEV))
print('loss oracle32 %.7f oracle64 %.7f ours %.7f' % (l32, l64, lm))
def rel(a, b): return ((a-b).norm() / (b.norm() + 1e-30)).item()
def mx(a, b): return ((a-b).abs().max() / (b.abs().max() + 1e-30)).item()
for n in g64:
    if 'mlp_convs' in n and n.endswith('bias'): continue
    print('%-26s |g| %.3e  o32-vs-64 %.2e (max %.2e)  ours-vs-64 %.2e (max %.2e)  ours-vs-o32 %.2e' % (
        n, g64[n].norm(), rel(g32[n], g64[n]), mx(g32[n], g64[n]), rel(gm[n], g64[n]), mx(gm[n], g64[n]), rel(gm[n], g32[n])))
