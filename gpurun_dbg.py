import sys, os, importlib, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import _inputs as I
from oracle import pn2_oracle as O
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
ops = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.ops")
DEV = 'cuda'
pn2.set_precision("bf16")
B, N, C, NC, P = 4, 1024, 9, 18, 3000
torch.manual_seed(3)
net = I.randomize_module_(pn2.get_model(NC, C - 6), 23).to(DEV).eval()
nb = 11
blocks = torch.cat([I.facade_batch(B, N, C, 900 + i) for i in range(3)])[:nb]
g = np.random.RandomState(1)
pidx = torch.from_numpy(g.randint(0, P, size=(nb, N)))
w = torch.from_numpy((g.rand(nb, N) > 0.1).astype(np.float64))
torch.manual_seed(8)
_, _, lab = pn2.predict_blocks(net, blocks, batch_size=B, device=DEV, pipeline=False)
want_pool = O.add_vote(np.zeros((P, NC)), pidx.numpy(), lab.numpy(), w.numpy())
orig = ops.add_vote
for pipeline in (False, True):
    seen = []
    def spy(pool, pi, pl, wt=None):
        seen.append((pi.clone(), pl.clone().cpu(), None if wt is None else wt.clone()))
        return orig(pool, pi, pl, wt)
    ops.add_vote = spy
    torch.manual_seed(8)
    labels, pool = pn2.predict_scene(net, blocks, pidx, w, P, NC, batch_size=B, device=DEV, pipeline=pipeline)
    ops.add_vote = orig
    got = pool.cpu().numpy()
    print("pipeline", pipeline, "votes", got.sum(), "want", want_pool.sum(), "diff entries", int((got != want_pool).sum()))
    cat = torch.cat([s[1] for s in seen])
    print("  label batches", [tuple(s[1].shape) for s in seen], "labels equal predict_blocks:", bool(torch.equal(cat, lab)),
          "mismatch", int((cat != lab).sum()))
    print("  idx equal", bool(torch.equal(torch.cat([s[0] for s in seen]), pidx)))
    again = O.add_vote(np.zeros((P, NC)), pidx.numpy(), cat.numpy(), w.numpy())
    print("  device pool == oracle on ITS labels:", bool(np.array_equal(got, again)))
