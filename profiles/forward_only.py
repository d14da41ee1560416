"""One eval-mode forward of the SSG network on a 32 x 4096 x 9 batch, run eagerly a few times -- the command the
ncu launch list of the inference path is taken from (python profiles/forward_only.py [B] [iterations])."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pn2.set_precision("bf16")
torch.manual_seed(0)
net = pn2.get_model(18, 3).cuda().eval()
x = I.facade_batch(B, 4096, 9, 3).cuda().transpose(2, 1)
with torch.no_grad():
    for _ in range(iters):
        labels, pred, _ = net.forward_labels(x)      # the test loop's forward + arg-max (localfunctions.py:398-400)
torch.cuda.synchronize()
print("ok", tuple(pred.shape))
