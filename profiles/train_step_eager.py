"""A few EAGER training steps (no CUDA graph, stream overlap off) of the bench configuration -- the command the ncu
captures of the training path are taken from.  Usage: python profiles/train_step_eager.py [steps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
mods = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pn2.set_precision("bf16")
torch.manual_seed(1234)
trainer = pn2.SemSegTrainer(18, 3, device="cuda")
mods.OVERLAP_WGRAD = False
trainer.model.overlap_geometry = False
pts = I.facade_batch(32, 4096, 9, 11).cuda()
lab = I.labels(32, 4096, 18, 111).cuda()
for _ in range(steps):
    loss = trainer.step_device(pts, lab)
torch.cuda.synchronize()
print("ok loss %.4f, %d pn2 launches" % (float(loss), pn2.launch_count()))
