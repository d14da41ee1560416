"""Kernel timeline of one CUDA-graph replay of the training step (or the inference forward) through torch.profiler
(CUPTI): start, duration and stream of every kernel, written as CSV -- the overlap / gap picture ncu's serialised
launch list cannot give.  Usage: python profiles/trace_step.py [train|forward] out.csv [--no-pipeline] [--replays=N]"""
import importlib, json, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
mode = args[0] if len(args) > 0 else "train"
out = args[1] if len(args) > 1 else "gpurun_out/trace_%s.csv" % mode
pn2.set_precision("bf16")
torch.manual_seed(1234)
B, N, C = 32, 4096, 9
pts = I.facade_batch(B, N, C, 11).cuda()
lab = I.labels(B, N, 18, 111).cuda()
if mode == "train":
    trainer = pn2.SemSegTrainer(18, 3, device="cuda")
    trainer.enable_cuda_graph(B, N, C, pipeline="--no-pipeline" not in sys.argv)
    run = lambda: trainer.step_device(pts, lab)
else:
    net = pn2.get_model(18, 3).cuda().eval()
    pipe = "--no-pipeline" not in sys.argv
    pred = pn2.SemSegPredictor(net, B, N, C, "cuda", pipeline=pipe)
    run = (lambda: pred.submit(pts, to_host=False)) if pipe else (lambda: pred.forward_device(pts))
for _ in range(4):
    run()
torch.cuda.synchronize()
replays = int([a.split("=")[1] for a in sys.argv if a.startswith("--replays=")][0]) if any(a.startswith("--replays=") for a in sys.argv) else 1
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(replays):       # > 1: back-to-back replays show the gap between consecutive graphs
        run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = []
for e in evs:
    rows.append((e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "stream", -1) if hasattr(e, "stream") else -1, e.name))
rows.sort()
t0 = rows[0][0]
with open(out, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for s, d, st, n in rows:
        f.write("%.2f,%.2f,%s,\"%s\"\n" % (s - t0, d, st, n.replace('"', "'")[:110]))
span = max(s + d for s, d, _, _ in rows) - t0
busy = sum(d for _, d, _, _ in rows)
print(json.dumps({"mode": mode, "kernels": len(rows), "span_us": round(span, 1), "sum_kernel_us": round(busy, 1)}))
