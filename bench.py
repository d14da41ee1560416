#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline measurement: points/sec of one PointNet++ SSG sem_seg
TRAINING STEP (zero_grad + forward + weighted NLL + backward + Adam) on synthetic
32 x 4096-point, 9-channel batches per GPU (BASELINE.json configs[1]), bf16 rows.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the meaning of every key.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

B_PER_GPU, NPOINT, CHANNELS, NUM_CLASSES = 32, 4096, 9, 18
METRIC, UNIT = "points/sec (PointNet++ SSG sem_seg train step)", "points/s"
WORKLOAD = "sem_seg train step, synthetic facade blocks %dx%dx%dch per GPU, %d classes" % (B_PER_GPU, NPOINT, CHANNELS, NUM_CLASSES)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _esz(dtype_code):
    return 2 if dtype_code == 1 else 4


# Algorithmic work of one call of a C-ABI entry point, from its arguments (include/pn2b200.h): (HBM bytes, flop or
# pair-evaluations x 6, bound).  "Algorithmic" = every operand read once and every result written once, in the
# storage types the call was given -- the figure DESIGN.md section 4 states per kernel.
KERNEL_MODEL = {
    "pn2_farthest_point_sample": lambda a: (a[4] * (12 * a[5] + 20 * a[6]), 9.0 * a[4] * a[5] * a[6], "latency"),
    "pn2_query_ball_point": lambda a: (a[8] * (12 * a[9] + 12 * a[10] + 8 * a[10] * a[12]), 6.0 * a[8] * a[9] * a[10], "alu"),
    "pn2_query_ball_point_grid": lambda a: (a[8] * (12 * a[9] + 12 * a[10] + 8 * a[10] * a[13]), 6.0 * a[8] * a[9] * a[10], "alu"),
    "pn2_three_nn": lambda a: (a[8] * (12 * a[9] + 12 * a[10] + 36 * a[9]), 6.0 * a[8] * a[9] * a[10], "alu"),
    "pn2_group_points": lambda a: (a[10] * a[12] * a[13] * (8 + (3 + a[14]) * 4 + a[16] * _esz(a[17])), 0.0, "hbm"),
    "pn2_group_points_bwd": lambda a: (a[4] * a[6] * a[7] * (8 + a[1] * _esz(a[2]) + a[8] * 4), 0.0, "hbm"),
    "pn2_linear_fwd": lambda a: (a[7] * (a[1] * _esz(a[2]) + a[11] * _esz(a[12])), 2.0 * a[7] * a[8] * a[9], "hbm"),
    "pn2_linear_bwd_data": lambda a: (a[4] * (a[1] * _esz(a[2]) + a[8] * _esz(a[9])), 2.0 * a[4] * a[5] * a[6], "hbm"),
    "pn2_linear_bwd_weight": lambda a: (a[8] * (a[1] * _esz(a[2]) + a[4] * _esz(a[5])), 2.0 * a[8] * a[9] * a[10], "hbm"),
    # fused backward layer (csrc/bwd_fused.cu): dA_l (or dZ_l) + Z_l + X read, dA_{l-1} written, bf16 rows, each once
    # (args: M, K, N, da_mode, ldda, ldz, ldx, lddx, has dX, has X -- recorded by _lib.call from the host struct)
    "pn2_mlp_bwd_layer": lambda a: (a[0] * 2 * ((a[4] if a[3] != 2 else 0) + (a[5] if a[3] != 3 else 0) + (a[6] if a[9] else 0) + (a[7] if a[8] else 0)),
                                    2.0 * a[0] * a[1] * a[2] * (2 if a[8] else 1), "hbm"),
    "pn2_bn_relu_max": lambda a: (a[5] * a[6] * a[7] * _esz(a[2]) + a[5] * a[7] * 8, 0.0, "hbm"),
    "pn2_bn_relu": lambda a: (a[5] * a[6] * (_esz(a[2]) + 4), 0.0, "hbm"),
    "pn2_bn_relu_bwd_reduce": lambda a: (a[10] * a[11] * (_esz(a[2]) + _esz(a[5])), 0.0, "hbm"),
    "pn2_pool_bn_relu_bwd_reduce": lambda a: (a[9] * a[11] * (8 + _esz(a[4])), 0.0, "hbm"),
    "pn2_bn_relu_bwd_dz": lambda a: (a[12] * a[13] * (_esz(a[2]) + _esz(a[5]) + _esz(a[16])), 0.0, "hbm"),
    "pn2_pool_bn_relu_bwd_dz": lambda a: (a[11] * a[13] * 8 + a[11] * a[12] * a[13] * (_esz(a[4]) + _esz(a[16])), 0.0, "hbm"),
    "pn2_interp_concat": lambda a: (a[10] * a[11] * (a[13] * 4 + 3 * a[14] * 4 + 36 + a[16] * _esz(a[17])), 6.0 * a[10] * a[11] * a[14], "hbm"),
    "pn2_interp_bwd": lambda a: (a[5] * a[6] * (a[1] * _esz(a[2]) + 36 + 3 * a[9] * 4), 6.0 * a[5] * a[6] * a[9], "hbm"),
    "pn2_to_rows": lambda a: (a[4] * a[5] * a[6] * 8, 0.0, "hbm"),
    "pn2_rows_to_f32": lambda a: (a[3] * a[5] * (_esz(a[2]) + 4), 0.0, "hbm"),
    # head tail (csrc/head.cu): Z rows in, log-probabilities (+ the bf16 activation kept for backward) out / their gradients back
    "pn2_head_tail_fwd": lambda a: (a[6] * (a[1] * 2 + a[8] * 4 + (a[13] * 2 if a[12] else 0)), 2.0 * a[6] * a[7] * a[8], "hbm"),
    "pn2_head_tail_bwd": lambda a: (a[3] * (a[5] * 8 + a[9] * 2 + (a[11] * 2 if a[10] else 0)), 2.0 * a[3] * a[4] * a[5], "hbm"),
    # ... fused with the weighted NLL: targets in, no dense [M, NC] gradient
    "pn2_head_tail_loss_fwd": lambda a: (a[6] * (a[1] * 2 + a[8] * 4 + 8 + (a[15] * 2 if a[14] else 0)), 2.0 * a[6] * a[7] * a[8], "hbm"),
    "pn2_head_tail_loss_bwd": lambda a: (a[6] * (a[8] * 4 + 8 + a[12] * 2 + (a[14] * 2 if a[13] else 0)), 2.0 * a[6] * a[7] * a[8], "hbm"),
    # Adam over the flat gradient buffer (csrc/optim.cu): p, m, v read + written, g read (chunks x elements is an upper bound)
    "pn2_adam_step": lambda a: (a[4] * a[5] * 28, 0.0, "hbm"),
}
# the launch-saving variants move the same bytes as the entry points they replace (same leading argument layout)
for _alias, _base in (("pn2_bn_relu_max_keep", "pn2_bn_relu_max"), ("pn2_linear_fwd_prepacked", "pn2_linear_fwd"), ("pn2_linear_bwd_data_prepacked", "pn2_linear_bwd_data"),
                      ("pn2_linear_bwd_weight_accum", "pn2_linear_bwd_weight"),
                      ("pn2_bn_relu_bwd_reduce_finalize", "pn2_bn_relu_bwd_reduce"),
                      ("pn2_pool_bn_relu_bwd_reduce_finalize", "pn2_pool_bn_relu_bwd_reduce")):
    KERNEL_MODEL[_alias] = KERNEL_MODEL[_base]


def kernel_table(calls, steps, step_ms, hbm_gbs):
    """Aggregate the event-timed entry-point calls of `steps` eager training steps (kernels serialised: no stream
    overlap) into one row per entry point: launches and ms per step, algorithmic GB/s and its fraction of the HBM peak."""
    agg = {}
    for name, a, ms in calls:
        row = agg.setdefault(name, {"entry": name, "calls": 0, "ms": 0.0, "bytes": 0.0, "flop": 0.0, "bound": "-"})
        row["calls"] += 1
        row["ms"] += ms
        model = KERNEL_MODEL.get(name)
        if model:
            b, f, bound = model(a)
            row["bytes"] += b
            row["flop"] += f
            row["bound"] = bound
    total = sum(r["ms"] for r in agg.values()) or 1.0
    out = []
    for r in sorted(agg.values(), key=lambda r: -r["ms"]):
        gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0.0
        out.append({"entry": r["entry"], "calls_per_step": r["calls"] / steps, "ms_per_step": round(r["ms"] / steps, 4),
                    "share_of_kernel_time": round(r["ms"] / total, 4), "bound": r["bound"],
                    "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / hbm_gbs, 4),
                    "alg_TFLOPs": round(r["flop"] / (r["ms"] * 1e-3) / 1e12, 3) if r["ms"] > 0 else 0.0})
    return out


def forward_points_per_s(pn2, model, host, resident, flush, steps, warmup, dev, pipeline=False):
    """Eval-mode forward of one batch through SemSegPredictor (the whole forward replayed as one CUDA graph):
    (ms per batch with resident inputs, ms per batch from pinned host buffers with the labels read back).
    pipeline: consecutive batches overlap inside the graph (index pipeline of batch i+1 next to the feature path of
    batch i); every timed call still does one batch's index work and one batch's feature work."""
    was_training = model.training
    predictor = pn2.SemSegPredictor(model, B_PER_GPU, NPOINT, CHANNELS, dev, pipeline=pipeline)
    if pipeline:
        for i in range(warmup + 1):
            predictor.submit(resident[i % len(resident)][0], to_host=False)
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            predictor.submit(resident[i % len(resident)][0], to_host=False)
            ends[i].record()
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) / steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            predictor.submit(host[i % len(host)][0])
        e0.record()
        for i in range(steps):
            predictor.submit(host[i % len(host)][0])
        e1.record()
        torch.cuda.synchronize()
        predictor.flush()
        model.train(was_training)
        return torch.tensor([ms, e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    for i in range(warmup):
        predictor.forward_device(resident[i % len(resident)][0])
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        flush.zero_()
        starts[i].record()
        predictor.forward_device(resident[i % len(resident)][0])
        ends[i].record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    predictor.predict_host(host[0][0])
    e0.record()
    for i in range(steps):
        predictor.predict_host(host[i % len(host)][0])
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / steps
    model.train(was_training)
    return torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)


def synthetic_batches(n, seed):
    """n pinned host batches [B, NPOINT, C] float32 + labels [B*NPOINT] int64 (S2 "facade block")."""
    import _inputs as I
    out = []
    for i in range(n):
        pts = I.facade_batch(B_PER_GPU, NPOINT, CHANNELS, seed + i)
        lab = I.labels(B_PER_GPU, NPOINT, NUM_CLASSES, seed + 100 + i)
        if torch.cuda.is_available():
            pts, lab = pts.pin_memory(), lab.pin_memory()
        out.append((pts, lab))
    return out


def reference_model_module():
    """(module with get_model / get_loss, kind): the UNMODIFIED reference (models/pointnet2_sem_seg.py on its own
    models/pointnet2_utils.py) from /root/reference or the copy staged under oracle/_ref by build(); the oracle port only
    if neither exists."""
    from oracle import ref_env as E
    if E.reference_root() is not None:
        mod, _ = E.load_model_module(ours=False)
        return mod, "reference", "unmodified reference (%s)" % E.reference_kind()
    from oracle import pn2_oracle as O

    class _Port:
        get_model = staticmethod(lambda nc, e: O.OracleSemSeg(nc, e))
        get_loss = staticmethod(lambda: (lambda pred, target, feat, w: O.nll(pred, target, w)))
    return _Port, "port", "oracle port of the reference (no reference tree staged)"


def cpu_train_step_rate(batch_clouds, steps, warmup, threads, budget_s=150.0):
    """The reference's own train step (localfunctions.py:203-218: zero_grad, forward, weighted NLL, backward, Adam of
    sem_seg_training.py:576-582) on the host cores.  Stops early once `budget_s` of timed work is spent; returns
    (points/s from the median step, s/step, timed steps, warm-up steps, kind, description)."""
    import _inputs as I
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    mod, kind, what = reference_model_module()
    net = mod.get_model(NUM_CLASSES, CHANNELS - 6).train()
    crit = mod.get_loss()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=1e-4)
    w = torch.ones(NUM_CLASSES)
    x = I.facade_batch(batch_clouds, NPOINT, CHANNELS, 11)
    y = I.labels(batch_clouds, NPOINT, NUM_CLASSES, 111)
    def one():
        t = time.perf_counter()
        opt.zero_grad()
        pred, feat = net(x.transpose(2, 1))
        loss = crit(pred.contiguous().view(-1, NUM_CLASSES), y, feat, w)
        loss.backward()
        opt.step()
        return time.perf_counter() - t

    t_begin, warm_done, times = time.perf_counter(), 0, []
    while warm_done < warmup and time.perf_counter() - t_begin < budget_s / 3:       # the driver's --warmup, bounded in time
        one()
        warm_done += 1
    warmup = warm_done
    while len(times) < steps and sum(times) < budget_s:
        times.append(one())
    med = sorted(times)[len(times) // 2]
    return batch_clouds * NPOINT / med, med, len(times), warmup, kind, what


def cpu_forward_rate(batch_clouds, threads, reps=2):
    """BASELINE.json configs[0]: eval-mode forward of `batch_clouds` x 4096 x 9 on the host cores (the reference itself when
    staged).  Returns (points/s, s per forward, kind, description)."""
    import _inputs as I
    torch.set_num_threads(threads)
    mod, kind, what = reference_model_module()
    torch.manual_seed(0)
    net = mod.get_model(NUM_CLASSES, CHANNELS - 6).eval()
    x = I.facade_batch(batch_clouds, NPOINT, CHANNELS, 21).transpose(2, 1)
    best = None
    with torch.no_grad():
        net(x[:2])
        for _ in range(reps):
            t = time.perf_counter()
            net(x)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
    return batch_clouds * NPOINT / best, best, kind, what


def torch_gpu_baseline(dev, steps=3):
    """The reference's own code path with `.cuda()` (localfunctions.py:208: eager PyTorch, cuBLAS / cuDNN / cub library kernels,
    cuDNN's default TF32 convolutions) on THIS GPU: eval forward and train step at the headline shape -- SURVEY 2b's
    "PyTorch-on-B200 bar".  Returns a dict (ms per forward / step, points/s)."""
    import _inputs as I
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = False, True        # torch's defaults
    out = {}
    try:
        mod, kind, what = reference_model_module()
        torch.manual_seed(0)
        net = mod.get_model(NUM_CLASSES, CHANNELS - 6).to(dev)
        crit = mod.get_loss()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=1e-4)
        w = torch.ones(NUM_CLASSES, device=dev)
        x = I.facade_batch(B_PER_GPU, NPOINT, CHANNELS, 11).to(dev)
        y = I.labels(B_PER_GPU, NPOINT, NUM_CLASSES, 111).to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        def train_step():
            opt.zero_grad()
            pred, feat = net(x.transpose(2, 1))
            loss = crit(pred.contiguous().view(-1, NUM_CLASSES), y, feat, w)
            loss.backward()
            opt.step()

        def forward():
            with torch.no_grad():
                net(x.transpose(2, 1))

        net.train()
        ms_train = timed(train_step)
        net.eval()
        ms_fwd = timed(forward)
        pts = B_PER_GPU * NPOINT
        out = {"kind": kind, "what": what + ", eager on cuda, torch default precision (fp32 matmul, TF32 cuDNN convolutions), "
               "resident inputs, %d timed iterations after 1 warm-up" % steps,
               "train_ms_per_step": ms_train, "train_points_per_s": pts / (ms_train * 1e-3),
               "forward_ms_per_batch": ms_fwd, "forward_points_per_s": pts / (ms_fwd * 1e-3), "unit": UNIT}
        del net, opt
    except Exception as exc:
        out = {"unavailable": repr(exc)[:200]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
        torch.cuda.empty_cache()
    return out


def dropin_rates(pn2, dev, host, steps=5):
    """Drop-in mode as a user of the reference gets it: the reference's UNCHANGED get_model / get_loss on top of this repo's
    models/pointnet2_utils.py, torch.optim.Adam, the eager batch body of localfunctions.py:203-220 (zero_grad, host->device
    copy, transpose, forward, loss, backward, step, `seg_pred.cpu()`), no CUDA graph, no fused head, no pipeline -- every
    operator call pays the Python + ctypes path.  ms per step for fp32 rows (the default precision) and bf16 rows."""
    from oracle import ref_env as E
    out = {}
    before = pn2.get_precision()
    try:
        if E.reference_root() is not None:
            mod, _ = E.load_model_module(ours=True)
            what = "unmodified reference get_model/get_loss on this repo's operators"
        else:
            mod = pn2
            what = "this repo's sem_seg.get_model (reference tree not staged)"
        for precision in ("fp32", "bf16"):
            pn2.set_precision(precision)
            torch.manual_seed(0)
            net = mod.get_model(NUM_CLASSES, CHANNELS - 6).to(dev).train()
            crit = mod.get_loss()
            opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=1e-4)
            w = torch.ones(NUM_CLASSES, device=dev)

            def body(points, target):
                opt.zero_grad()
                points, target = points.float().to(dev, non_blocking=True), target.long().to(dev, non_blocking=True)
                pred, feat = net(points.transpose(2, 1))
                pred = pred.contiguous().view(-1, NUM_CLASSES)
                loss = crit(pred, target.view(-1), feat, w)
                loss.backward()
                opt.step()
                return pred.cpu().data.max(1)[1]

            launches0 = pn2.launch_count()
            for i in range(2):
                body(*host[i % len(host)])
            torch.cuda.synchronize()
            per_step = (pn2.launch_count() - launches0) // 2
            t = time.perf_counter()
            for i in range(steps):
                body(*host[i % len(host)])
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t) * 1e3 / steps
            out[precision] = {"ms_per_step": ms, "points_per_s": B_PER_GPU * NPOINT / (ms * 1e-3), "pn2_entry_calls_per_step": per_step}
            del net, opt
        out["what"] = what + "; eager, torch.optim.Adam, host batch in, seg_pred.cpu() out (localfunctions.py:203-220); wall clock over %d steps" % steps
    except Exception as exc:
        out["unavailable"] = repr(exc)[:200]
    finally:
        pn2.set_precision(before)
        torch.cuda.empty_cache()
    return out


def shutdown(world, *trainers):
    """Multi-rank exit.  NCCL requires every CUDA graph that captured a communicator's collectives to be destroyed
    before the communicator is (otherwise the destroy blocks for ever): drop the captured training step first, give
    the process-group teardown a bounded time, then leave without running further destructors."""
    if world <= 1:
        return
    import gc
    import torch.distributed as dist
    sys.stdout.flush()
    for trainer in trainers:
        if trainer is not None:
            trainer.release_graphs()         # every reference to the captured graphs (they hold NCCL nodes)
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(20.0)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_clouds = args.ref_sample_clouds
    rate, sec, done, warm, kind, what = cpu_train_step_rate(sample_clouds, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "%s on the host cores (torch CPU, fp32), %d of the batch's 32 clouds per step; "
                   "median step; at most 150 s of timed steps" % (what, sample_clouds)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d x %d-point clouds per step (train step: fwd+bwd+Adam)" % (sample_clouds, NPOINT)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# =================================================================================================
# The other configurations of BASELINE.json (run by hand, results under profiles/): same launcher, same JSON keys.
# =================================================================================================
def _rank_max(t, world):
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def run_fps_ball(args, rank, world, dev, pn2, lib_mod, barrier, sampler):
    """BASELINE.json configs[2]: FPS 65536 -> 16384 and ball query r = 0.1 / nsample = 32 on 64 unit-cube clouds (SURVEY
    8(d) "Cfg 3"), the 64 clouds sharded contiguously over the ranks (strong scaling, no collective).  A step = both
    kernels over the rank's clouds."""
    import _inputs as I
    B_total, N, S = 64, 65536, 16384
    lo, hi = pn2.shard_range(B_total, rank, world)
    cube_h = I.cube_xyz(B_total, N, 0)[lo:hi].contiguous().pin_memory()
    start = I.start_indices(B_total, N, 2)[lo:hi].to(dev)
    cube = cube_h.to(dev)

    def step():
        _, new_xyz = pn2.farthest_point_sample(cube, S, start=start, return_xyz=True)
        return pn2.query_ball_point(0.1, 32, cube, new_xyz)

    for _ in range(args.warmup):
        step()
    lib_mod.time_entry_point("*")
    step()
    calls = lib_mod.timed_calls()
    lib_mod.time_entry_point(None)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = _rank_max(torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64), world).item()
    # end to end: pinned host coordinates in, ball-query indices (int64, as the reference returns them) back on the host
    out_h = torch.empty(hi - lo, S, 32, dtype=torch.int64).pin_memory()
    barrier()
    e0.record()
    for _ in range(max(1, args.steps // 4)):
        cube.copy_(cube_h, non_blocking=True)
        out_h.copy_(step(), non_blocking=True)
    e1.record()
    barrier()
    e2e = _rank_max(torch.tensor([e0.elapsed_time(e1) / max(1, args.steps // 4)], device=dev, dtype=torch.float64), world).item()
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None
    pk, pk_kind = peaks()
    by = {n: (a, t) for n, a, t in calls}
    ball_entry = "pn2_query_ball_point_grid" if "pn2_query_ball_point_grid" in by else "pn2_query_ball_point"
    fps_ms, ball_ms = by["pn2_farthest_point_sample"][1], by[ball_entry][1]
    b = hi - lo
    ball_bytes = b * (12 * N + 12 * S + 8 * S * 32)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import c_oracle as C
        n_cpu = 4
        t = time.perf_counter()
        xs = cube_h[:n_cpu].numpy()
        idx = C.fps(xs, S, start[:n_cpu].cpu().numpy())
        nx = np.take_along_axis(xs, idx[:, :, None].repeat(3, 2), 1)
        C.ball_query(0.1, 32, xs, np.ascontiguousarray(nx))
        sec = time.perf_counter() - t
        cpu = {"value": n_cpu * N / sec, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": "C oracle (scalar, one core), the complete workload for %d of the 64 clouds: %.1f s" % (n_cpu, sec)}
    line = {
        "metric": "points/sec (FPS 65536->16384 + ball query r=0.1 k=32, 64 clouds)", "value": B_total * N / (ms * 1e-3),
        "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "kernel microbench: FPS 65536 -> 16384 + ball query r=0.1 nsample=32, 64 unit-cube clouds sharded over the ranks",
                   "clouds_per_rank": b, "l2": "inputs (50 MB) and outputs (268 MB) per rank exceed nothing on their own; each kernel streams its cloud once"},
        "e2e": {"value": B_total * N / (e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": b * N * 12, "d2h_bytes_per_step": b * S * 32 * 8,
                "ms_per_step": e2e},
        "gpu_launches": 2 * args.steps, "clocks": clocks,
        "roofline": {"kernel": ("bg_build_kernel + bg_query_kernel (cell grid)" if ball_entry.endswith("_grid") else "ball_query_kernel"), "bound": "hbm", "achieved": ball_bytes / (ball_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": ball_bytes / (ball_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                     "peak_kind": pk_kind + " (burst copy bandwidth)", "avg_launch_ms": ball_ms,
                     "note": "the int64 index output dominates the bytes; fp32-ALU view: %.2f Tpair/s brute-force basis. "
                             "fps_kernel: %.2f ms (%.3f us per iteration, latency chain, %d 8-CTA clusters at a time)"
                             % (b * S * N / (ball_ms * 1e-3) / 1e12, fps_ms, fps_ms * 1e3 / S, 16)},
        "cpu_baseline": cpu, "kernel_ms": {"fps": fps_ms, "ball_query": ball_ms},
    }
    return line


def run_facade_from_scene(args, rank, world, dev, pn2, barrier, sampler):
    """BASELINE.json configs[3] from the RAW scene: one synthetic facade (SURVEY 8(d) "Cfg 4": x ~ U(0,60), y ~ N(0,0.15),
    z ~ U(0,25) float64, labels U{0..17}, rgb U{0..255}, np.random.seed(0)) goes host -> device once, is sliced into
    4096-point blocks on the device (pn2.slice_scene: sem_seg_testing.py:182-254), this rank's shard of the blocks runs
    through the pipelined predictor with device votes (pn2.predict_scene), the ranks' pools are summed and the scene labels
    read back.  Every rank slices the whole scene with the same generator seed (identical blocks), then takes its shard."""
    import numpy as np
    P, batch = args.scene_points, args.batch
    np.random.seed(0)
    scene = np.stack([np.random.uniform(0, 60, P), np.random.normal(0, 0.15, P), np.random.uniform(0, 25, P)], axis=1)
    labels = np.random.randint(0, NUM_CLASSES, P).astype(np.int64)
    rgb = np.random.randint(0, 256, (3, P)).astype(np.float64)
    hist = np.histogram(labels, range(NUM_CLASSES + 1))[0].astype(np.float32)      # sem_seg_testing.py:172-180
    hist = hist / np.sum(hist)
    lw = torch.from_numpy(np.power(np.amax(hist) / hist, 1 / 3.0))
    h_scene, h_labels, h_rgb = torch.from_numpy(scene).pin_memory(), torch.from_numpy(labels).pin_memory(), torch.from_numpy(rgb).pin_memory()
    names = ["red", "blue", "green"][:CHANNELS - 6]
    torch.manual_seed(1234)
    net = pn2.get_model(NUM_CLASSES, CHANNELS - 6).to(dev).eval()
    pipeline = not args.no_pipeline
    predictor = pn2.SemSegPredictor(net, batch, NPOINT, CHANNELS, dev, pipeline=pipeline)

    def run(n_points):
        gen = torch.Generator(device=dev).manual_seed(5)
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        d_scene = h_scene[:n_points].to(dev, non_blocking=True)
        d_labels = h_labels[:n_points].to(dev, non_blocking=True)
        d_rgb = h_rgb[:len(names), :n_points].to(dev, non_blocking=True) if names else None
        data, lab, w, idx = pn2.slice_scene(d_scene, d_labels, d_rgb, names, lw, generator=gen)
        t0[1].record()
        lo, hi = pn2.shard_range(data.shape[0], rank, world)
        pool = pn2.new_vote_pool(n_points, NUM_CLASSES, dev)
        if hi > lo:
            pn2.predict_scene(net, data[lo:hi], idx[lo:hi], w[lo:hi], n_points, NUM_CLASSES, predictor=predictor, device=dev,
                              vote_pool=pool, merge=False)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(pool, op=dist.ReduceOp.SUM)
        out = pn2.vote_argmax(pool, torch.uint8).cpu()
        t0[2].record()
        torch.cuda.synchronize()
        return data.shape[0], t0[0].elapsed_time(t0[1]), t0[0].elapsed_time(t0[2]), int(pool.sum().item()), out

    for _ in range(max(1, args.warmup - 2)):
        run(min(P, 400_000))
    barrier()
    first = run(P)                 # first full-size pass: the caching allocator obtains its multi-GB blocks (cudaMalloc) here
    barrier()
    wall0 = time.perf_counter()
    nb, slice_ms, total_ms, votes_total, out = run(P)
    barrier()
    wall = time.perf_counter() - wall0
    ms = _rank_max(torch.tensor([total_ms], device=dev, dtype=torch.float64), world).item()
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None
    assert votes_total == nb * NPOINT, (votes_total, nb * NPOINT)
    lo, hi = pn2.shard_range(nb, 0, world)
    n_batches = (hi - lo + batch - 1) // batch
    value = nb * NPOINT / (ms * 1e-3)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import pn2_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sub = min(P, 500_000)                      # the reference scans the WHOLE scene once per grid cell: cost ~ cells x points
        t = time.perf_counter()
        np.random.seed(1)
        d, l, w_, i_, cells = O.slice_scene(scene[:sub].copy(), labels[:sub], [rgb[k][:sub] for k in range(len(names))], names,
                                            lw.numpy(), block_points=NPOINT)
        t_slice = time.perf_counter() - t
        ref = O.OracleSemSeg(NUM_CLASSES, CHANNELS - 6).eval()
        x = torch.Tensor(d[:16]).transpose(2, 1)
        with torch.no_grad():
            ref(x)
            t = time.perf_counter()
            pred, _ = ref(x)
            O.add_vote(np.zeros((sub, NUM_CLASSES)), i_[:16], pred.argmax(2).numpy(), w_[:16])
            t_fwd = time.perf_counter() - t
        est = t_slice * (P / sub) + t_fwd * (nb / 16.0)
        cpu = {"value": nb * NPOINT / est, "unit": "points/s", "cores": threads, "kind": "port",
               "sample": "oracle slicer on the first %d scene points (%.1f s, scaled x%.0f: the per-cell np.where scans are linear in the "
                         "scene size) + oracle eval forward and numpy votes of 16 blocks (%.2f s, scaled to %d blocks)" % (
                             sub, t_slice, P / sub, t_fwd, nb)}
    line = {
        "metric": "points/sec (whole-facade sliding-block inference, num_votes=1)", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": n_batches, "warmup": args.warmup, "ms_per_step": ms / n_batches, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "raw synthetic facade of %d points (float64 xyz, labels, rgb) -> device slicer (block 1.0, stride 0.5) -> %d block "
                               "slots x %d points x %d ch -> network (batch %d, pipelined graph) -> device votes -> scene labels on the host; "
                               "points/s counts block-slot points" % (P, nb, NPOINT, CHANNELS, batch),
                   "slice_ms": slice_ms, "total_ms": ms, "first_pass_total_ms": first[2], "blocks": nb, "blocks_per_rank": hi - lo,
                   "wall_s_rank0": wall},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": int(P * (24 + 8 + 8 * len(names)) / n_batches),
                "d2h_bytes_per_step": int(P / n_batches), "ms_per_step": ms / n_batches},
        "gpu_launches": pn2.launch_count(), "clocks": clocks, "roofline": None, "cpu_baseline": cpu,
    }
    return line


def run_facade(args, rank, world, dev, pn2, barrier, sampler):
    """BASELINE.json configs[3]: sem_seg_testing-style whole-facade inference (num_votes = 1): the block slots of a
    synthetic ~10 M-point facade (8544 blocks of 4096 points, SURVEY 8(d) "Cfg 4": ~3.5 block slots per point) sharded
    contiguously over the ranks, labelled batch by batch through the CUDA-graph predictor, voted into the per-point pool
    on the device (pn2.predict_scene: the reference's add_vote loop, localfunctions.py:336-343), the ranks' pools summed
    with one all-reduce and the scene labels (arg-max of the pool, :405) read back to the host.  The whole measurement is
    end to end (host blocks, point indices and sample weights in, host scene labels out); a step = one batch."""
    import _inputs as I
    import numpy as np
    nb_total, batch = args.blocks, args.batch
    lo, hi = pn2.shard_range(nb_total, rank, world)
    torch.manual_seed(1234)
    net = pn2.get_model(NUM_CLASSES, CHANNELS - 6).to(dev).eval()
    chunks = [I.facade_batch(min(256, hi - s), NPOINT, CHANNELS, 7000 + s) for s in range(lo, hi, 256)]
    blocks = torch.cat(chunks).pin_memory()            # only this rank's shard is materialised
    n_scene = int(nb_total * NPOINT / 3.5)
    g = np.random.RandomState(77)                      # same slot -> point map on every rank
    pidx_all = torch.from_numpy(g.randint(0, n_scene, size=(nb_total, NPOINT)))
    pidx = pidx_all[lo:hi].clone().pin_memory()
    smpw = torch.ones(hi - lo, NPOINT, dtype=torch.float32).pin_memory()
    pipeline = not args.no_pipeline
    predictor = pn2.SemSegPredictor(net, batch, NPOINT, CHANNELS, dev, pipeline=pipeline)
    pool = pn2.new_vote_pool(n_scene, NUM_CLASSES, dev)
    for i in range(args.warmup):                       # same code path on a few batches (votes discarded)
        pn2.predict_scene(net, blocks[:2 * batch], pidx[:2 * batch], smpw[:2 * batch], n_scene, NUM_CLASSES, predictor=predictor,
                          device=dev, merge=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    # rank r's shard is blocks[0 : hi-lo] locally: run it as a one-rank scene into this rank's pool, then merge the pools
    labels, pool = pn2.predict_scene(net, blocks[:hi - lo], pidx, smpw, n_scene, NUM_CLASSES, predictor=predictor, device=dev,
                                     vote_pool=pool, merge=False)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(pool, op=dist.ReduceOp.SUM)
        labels = pn2.vote_argmax(pool).cpu()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    n_batches = (hi - lo + batch - 1) // batch
    ms = _rank_max(torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64), world).item()
    clocks = sampler.stop() if sampler else None
    votes_total = int(pool.sum().item())
    if rank != 0:
        return None
    assert votes_total == nb_total * NPOINT, (votes_total, nb_total * NPOINT)       # every (slot, point) pair voted once
    pts = nb_total * NPOINT
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import pn2_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        ref = O.OracleSemSeg(NUM_CLASSES, CHANNELS - 6).eval()
        x = blocks[:16].float().transpose(2, 1)
        with torch.no_grad():
            ref(x)
            t = time.perf_counter()
            pred, _ = ref(x)
            lab = pred.argmax(2).numpy()
            O.add_vote(np.zeros((n_scene, NUM_CLASSES)), pidx[:16].numpy(), lab, smpw[:16].numpy())
            sec = time.perf_counter() - t
        cpu = {"value": 16 * NPOINT / sec, "unit": "points/s", "cores": threads, "kind": "port",
               "sample": "oracle port, eval forward of 16 blocks + numpy vote accumulation (%.2f s); the reference's own add_vote "
                         "is a Python double loop and far slower than the numpy restatement timed here" % sec}
    value = pts / (ms * 1e-3)
    line = {
        "metric": "points/sec (whole-facade sliding-block inference, num_votes=1)", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": n_batches, "warmup": args.warmup, "ms_per_step": ms / n_batches, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "sem_seg_testing-style inference, %d block slots x %d points x %d ch (synthetic ~10 M-point facade, %d scene "
                               "points), batch %d, blocks sharded over the ranks, votes accumulated on the device, pools merged with one "
                               "all-reduce, scene labels returned to the host" % (nb_total, NPOINT, CHANNELS, n_scene, batch),
                   "blocks_per_rank": hi - lo, "launch": "one CUDA graph per batch" + (
                       ", consecutive batches software-pipelined (index pipeline of batch i+1 beside the feature path of batch i)" if pipeline else ""),
                   "wall_s_rank0": wall},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": batch * NPOINT * (CHANNELS * 4 + 8 + 4),
                "d2h_bytes_per_step": int(n_scene * 8 / n_batches), "ms_per_step": ms / n_batches},
        "gpu_launches": pn2.launch_count(), "clocks": clocks, "roofline": None, "cpu_baseline": cpu,
    }
    return line


def _compact(line):
    """the few keys of another workload's full line that the headline line carries under `other_workloads`"""
    if line is None:
        return None
    keep = {k: line.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "dtype")}
    keep["workload"] = line["config"]["workload"]
    keep["e2e_value"] = (line.get("e2e") or {}).get("value")
    for k in ("kernel_ms",):
        if k in line:
            keep[k] = line[k]
    for k in ("slice_ms", "total_ms", "blocks"):
        if k in line["config"]:
            keep[k] = line["config"][k]
    return keep


def other_workloads(args, rank, world, dev, pn2, lib_mod, barrier):
    """BASELINE.json configs[2], [3] and [4] at the current N, in short form (the full lines: --workload fps_ball / facade
    --from-scene, --channels 6).  No CPU legs here; every rank takes part (the shards and the vote merge need all of them)."""
    import argparse as _ap
    out = {}
    sub = _ap.Namespace(**vars(args))
    sub.no_cpu_baseline, sub.steps, sub.warmup = True, 3, 3
    try:
        out["fps_ball"] = _compact(run_fps_ball(sub, rank, world, dev, pn2, lib_mod, barrier, None))
    except Exception as exc:
        out["fps_ball"] = {"unavailable": repr(exc)[:200]}
    torch.cuda.empty_cache()
    try:
        out["facade_from_scene"] = _compact(run_facade_from_scene(sub, rank, world, dev, pn2, barrier, None))
    except Exception as exc:
        out["facade_from_scene"] = {"unavailable": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def train_rate_other_channels(pn2, dev, world, channels, steps, warmup, barrier, rank):
    """The headline measurement for another input width (--RGB_OFF: 6 channels, BASELINE.json configs[4]): ms per step
    resident and from host buffers, through the same pipelined graph."""
    import _inputs as I
    torch.manual_seed(1234)
    trainer = pn2.SemSegTrainer(NUM_CLASSES, channels - 6, device=dev)
    host = []
    for i in range(4):
        pts = I.facade_batch(B_PER_GPU, NPOINT, channels, 1000 * rank + 51 + i).pin_memory()
        lab = I.labels(B_PER_GPU, NPOINT, NUM_CLASSES, 1000 * rank + 151 + i).pin_memory()
        host.append((pts, lab))
    resident = [(p.to(dev), t.to(dev)) for p, t in host]
    trainer.enable_cuda_graph(B_PER_GPU, NPOINT, channels, pipeline=True)
    for i in range(warmup + 1):
        trainer.step_device(*resident[i % 4])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        trainer.step_device(*resident[i % 4])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    trainer.flush()
    for i in range(4):
        trainer.step(*host[i % 4])
    barrier()
    e0.record()
    for i in range(steps):
        trainer.step(*host[i % 4])
    e1.record()
    barrier()
    e2e = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    trainer.flush()
    _rank_max(ms, world), _rank_max(e2e, world)
    pts = world * B_PER_GPU * NPOINT
    return trainer, {"workload": "sem_seg train step, %dx%dx%dch per GPU (--RGB_OFF), %d classes" % (B_PER_GPU, NPOINT, channels, NUM_CLASSES),
                     "n_gpus": world, "ms_per_step": ms.item(), "value": pts / (ms.item() * 1e-3), "e2e_ms_per_step": e2e.item(),
                     "e2e_value": pts / (e2e.item() * 1e-3), "unit": UNIT, "scaling": "weak", "dtype": "bf16",
                     "note": "L2 not flushed between these steps (short form); graph + pipeline as the headline"}


def data_parallel_check(pn2, trainer, resident, world, rank, dev):
    """SURVEY.md 8(e) on the GPUs: (1) the all-reduced flat gradient equals the mean of the ranks' local gradients
    (all-gathered and averaged in fp64); (2) a rank's local gradient is what a single GPU computes: every rank runs RANK 0's
    batch with the collective switched off and the results are compared across the ranks (fp32 atomics order is the only
    freedom).  Eager steps; parameters, buffers and optimizer state are restored afterwards."""
    import torch.distributed as dist
    model, opt = trainer.model, trainer.optimizer
    saved = [t.detach().clone() for t in model.state_dict().values()]
    saved_opt = (opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_count.clone()) if hasattr(opt, "exp_avg") else None
    graph, trainer._graph = trainer._graph, None
    out = {}
    try:
        trainer.grads.check_next = True
        torch.manual_seed(77 + rank)
        trainer.step_device(*resident[0])
        torch.cuda.synchronize()
        out["allreduce_vs_mean_of_rank_gradients"] = trainer.grads.last_check
        # (2) rank 0's batch everywhere, no collective
        pts, tgt = resident[0][0].clone(), resident[0][1].clone()
        dist.broadcast(pts, src=0)
        dist.broadcast(tgt, src=0)
        with torch.no_grad():
            for t, sv in zip(model.state_dict().values(), saved):
                t.copy_(sv)
        overlap, trainer.overlap_allreduce = trainer.overlap_allreduce, False
        real = trainer.grads.all_reduce_mean
        trainer.grads.all_reduce_mean = lambda *a, **k: None
        try:
            torch.manual_seed(99)
            trainer.step_device(pts, tgt)
            torch.cuda.synchronize()
        finally:
            trainer.grads.all_reduce_mean = real
            trainer.overlap_allreduce = overlap
        local = trainer.grads.flat.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ref = gathered[0].double()
        worst = max(float((g.double() - ref).norm() / ref.norm().clamp_min(1e-30)) for g in gathered)
        out["rank_local_vs_rank0_on_the_same_batch"] = {"max_rel_l2": worst, "world": world,
                                                        "note": "no collective in this step; differences = fp32 atomic order only"}
    except Exception as exc:
        out["error"] = repr(exc)[:300]
    finally:
        with torch.no_grad():
            for t, sv in zip(model.state_dict().values(), saved):
                t.copy_(sv)
            if saved_opt is not None:
                opt.exp_avg.copy_(saved_opt[0]), opt.exp_avg_sq.copy_(saved_opt[1]), opt.step_count.copy_(saved_opt[2])
        trainer._graph = graph
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything a library writes to file descriptor 1 (NCCL prints its version banner there when NCCL_DEBUG=VERSION is
    in the environment) goes to stderr; the ONE JSON line of the contract is written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample-clouds", type=int, default=32,
                    help="clouds per step of the CPU reference arm (32 = the whole batch of the configuration; fewer = a bounded sample)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="do not overlap the index pipeline (FPS, ball query, 3-NN) of batch i+1 with the feature path of batch i")
    ap.add_argument("--workload", default="train", choices=["train", "fps_ball", "facade"],
                    help="train: BASELINE.json configs[1] (the headline; --channels 6 gives configs[4], the --RGB_OFF data-parallel run); "
                         "fps_ball: configs[2]; facade: configs[3]")
    ap.add_argument("--channels", type=int, default=9, choices=[6, 9], help="input channels: 9 = xyz+norm-xyz+RGB, 6 = --RGB_OFF")
    ap.add_argument("--blocks", type=int, default=8544, help="facade workload: block slots of the synthetic facade")
    ap.add_argument("--from-scene", action="store_true",
                    help="facade workload: start from a raw synthetic 10 M-point scene (SURVEY 8(d) Cfg 4) and slice it into blocks on "
                         "the device (pn2.slice_scene) inside the timed region, instead of starting from ready host blocks")
    ap.add_argument("--scene-points", type=int, default=10_000_000, help="facade --from-scene: points of the synthetic scene")
    ap.add_argument("--batch", type=int, default=128, help="facade workload: blocks per forward")
    ap.add_argument("--headline-only", action="store_true",
                    help="train workload: skip the extra legs (other_workloads, torch_gpu_baseline, dropin) -- for profiling runs")
    args = ap.parse_args()
    global CHANNELS, WORKLOAD
    CHANNELS = args.channels
    WORKLOAD = "sem_seg train step, synthetic facade blocks %dx%dx%dch per GPU, %d classes" % (B_PER_GPU, NPOINT, CHANNELS, NUM_CLASSES)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
    lib_mod = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200._lib")
    pn2.load()                                        # fails loudly if the CUDA library is missing
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pn2.set_precision(args.precision)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if args.workload != "train":
        if args.workload == "fps_ball":
            line = run_fps_ball(args, rank, world, dev, pn2, lib_mod, barrier, sampler)
        elif args.from_scene:
            line = run_facade_from_scene(args, rank, world, dev, pn2, barrier, sampler)
        else:
            line = run_facade(args, rank, world, dev, pn2, barrier, sampler)
        if line is not None:
            emit(line)
        if world > 1:
            sys.stdout.flush()
            t = threading.Thread(target=dist.destroy_process_group, daemon=True)
            t.start()
            t.join(20.0)
            os._exit(0)
        return

    torch.manual_seed(1234)                           # identical initial weights on every rank
    trainer = pn2.SemSegTrainer(NUM_CLASSES, CHANNELS - 6, device=dev)
    n_batches = 4
    host = synthetic_batches(n_batches, 1000 * rank + 11)
    resident = [(p.to(dev), t.to(dev)) for p, t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- every kernel of the path, event-timed live on its launching stream in an eager pass with the stream
    #      overlap switched off (a graph replay cannot carry events around one node; overlapping kernels would
    #      time each other) -------------------------------------------------------------------------------
    modules_mod = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules")
    for i in range(2):
        trainer.step_device(*resident[i % n_batches])
    overlap_saved = (modules_mod.OVERLAP_WGRAD, trainer.model.overlap_geometry)
    modules_mod.OVERLAP_WGRAD, trainer.model.overlap_geometry = False, False
    trainer.step_device(*resident[0])
    lib_mod.time_entry_point("*")
    eager0 = pn2.launch_count()
    kernel_steps = 3
    for i in range(kernel_steps):
        trainer.step_device(*resident[i % n_batches])
    launches_per_step = (pn2.launch_count() - eager0) // kernel_steps
    timed_calls = lib_mod.timed_calls()
    lib_mod.time_entry_point(None)
    modules_mod.OVERLAP_WGRAD, trainer.model.overlap_geometry = overlap_saved

    graphed = False
    pipelined = not args.no_pipeline and not args.no_graph
    unpipelined_ms = None
    if not args.no_graph:
        try:
            if pipelined:
                # the same step with the batches NOT overlapped, for the record (short: it is not the headline)
                trainer.enable_cuda_graph(B_PER_GPU, NPOINT, CHANNELS)
                for i in range(args.warmup):
                    trainer.step_device(*resident[i % n_batches])
                barrier()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
                for i, (a, b) in enumerate(ev):
                    flush.zero_()
                    a.record()
                    trainer.step_device(*resident[i % n_batches])
                    b.record()
                barrier()
                unpipelined_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
            trainer.enable_cuda_graph(B_PER_GPU, NPOINT, CHANNELS, pipeline=pipelined)
            graphed = True
        except Exception as exc:                       # report, then measure the eager path instead
            print("cuda graph capture failed, running eagerly: %r" % (exc,), file=sys.stderr)
            trainer._graph = None
            pipelined = False

    # ---- device-resident arm ("value") -------------------------------------------------------
    for i in range(args.warmup + (1 if pipelined else 0)):      # the first pipelined call only primes the index slot
        trainer.step_device(*resident[i % n_batches])
    barrier()
    launches0 = pn2.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations (outside the events)
        starts[i].record()
        trainer.step_device(*resident[i % n_batches])
        ends[i].record()
    barrier()
    # a graph replay launches the captured kernels without passing through the library's host entry
    # points, so the count is taken from the eager pass (same kernels, same order) times the steps
    launches = launches_per_step * args.steps if graphed else pn2.launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)

    # ---- end-to-end arm ("e2e"): host buffers in, loss out, copies inside the timed region ------
    for i in range(4):                                 # (the host-input pipeline is four stages deep: copy, indices, features, loss read-back)
        trainer.step(*host[i % n_batches])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.step(*host[i % n_batches])
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # ---- inference forward (BASELINE.json's "fwd" half of the metric): eval-mode get_model forward under no_grad on
    #      the same 32 x 4096 x 9 batch, inputs resident, L2 flushed between iterations ---------------------------
    trainer.flush()
    fwd_ms = forward_points_per_s(pn2, trainer.model, host, resident, flush, args.steps, args.warmup, dev, pipeline=pipelined)
    barrier()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(fwd_ms, op=dist.ReduceOp.MAX)
        if unpipelined_ms is not None:
            dist.all_reduce(unpipelined_ms, op=dist.ReduceOp.MAX)
    # ---- the rest of the measurement contract, outside the timed regions above ---------------------------------------
    dp_check = data_parallel_check(pn2, trainer, resident, world, rank, dev) if world > 1 else None
    trainer6, other = None, None
    if not args.headline_only:
        other = other_workloads(args, rank, world, dev, pn2, lib_mod, barrier)
        pn2.set_precision(args.precision)
        try:
            trainer6, ch6 = train_rate_other_channels(pn2, dev, world, 6, 10, 3, barrier, rank)
        except Exception as exc:
            ch6 = {"unavailable": repr(exc)[:200]}
        if other is not None:
            other["train_6ch"] = ch6
    if rank != 0:
        return shutdown(world, trainer, trainer6)

    points_per_step = world * B_PER_GPU * NPOINT
    ms_per_step = total_ms.item() / args.steps
    value = points_per_step / (ms_per_step * 1e-3)
    e2e_value = points_per_step / (e2e_ms.item() / args.steps * 1e-3)
    pk, pk_kind = peaks()

    # ---- roofline: per entry point, and the dominant one as the contract's "roofline" object
    kernels = kernel_table(timed_calls, kernel_steps, ms_per_step, pk["hbm_gbs"])
    roof = None
    hbm_rows = [k for k in kernels if k["bound"] == "hbm"]
    if hbm_rows:
        top = hbm_rows[0]
        top_calls = [(a, ms) for n, a, ms in timed_calls if n == top["entry"]]
        alg = sum(KERNEL_MODEL[top["entry"]](a)[0] for a, _ in top_calls)
        dur = sum(ms for _, ms in top_calls) * 1e-3
        # DRAM traffic of the same entry point from the committed ncu pass (profiles/make_traffic.py), per call like `achieved`
        traffic, traffic_src = None, None
        try:
            tname = next(n for n in ("r02_dram_traffic.json", "r01_dram_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", n)))
            with open(os.path.join(ROOT, "profiles", tname)) as f:
                tj = json.load(f)
            if top["entry"] in tj:
                traffic = tj[top["entry"]]["dram_bytes_per_step"] / top["calls_per_step"]
                traffic_src = "profiles/%s: " % tname + tj.get("_source", "")
        except Exception:
            pass
        roof = {"kernel": top["entry"] + " (all %d launches of a step; the largest share of kernel time among the streaming kernels)" % round(top["calls_per_step"]),
                "bound": "hbm", "achieved": alg / dur / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": alg / dur / 1e9 / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": pk_kind + " (burst copy bandwidth)", "avg_launch_ms": dur * 1e3 / len(top_calls),
                "bytes_per_launch_avg": alg / len(top_calls), "share_of_step": top["ms_per_step"] / ms_per_step,
                "note": "achieved = sum of algorithmic bytes / sum of CUDA-event durations over the launches of %d eager steps; "
                        "the FPS dependency chain (pn2_farthest_point_sample) is latency-bound and listed under kernels" % kernel_steps}

    cpu = cpu_fwd = gpu_base = dropin = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, sec, done, warm, kind, what = cpu_train_step_rate(B_PER_GPU, 5, 1, threads, budget_s=25.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d timed train steps (after %d warm-up) on the whole %d x %d-point batch (%s, torch CPU fp32, fwd+bwd+Adam), "
                         "%.1f s/step, median" % (done, warm, B_PER_GPU, NPOINT, what, sec)}
        # BASELINE.json configs[0]: eval forward 16 x 4096 x 9 on the host cores
        rate_f, sec_f, kind_f, what_f = cpu_forward_rate(16, threads)
        cpu_fwd = {"value": rate_f, "unit": UNIT, "cores": threads, "kind": kind_f, "s_per_forward": sec_f,
                   "sample": "eval-mode forward of 16 x %d x %d ch under no_grad (%s), best of 2" % (NPOINT, CHANNELS, what_f)}
    if not args.headline_only and world == 1:
        gpu_base = torch_gpu_baseline(dev)
        dropin = dropin_rates(pn2, dev, host)

    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch_clouds": world * B_PER_GPU, "points_per_cloud": NPOINT,
                   "parallelism": "dp%d (blocks sharded, flat-gradient NCCL all-reduce)" % world,
                   "l2": "256 MiB buffer written between timed steps (L2 flush), outside the per-step events",
                   "optimizer": "Adam(lr 1e-3, wd 1e-4) inside the step",
                   "launch": ("whole step replayed as one CUDA graph" + (" (two captured graphs alternate over two input/index slots)" if pipelined else "")) if graphed else "eager launches",
                   "pipeline": ("depth 2: inside the graph the index pipeline (FPS, ball query, 3-NN) of the batch submitted by this call runs "
                                "beside forward/backward/Adam of the batch submitted by the previous call; every step does one batch of each; "
                                "the loss read back is the previous batch's; e2e adds two stages: the host->device copy of a batch runs on a copy "
                                "stream beside the replay before it, and a replay's loss is read back (4 bytes, every step) by the next call") if pipelined else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms.item() / args.steps},
        "forward": {"value": points_per_step / (fwd_ms[0].item() * 1e-3), "unit": UNIT, "ms_per_batch": fwd_ms[0].item(),
                    "e2e_value": points_per_step / (fwd_ms[1].item() * 1e-3), "e2e_ms_per_batch": fwd_ms[1].item(),
                    "d2h_bytes_per_batch": B_PER_GPU * NPOINT * 8,
                    "what": "eval-mode forward (no_grad) of the same batch shape, replayed as one CUDA graph; value: resident "
                            "inputs, e2e: pinned host points in, arg-max labels read back to the host"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
    }
    if cpu_fwd is not None:
        line["cpu_baseline_forward"] = cpu_fwd
    if gpu_base is not None:
        line["torch_gpu_baseline"] = gpu_base
        if "train_ms_per_step" in gpu_base:
            line["torch_gpu_baseline"]["speedup_train"] = gpu_base["train_ms_per_step"] / ms_per_step
            line["torch_gpu_baseline"]["speedup_forward"] = gpu_base["forward_ms_per_batch"] / fwd_ms[0].item()
    if dropin is not None:
        line["dropin"] = dropin
    if other is not None:
        line["other_workloads"] = other
    if dp_check is not None:
        line["data_parallel_check"] = dp_check
    if unpipelined_ms is not None:
        line["unpipelined"] = {"ms_per_step": unpipelined_ms.item(), "value": points_per_step / (unpipelined_ms.item() * 1e-3), "unit": UNIT,
                               "what": "the same graph-replayed train step without overlapping consecutive batches (resident inputs)"}
    emit(line)
    shutdown(world, trainer, trainer6)


if __name__ == "__main__":
    main()
