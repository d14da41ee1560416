"""Host-side drivers around the hot path: one training step and block-sharded inference.

``SemSegTrainer.step`` is the body of the reference's batch loop
(/root/reference/localfunctions.py:202-218: zero_grad, host->device copy, transpose,
forward, weighted NLL, backward, optimizer step) with the optimizer the reference builds
(/root/reference/sem_seg_training.py:576-582: Adam, betas (0.9, 0.999), eps 1e-8, weight
decay 1e-4).  Multi-GPU is pure data parallelism over point-cloud blocks, one process per
GPU: inference needs no collective; training all-reduces ONE flat fp32
gradient buffer (~3.9 MB) over NCCL/NVLink per step.  BatchNorm statistics stay per rank.
"""
import contextlib
import gc
import os

import torch
import torch.distributed as dist

from . import modules
from .sem_seg import get_loss, get_model


@contextlib.contextmanager
def _no_gc():
    """No cyclic garbage collection inside a CUDA-graph capture: collecting an object that owns pinned host memory makes
    the host allocator record events on the streams that used it -- an invalid operation on a capturing stream."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def _fps_sm_budget(batch_clouds):
    """{set-abstraction index: SMs} for the levels whose forward kernels run next to the FPS kernels of the NEXT batch in a
    pipelined graph.  FPS holds one CTA per cloud for ~0.45 ms (levels 1 and 2 of the index pipeline) and that CTA's shared
    memory keeps the SM from taking its full share of a persistent feature-path kernel, whose grid (sized for 148 SMs) then
    ends in a partial second wave: sa1's three forward layers took 192 us next to FPS, 129 us alone.  Sizing those grids for
    148 - clouds SMs keeps every CTA resident at once.  PN2_SA_SM_BUDGET=0 disables, =a,b,.. picks the levels."""
    env = os.environ.get("PN2_SA_SM_BUDGET", "0,1")
    if env in ("", "0") or not (0 < batch_clouds <= 74):
        return None
    sms = int(os.environ.get("PN2_SA_SM_BUDGET_SMS", "0")) or 148 - int(batch_clouds)
    return {int(v): sms for v in env.split(",")}


def _capture_stream(dev):
    """The stream a step / forward graph is captured on: HIGH priority, so its kernel nodes win the block scheduler
    against the index-pipeline branch (captured from default-priority streams).  Without it the big-grid ball-query and
    3-NN kernels of the next batch, once launched, keep every free block slot until their last wave and the feature
    path's next kernel waits ~0.1 ms behind them although it is the critical path.  PN2_MAIN_PRIORITY=0 switches it off."""
    prio = int(os.environ.get("PN2_MAIN_PRIORITY", "-1"))
    return torch.cuda.Stream(device=dev, priority=prio)


def shard_range(n_items, rank, world):
    """Contiguous block partition: rank r owns [r*ceil(n/world), min(n, (r+1)*ceil(n/world)))."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class _HostStager:
    """Two device staging slots fed from (pinned) host memory over a dedicated copy stream, so the host->device copy of the
    batch handed in by call i runs while the GPU is still busy with the replay launched by call i-1.  put() enqueues the
    copies of this call's batch and returns its slot; the slot filled by the previous call is then consumed on the
    compute stream (take -> use -> release).  One more stage of software pipelining: results come back one call later."""

    def __init__(self, templates, device):
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [[torch.empty_like(t) for t in templates] for _ in range(2)]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]
        self.rows = [None, None]          # leading-dimension sizes of the tensors staged in each slot, None = empty
        self.k = 0

    def put(self, *host_tensors):
        k, self.k = self.k, self.k ^ 1
        with torch.cuda.stream(self.stream):
            if self.consumed[k] is not None:                 # the compute stream has finished reading this slot
                self.stream.wait_event(self.consumed[k])
            for dst, src in zip(self.slots[k], host_tensors):
                dst[:src.shape[0]].copy_(src, non_blocking=True)
            self.copied[k].record(self.stream)
        self.rows[k] = [t.shape[0] for t in host_tensors]
        return k

    def take(self, k):
        """the staged tensors of slot k, readable on the current stream"""
        torch.cuda.current_stream(self.slots[k][0].device).wait_event(self.copied[k])
        return [t[:n] for t, n in zip(self.slots[k], self.rows[k])]

    def release(self, k):
        ev = self.consumed[k] = self.consumed[k] or torch.cuda.Event()
        ev.record()
        self.rows[k] = None

    def pending(self):
        """slots that hold a staged, not yet consumed batch, oldest first"""
        order = [self.k, self.k ^ 1]      # self.k is the slot the NEXT put() will use, i.e. the older one
        return [k for k in order if self.rows[k] is not None]


class _BnMomentum:
    """Every BatchNorm's momentum in ONE device tensor the fused finalize kernels read at run time
    (modules.set_momentum_buffers), kept in step with `bn.momentum` by sync(): the reference's per-epoch schedule
    (/root/reference/localfunctions.py:191-195, `m.momentum = momentum` on every BatchNorm1d/2d) reaches a captured
    training step without re-capture, the way the learning rate does through FlatAdam's device hyper-parameters."""

    def __init__(self, model, device):
        self.bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self.dev = torch.zeros(max(1, len(self.bns)), device=device, dtype=torch.float32)
        self.host = torch.zeros(max(1, len(self.bns)), dtype=torch.float32).pin_memory()
        self.seen = None
        modules.set_momentum_buffers({bn: self.dev[i:i + 1] for i, bn in enumerate(self.bns)})
        self.sync()

    def sync(self):
        now = tuple(-1.0 if bn.momentum is None else float(bn.momentum) for bn in self.bns)
        if now != self.seen:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("BatchNorm momentum changed during CUDA-graph capture")
            torch.cuda.synchronize(self.dev.device)          # rare (once per schedule step); the last copy may be in flight
            self.host[:len(now)].copy_(torch.tensor(now, dtype=torch.float32))
            self.dev.copy_(self.host, non_blocking=True)
            self.seen = now


class FlatGradients:
    """One flat fp32 buffer holding every parameter gradient, so data-parallel training needs a single
    all-reduce per step (no bucketing: the buffer is latency-, not bandwidth-bound).  The buffer is also the
    gradient SINK of the hot path (modules.set_grad_sink): backward kernels write weight / BatchNorm gradients
    straight into it and autograd adopts views of it as `.grad` -- no accumulate or fill kernels per parameter.
    Gradients produced elsewhere (the PyTorch head) are copied in by adopt()."""

    def __init__(self, params):
        from . import modules
        params = list(params)
        if params and isinstance(params[0], tuple):          # named_parameters(): names enable bucket_of()
            self.names = [n for n, p in params if p.requires_grad]
            params = [p for _, p in params]
        else:
            self.names = ["p%d" % i for i, p in enumerate(params) if p.requires_grad]
        self.check_next, self.last_check = False, None
        self.params = [p for p in params if p.requires_grad]
        # every parameter's slice starts on a 16-byte boundary (the weight-gradient kernels add into it with
        # red.global.add.v4.f32); the few padding elements stay zero
        offsets, total = [], 0
        for p in self.params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.offsets = offsets
        self.flat = torch.zeros(total, device=self.params[0].device, dtype=torch.float32)
        self.views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, offsets)]
        if self.flat.is_cuda:
            modules.set_grad_sink(self.params, self.views)

    def zero(self):
        for p in self.params:
            p.grad = None              # autograd adopts the incoming gradient tensor instead of adding into an old one
        self.flat.zero_()

    def adopt(self):
        """After backward: make every .grad a view of the flat buffer (copy the few that were produced elsewhere)."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                p.grad = v             # no gradient this step: the zeroed slice
            elif g.data_ptr() != v.data_ptr() or g.stride() != v.stride():
                v.copy_(g)
                p.grad = v

    def bucket_of(self, names):
        """[lo, hi) of the flat buffer covered by the parameters whose qualified name starts with one of `names`
        (they must be contiguous in parameter order); used to all-reduce a finished part of the gradient early."""
        idx = [i for i, n in enumerate(self.names) if any(n == m or n.startswith(m + ".") for m in names)]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError("bucket %s is not a contiguous run of parameters" % (names,))
        hi = self.offsets[idx[-1]] + (self.params[idx[-1]].numel() + 3) // 4 * 4
        return self.offsets[idx[0]], hi

    def all_reduce_mean(self, group=None, lo=0, hi=None):
        """Mean over the ranks of flat[lo:hi) (default: the whole buffer): ONE collective, averaged inside NCCL
        (ReduceOp.AVG -- no separate scale kernel); gloo (CPU tests) sums and scales.  With `check_next` set, the next
        whole-buffer call also all-gathers the ranks' LOCAL gradients and records how far the reduced buffer is from their
        mean (`last_check`) -- SURVEY.md 8(e)'s contract, reported by bench.py at N > 1."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        world = dist.get_world_size(group)
        part = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        check = self.check_next and lo == 0 and hi is None and not (self.flat.is_cuda and torch.cuda.is_current_stream_capturing())
        local = self.flat.clone() if check else None
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(part, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
            part.div_(world)
        if check:
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local, group=group)
            mean = torch.stack(gathered).double().mean(0)
            err = float((self.flat.double() - mean).abs().max())
            self.last_check = {"world": world, "max_abs_err": err, "max_abs_mean": float(mean.abs().max()),
                               "rel_l2": float((self.flat.double() - mean).norm() / mean.norm().clamp_min(1e-30)),
                               "rank_gradients_differ": bool(world > 1 and not torch.equal(gathered[0], gathered[-1]))}
            self.check_next = False


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam (the reference's optimizer, /root/reference/sem_seg_training.py:576-582; amsgrad / maximize off)
    as ONE kernel launch over a FlatGradients buffer (csrc/optim.cu: pn2_adam_step) instead of the six multi-tensor
    launches of the fused PyTorch implementation.  The moments are two flat buffers laid out like the gradients;
    ``state[p]`` holds views of them plus the shared step counter, so ``state_dict()`` has torch.optim.Adam's layout and
    ``load_state_dict()`` accepts one of its checkpoints (localfunctions.py:322 saves optimizer.state_dict()).
    Learning rate, betas, eps and weight decay live in DEVICE memory: ``param_groups[0]`` is re-read on every eager step
    and by ``sync_hyper()`` (call it before replaying a captured graph after a schedule changed them), so a captured
    step follows the reference's per-epoch learning-rate decay without re-capture.  One parameter group; CUDA fp32."""

    CHUNK = 1024          # elements per CTA

    def __init__(self, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not grads.flat.is_cuda:
            raise ValueError("FlatAdam needs CUDA parameters (pn2-b200 has no CPU path)")
        super().__init__(grads.params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.grads = grads
        dev = grads.flat.device
        for p in grads.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise TypeError("FlatAdam: parameters must be contiguous fp32 tensors on %s" % dev)
        self.exp_avg = torch.zeros_like(grads.flat)
        self.exp_avg_sq = torch.zeros_like(grads.flat)
        self.step_count = torch.zeros((), device=dev, dtype=torch.float32)
        self._ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        self._hyper = torch.zeros(5, device=dev, dtype=torch.float64)
        self._hyper_host = torch.zeros(5, dtype=torch.float64).pin_memory()
        self._hyper_seen = None
        chunks = [(t, s) for t, p in enumerate(grads.params) for s in range(0, p.numel(), self.CHUNK)]
        self._n_chunks = len(chunks)
        self._chunks = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(dev)
        self._off = torch.tensor(grads.offsets, dtype=torch.int64).to(dev)
        self._n = torch.tensor([p.numel() for p in grads.params], dtype=torch.int64).to(dev)
        self._ptrs = torch.tensor([p.data_ptr() for p in grads.params], dtype=torch.int64).to(dev)
        self._ptrs_host = [p.data_ptr() for p in grads.params]
        self._bind_state()
        self.sync_hyper()

    def _bind_state(self):
        for p, off in zip(self.grads.params, self.grads.offsets):
            n = p.numel()
            self.state[p] = {"step": self.step_count, "exp_avg": self.exp_avg[off:off + n].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p)}

    def sync_hyper(self):
        """param_groups[0] -> the device copy the kernel reads (a 40-byte copy, only when something changed)."""
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam supports one parameter group")
        g = self.param_groups[0]
        now = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if now != self._hyper_seen:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FlatAdam: hyper-parameters changed during CUDA-graph capture; call sync_hyper() before")
            torch.cuda.synchronize(self._hyper.device)      # rare; the pinned source of the last change may still be in flight
            self._hyper_host.copy_(torch.tensor(now, dtype=torch.float64))
            self._hyper.copy_(self._hyper_host, non_blocking=True)
            self._hyper_seen = now

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise ValueError("FlatAdam.step takes no closure")
        from ._lib import call, ptr, stream
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
            if [p.data_ptr() for p in self.grads.params] != self._ptrs_host:
                raise RuntimeError("FlatAdam: a parameter was re-allocated (model.to()/load with assign?) after the optimizer was built")
        for p, v in zip(self.grads.params, self.grads.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                raise RuntimeError("FlatAdam: a .grad is not its slice of the flat buffer; call FlatGradients.adopt() after backward")
        call("pn2_adam_step", ptr(self._ptrs), ptr(self._off), ptr(self._n), ptr(self._chunks), self._n_chunks, self.CHUNK,
             ptr(self.grads.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self._hyper), ptr(self.step_count),
             ptr(self._ticket), stream())

    def zero_grad(self, set_to_none=True):
        self.grads.zero()

    def reset_state(self):
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count.zero_()

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam (or FlatAdam) state_dict: the moments are copied INTO the flat buffers."""
        super().load_state_dict(state_dict)
        loaded = dict(self.state)
        steps = [float(st["step"]) for st in loaded.values() if "step" in st]
        with torch.no_grad():
            self.reset_state()
            for p, off in zip(self.grads.params, self.grads.offsets):
                st = loaded.get(p)
                if st:
                    n = p.numel()
                    self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            if steps:
                if min(steps) != max(steps):
                    raise ValueError("FlatAdam keeps ONE step counter; the checkpoint has %s..%s" % (min(steps), max(steps)))
                self.step_count.fill_(steps[0])
        self.state.clear()
        self._bind_state()
        self._hyper_seen = None
        self.sync_hyper()


class SemSegTrainer:
    def __init__(self, num_classes=18, num_extra_features=3, lr=1e-3, weight_decay=1e-4, device="cuda",
                 class_weights=None, model=None, fused_optimizer=True, augment_rotate_z=False, flat_optimizer=True, fused_loss=True):
        """fused_loss: evaluate the loss through get_model.forward_loss (inside the head kernels when the fused head applies).
        flat_optimizer: Adam as one launch of this library over the flat gradient buffer (FlatAdam) instead of
        torch.optim.Adam.  augment_rotate_z: apply the training loop's augmentation (provider.rotate_point_cloud_z on points[:, :, :3],
        /root/reference/localfunctions.py:205) to every batch ON THE DEVICE, with the reference's numpy angle draws."""
        self.augment_rotate_z = bool(augment_rotate_z)
        self.overlap_allreduce = os.environ.get("PN2_OVERLAP_ALLREDUCE", "1") != "0"
        # where the pipelined graph forks the next batch's index pipeline: "start" (top of the step), "sa<l>_issued" (after that
        # level's forward kernels) or "fp<l>" (when the gradient of that level's output is complete); see enable_cuda_graph
        self.index_anchor = os.environ.get("PN2_INDEX_ANCHOR", "start")
        self.fused_loss = bool(fused_loss)
        self.prepack = True                  # every MLP's weight images in one launch per step (modules.prepack_mlps)
        self._rot_staging = None
        self.device = torch.device(device)
        on_gpu = self.device.type == "cuda"
        if on_gpu:
            # one GPU per process: the library launches on the CURRENT device's streams (include/pn2b200.h takes a stream, not
            # a device), so the trainer's device becomes the current one
            if self.device.index is not None:
                torch.cuda.set_device(self.device)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_classes = num_classes
        self.model = (model if model is not None else get_model(num_classes, num_extra_features)).to(self.device)
        self.criterion = get_loss()
        self.broadcast_state()
        self.grads = FlatGradients(self.model.named_parameters())
        self._bn_momentum = _BnMomentum(self.model, self.device) if on_gpu else None
        if on_gpu and flat_optimizer:
            self.optimizer = FlatAdam(self.grads, lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=weight_decay)
        else:
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-08,
                                              weight_decay=weight_decay, fused=bool(fused_optimizer and on_gpu),
                                              capturable=on_gpu)
        self.class_weights = (torch.ones(num_classes) if class_weights is None else class_weights).to(self.device)
        self._graph = None

    def broadcast_state(self):
        """Data-parallel replicas start (and, after a checkpoint load on rank 0, restart) from rank 0's parameters and
        BatchNorm buffers -- DDP's construction-time broadcast.  Afterwards the replicas only exchange gradients; the running
        statistics stay per rank and rank 0's are the ones a checkpoint should save.  No-op without an
        initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            with torch.no_grad():
                for t in self.model.state_dict().values():
                    dist.broadcast(t, src=0)

    def release_graphs(self):
        """Drop every captured graph (and what only they keep alive).  NCCL requires the graphs that captured a
        communicator's collectives to be destroyed before the communicator is."""
        self.__dict__.pop("_graphs", None)
        self.__dict__.pop("_g_losses", None)
        self._graph = self._g_loss = None
        self._slots = []
        gc.collect()

    def _early_buckets(self):
        """Data-parallel overlap: which parts of the flat gradient are final before backward ends.  Backward runs
        head -> fp1 .. fp4 -> sa4 .. sa1 and the library's kernels write every weight / BatchNorm gradient of those modules
        straight into the flat buffer (the gradient sink), so [fp4 .. fp1, head] is final when the gradient of sa4's output
        arrives and [sa3, sa4] when the gradient of sa2's output is complete; only [sa1, sa2] (80 KB) is left for the
        all-reduce on the critical path after the last backward kernel.  Returns {feature level: (lo, hi)} or None when
        some gradient of those buckets is produced outside the sink (PyTorch head, fp32 rows, custom modules)."""
        from . import ops
        m = self.model
        if not (self.overlap_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return None
        if not (self.fused_loss and hasattr(m, "forward_loss") and getattr(m, "fused_head", False) and self.device.type == "cuda"
                and ops.rows_dtype() == torch.bfloat16 and type(m).__name__ == "get_model" and hasattr(m, "feature_grad_hooks")):
            return None
        try:
            return {4: self.grads.bucket_of(["fp4", "fp3", "fp2", "fp1", "conv1", "bn1", "conv2"]),
                    2: self.grads.bucket_of(["sa3", "sa4"]), 0: self.grads.bucket_of(["sa1", "sa2"])}
        except ValueError:
            return None

    def _step_impl(self, points, target, geometry=None):
        if self._bn_momentum is not None and not torch.cuda.is_current_stream_capturing():
            self._bn_momentum.sync()
        self.grads.zero()
        if self.prepack and hasattr(self.model, "training_chains") and points.is_cuda:
            modules.prepack_mlps(self.model.training_chains())      # every MLP's weight images in one launch
        buckets = None if self.grads.check_next else self._early_buckets()
        if buckets is not None:
            main = torch.cuda.current_stream(self.device)
            comm = self._comm_stream = getattr(self, "_comm_stream", None) or torch.cuda.Stream(device=self.device)

            def reduce_when_ready(level):
                lo, hi = buckets[level]

                def hook(grad):          # fires when the gradient w.r.t. that level's features is complete
                    comm.wait_stream(main)
                    with torch.cuda.stream(comm):
                        self.grads.all_reduce_mean(lo=lo, hi=hi)
                    return grad
                return hook

            self.model.feature_grad_hooks = {4: reduce_when_ready(4), 2: reduce_when_ready(2)}
        index_hook, self._index_hook = getattr(self, "_index_hook", None), None
        if index_hook is not None and hasattr(self.model, "feature_grad_hooks"):
            self.model.feature_grad_hooks = dict(self.model.feature_grad_hooks or {}, **{self.index_anchor: index_hook})
        try:
            if self.fused_loss and hasattr(self.model, "forward_loss"):
                # forward + weighted NLL in one pass (the loss and its gradient come out of the head kernels when they apply)
                loss, _, _ = self.model.forward_loss(points.transpose(2, 1), target, self.class_weights, geometry=geometry)
            else:
                if geometry is None:
                    pred, feat = self.model(points.transpose(2, 1))
                else:
                    pred, feat = self.model(points.transpose(2, 1), geometry=geometry)
                loss = self.criterion(pred.contiguous().view(-1, self.num_classes), target, feat, self.class_weights)
        finally:
            if buckets is not None or index_hook is not None:
                self.model.feature_grad_hooks = None
        modules._STEP_IMAGES.clear()          # images a forward did not pick up must not outlive the parameters they were packed from
        loss.backward()
        self.grads.adopt()
        if buckets is not None:
            self.grads.all_reduce_mean(lo=buckets[0][0], hi=buckets[0][1])      # what is left: sa1 + sa2
            main.wait_stream(comm)
        else:
            self.grads.all_reduce_mean()
        self.optimizer.step()
        return loss.detach()

    def enable_cuda_graph(self, batch_clouds, npoint, channels, warmup=3, pipeline=False):
        """Capture the whole training step (forward, loss, backward, gradient all-reduce, Adam) into ONE
        CUDA graph replayed per step: the ~300 kernel launches of a step cost no host time any more.
        Inputs are copied into static buffers; the FPS start indices stay a fresh CPU-generator draw per
        step (drawn on the host before each replay into a pinned ring, one host->device copy ahead of the replay).
        The reference's per-epoch schedules need no re-capture with bf16 rows: FlatAdam reads the learning rate and the fused
        BatchNorm finalize reads each module's momentum from device memory (torch.optim.Adam / fp32 rows bake them in:
        re-capture after changing them).  Re-capturing keeps the optimizer state (moments, step counter) and the parameters;
        it refuses to run while pipelined batches are in flight (flush() first).

        pipeline=True software-pipelines consecutive batches inside the graph: a forked branch runs the
        coordinate-only index pipeline (FPS, ball query, 3-NN: get_model.geometry_all) of the batch just SUBMITTED
        while the main branch runs forward/backward/Adam of the batch submitted one call earlier with the indices
        computed for it during the previous replay.  The FPS dependency chain (0.5 ms on 32 of 148 SMs) thereby leaves
        the critical path.  Inputs and indices live in TWO slots and there are two captured graphs (sharing one memory
        pool): graph k trains on slot k while its index branch fills slot 1-k (ops.reuse_outputs), and consecutive calls
        alternate between them -- nothing is copied from a "next" to a "current" slot.
        step_device() then returns the loss of the PREVIOUS batch (None on the first call).  step() -- host buffers --
        adds two more stages: the host->device copy of the batch handed in runs on a copy stream beside the replay that
        works on the two batches before it, and the loss of a replay is read back by the NEXT call (the host never waits on
        the replay it has just launched), so it returns the loss of the batch handed in THREE calls earlier (None 3 times).
        flush() finishes what is in flight and returns the remaining losses in batch order.  Same arithmetic per batch
        as pipeline=False."""
        from . import ops
        from .modules import PointNetSetAbstraction
        dev = self.device
        if getattr(self, "_pipeline", False) and (self._primed or self._stager.pending() or any(self._loss_valid)):
            raise RuntimeError("enable_cuda_graph: batches are still in flight in the pipeline; call flush() first")
        rng_state = torch.get_rng_state()     # warm-up / capture must not advance the CPU generator the FPS start draws use
        self.model.train()
        self._sa = [m for m in self.model.modules() if isinstance(m, PointNetSetAbstraction) and not m.group_all]
        for m in self._sa:
            m.use_static_start_buffers(True)
        self._pipeline, self._primed, self._parity, self._start_group = bool(pipeline), False, 0, None
        # training hides the whole index branch behind ~3 ms of feature work: issuing it as ONE chain disturbs the feature
        # path least (measured: 3.02 vs 3.05 ms per step with the branch forked over helper streams); PN2_TRAIN_FORK=1 forks
        self._fork = os.environ.get("PN2_TRAIN_FORK", "0") != "0"
        n_slots = 2 if pipeline else 1
        self._pts = [torch.zeros(batch_clouds, npoint, channels, device=dev).uniform_(-0.5, 0.5) for _ in range(n_slots)]
        self._tgt = [torch.zeros(batch_clouds * npoint, dtype=torch.int64, device=dev) for _ in range(n_slots)]
        self._g_points, self._g_target = self._pts[0], self._tgt[0]       # the slot of the batch submitted last
        self._slots = [None] * n_slots           # pipeline: (geometry, its tensors in allocation order) per slot
        side = torch.cuda.Stream(device=dev)
        if pipeline:
            self._geo_stream = torch.cuda.Stream(device=dev)
            self._stager = _HostStager([self._pts[0], self._tgt[0]], dev)
            self._loss_host = torch.zeros(2).pin_memory()            # step(): losses come back through a pinned ring,
            self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]  # read one call after their replay was launched
            self._loss_valid, self._loss_slot = [False, False], 0
            with torch.no_grad():       # persistent index tensors of the two slots (outside any graph pool)
                for k in range(2):
                    with ops.record_outputs() as rec:
                        geo = self.model.geometry_all(self._pts[k].transpose(2, 1)[:, :3, :], fork=self._fork)
                    self._slots[k] = (geo, rec.tensors)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = [(p.detach().clone()) for p in self.model.state_dict().values()]
        # ... nor as optimizer steps: the moments and the step counter (a loaded checkpoint's, or those of the epochs trained
        # so far when the caller re-captures) come back exactly as they were
        saved_opt = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in self.optimizer.state.items()}
        geo0 = self._slots[0][0] if pipeline else None
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self._pts[0], self._tgt[0], geo0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # one pinned ring / one host->device copy for the start indices of all levels (their device buffers become views of
        # the group's: must happen before the capture)
        stagings = [m.start_staging for m in self._sa]
        if stagings and all(isinstance(st, ops.StartIndexStaging) for st in stagings):
            self._start_group = ops.StartIndexGroup(stagings)
        # the warm-up steps must not count as training: restore parameters/buffers and optimizer moments
        with torch.no_grad():
            for t, s in zip(self.model.state_dict().values(), saved):
                t.copy_(s)
            for p, st in self.optimizer.state.items():
                old = saved_opt.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and torch.is_tensor(old.get(k)):
                            v.copy_(old[k])
                        else:
                            v.zero_()              # state created by the warm-up itself: a zero state is a fresh state
        self._graphs, self._g_losses = [], []
        for k in range(n_slots):
            graph = torch.cuda.CUDAGraph()
            pool = {} if not self._graphs else {"pool": self._graphs[0].pool()}     # never replayed concurrently
            with _no_gc(), torch.cuda.graph(graph, stream=_capture_stream(dev), **pool):
                if pipeline:
                    main = torch.cuda.current_stream(dev)
                    forked = []

                    def index_branch(grad=None, k=k, main=main, forked=forked):
                        # the index pipeline of the batch waiting in the other slot, forked off the main stream HERE
                        if not forked:
                            forked.append(True)
                            self._geo_stream.wait_stream(main)
                            with torch.cuda.stream(self._geo_stream), torch.no_grad():
                                with ops.reuse_outputs(self._slots[1 - k][1]):
                                    self.model.geometry_all(self._pts[1 - k].transpose(2, 1)[:, :3, :], fork=self._fork)
                        return grad

                    # Where the branch forks.  Its two long FPS kernels pin one CTA per cloud to 32 SMs for ~0.45 ms, and a
                    # persistent 148-CTA kernel of the feature path that meets them runs as two waves (sa1's three forward
                    # layers: 192 us next to FPS, 129 us alone).  Forking later -- next to the few-CTA layers of fp2..fp4 /
                    # sa4 / sa3 -- was measured (B200, 32 x 4096, ms/step): start 2.570, sa1_issued 2.598, sa2_issued 2.563,
                    # sa3_issued 2.628, fp4 2.604, fp2 2.594: the 32 SM x 0.45 ms the branch takes cost the same wherever they
                    # land, so the default stays at the top of the step.
                    if self.index_anchor != "start" and hasattr(self.model, "feature_grad_hooks"):
                        self._index_hook = index_branch
                    else:
                        index_branch()
                    budget = _fps_sm_budget(batch_clouds) if self.index_anchor == "start" else None
                    if budget and hasattr(self.model, "sa_sm_budget"):
                        self.model.sa_sm_budget = budget
                    try:
                        loss = self._step_impl(self._pts[k], self._tgt[k], self._slots[k][0])
                    finally:
                        if budget and hasattr(self.model, "sa_sm_budget"):
                            self.model.sa_sm_budget = None
                    index_branch()          # (an anchor that never fired: fork here, the join below still orders it)
                    main.wait_stream(self._geo_stream)
                else:
                    loss = self._step_impl(self._pts[0], self._tgt[0])
            self._graphs.append(graph)
            self._g_losses.append(loss)
        with torch.no_grad():                         # capture does not execute, but keep the state pristine anyway
            for t, s in zip(self.model.state_dict().values(), saved):
                t.copy_(s)
        self._graph, self._g_loss = self._graphs[0], self._g_losses[0]
        torch.set_rng_state(rng_state)
        return self

    def _augment(self, points):
        """in-place z rotation of a device batch [B, N, C]; angles drawn like the reference does"""
        if self.augment_rotate_z:
            from . import ops
            if self._rot_staging is None or self._rot_staging.B != points.shape[0]:
                self._rot_staging = ops.RotationStaging(points.shape[0], self.device)
            ops.rotate_point_cloud_z_(points, ops.draw_rotation_angles(points.shape[0]), staging=self._rot_staging)
        return points

    def _submit(self, points, target):
        """pipeline mode: put the new batch into the free slot, run its index pipeline next to the previous batch's
        feature path (the graph that trains on the other slot)."""
        from . import ops
        k = self._parity
        self._parity = k ^ 1
        self._g_points, self._g_target = self._pts[k], self._tgt[k]
        self._pts[k].copy_(points, non_blocking=True)
        self._tgt[k].copy_(target.view(-1), non_blocking=True)
        self._augment(self._pts[k])
        if not self._primed:                          # first batch: only its index pipeline, eagerly
            with torch.no_grad(), ops.reuse_outputs(self._slots[k][1]):
                self.model.geometry_all(self._pts[k].transpose(2, 1)[:, :3, :], fork=self._fork)
            self._primed = True
            return None
        self._pre_replay()
        self._graphs[k ^ 1].replay()                  # trains on slot k^1 (the batch before), indexes slot k
        return self._g_losses[k ^ 1]

    def _pre_replay(self):
        """host work a replay depends on: the reference's per-forward FPS start draws (module order) and, for FlatAdam,
        the device copy of the optimizer's hyper-parameters (follows a learning-rate schedule without re-capture)"""
        if self._start_group is not None:
            self._start_group.draw()
        else:
            for m in self._sa:
                m.start_staging.draw()
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.sync_hyper()
        if self._bn_momentum is not None:
            self._bn_momentum.sync()

    def _take_loss(self, j):
        if not self._loss_valid[j]:
            return None
        self._loss_ev[j].synchronize()
        self._loss_valid[j] = False
        return float(self._loss_host[j])

    def flush(self):
        """pipeline mode: finish every batch still in flight (a loss whose read-back is queued, staged host copies, then the
        feature path of the last one, eagerly); returns their losses as floats in batch order ([] if nothing was pending)."""
        out = []
        if not getattr(self, "_pipeline", False):
            return out
        for j in (self._loss_slot, self._loss_slot ^ 1):         # older slot first
            v = self._take_loss(j)
            if v is not None:
                out.append(v)
        for k in self._stager.pending():
            pts, tgt = self._stager.take(k)
            loss = self._submit(pts, tgt)
            self._stager.release(k)
            if loss is not None:
                out.append(float(loss))
        if self._primed:
            self._primed = False
            k = self._parity ^ 1                      # the slot of the batch submitted last: indexed, not yet trained on
            out.append(float(self._step_impl(self._pts[k], self._tgt[k], self._slots[k][0])))
        return out

    def step_device(self, points, target):
        """points [B, N, C] (point-major, as the DataLoader yields it) and target [B*N], on the device."""
        self.model.train()
        if self._graph is None:
            return self._step_impl(self._augment(points.clone()) if self.augment_rotate_z else points, target)
        if self._pipeline:
            return self._submit(points, target)
        self._g_points.copy_(points, non_blocking=True)
        self._g_target.copy_(target.view(-1), non_blocking=True)
        self._augment(self._g_points)
        self._pre_replay()
        self._graph.replay()
        return self._g_loss

    def _check_labels(self, target_host):
        """F.nll_loss (the reference's criterion, pointnet2_sem_seg.py:47-48) raises on a label outside [0, classes) other
        than ignore_index = -100; the fused loss kernel would silently give such points weight 0 -- e.g. a wrong class
        mapping (class8 remap vs NUM_CLASSES) would train on a subset.  Host labels are checked here (two reductions over
        ~130 k int64, well inside the pipelined step's host slack)."""
        if not target_host.is_cuda and target_host.numel():
            lo, hi = int(target_host.min()), int(target_host.max())
            if hi >= self.num_classes or (lo < 0 and bool(((target_host < 0) & (target_host != -100)).any())):
                raise ValueError("labels must lie in [0, %d) or be -100 (ignored); got range [%d, %d]" % (self.num_classes, lo, hi))

    def step(self, points_host, target_host):
        """One training step from HOST buffers (pinned memory recommended); returns the loss as a float
        (a device->host read, like the reference's per-batch `seg_pred.cpu()`)."""
        self._check_labels(target_host)
        if self._graph is not None and self._pipeline:
            self.model.train()
            st = self._stager
            k = st.put(points_host, target_host.view(-1))          # this batch: host -> device on the copy stream
            ret = None
            if st.rows[k ^ 1] is not None:                         # the previous call's batch has (long) arrived
                pts, tgt = st.take(k ^ 1)
                loss = self._submit(pts, tgt)
                st.release(k ^ 1)
                if loss is not None:
                    # queue the read-back of the loss this replay will produce; hand out the one queued by the call before
                    # (its replay has finished meanwhile), so the host never waits on the replay it has just launched
                    j, self._loss_slot = self._loss_slot, self._loss_slot ^ 1
                    ret = self._take_loss(j ^ 1)
                    self._loss_host[j].copy_(loss, non_blocking=True)
                    self._loss_ev[j].record()
                    self._loss_valid[j] = True
            st.copied[k].synchronize()                             # the caller may reuse its host buffers
            return ret
        if self._graph is not None:                  # host -> static device buffers directly
            self.model.train()
            self._g_points.copy_(points_host, non_blocking=True)
            self._g_target.copy_(target_host.view(-1), non_blocking=True)
            self._augment(self._g_points)
            self._pre_replay()
            self._graph.replay()
            return float(self._g_loss)
        points = points_host.to(self.device, non_blocking=True).float()
        target = target_host.to(self.device, non_blocking=True).long().view(-1)
        return float(self.step_device(points, target))


class SemSegPredictor:
    """Inference forward of the SSG network on fixed-shape batches, captured ONCE as a CUDA graph
    (the ~100 kernel launches of a forward cost no host time per batch).  The body is the
    reference's test-time batch step (/root/reference/localfunctions.py:396-400: host->device copy,
    transpose, `classifier(torch_data)`, arg-max of the log-probabilities); the FPS start indices stay a
    fresh CPU-generator draw per forward (pointnet2_utils.py:75), staged through pinned buffers.

    The parameters are frozen at construction (folded BatchNorm / packed weights are computed once, not per forward): build a
    new predictor after changing them.

    pipeline=True: the graph additionally runs, on a forked branch, the coordinate-only index pipeline
    (get_model.geometry_all) of the batch just submitted while the main branch runs the feature path of the
    batch submitted one call earlier (see SemSegTrainer.enable_cuda_graph); use submit()/flush(), which hand
    back the labels of the PREVIOUS batch."""

    def __init__(self, model, batch_clouds, npoint, channels, device="cuda", warmup=2, pipeline=False):
        from .modules import PointNetSetAbstraction
        self.model = model.eval()
        self.device = dev = torch.device(device)
        self.batch = batch_clouds
        self.pipeline = bool(pipeline)
        rng_state = torch.get_rng_state()     # construction must not advance the CPU generator the FPS start draws use
        self._sa = [m for m in model.modules() if isinstance(m, PointNetSetAbstraction) and not m.group_all]
        for m in self._sa:
            m.use_static_start_buffers(True)
        from . import ops
        n_slots = 2 if self.pipeline else 1
        self._pts = [torch.zeros(batch_clouds, npoint, channels, device=dev).uniform_(-0.5, 0.5) for _ in range(n_slots)]
        self.points = self._pts[0]
        self._slots, self._pending, self._parity, self._start_group = [None] * n_slots, None, 0, None
        side = torch.cuda.Stream(device=dev)
        if self.pipeline:
            self._geo_stream = torch.cuda.Stream(device=dev)
            self._stager = _HostStager([self.points], dev)
            with torch.no_grad():       # persistent index tensors of the two slots (outside any graph pool)
                for k in range(2):
                    with ops.record_outputs() as rec:
                        geo = self.model.geometry_all(self._pts[k].transpose(2, 1)[:, :3, :])
                    self._slots[k] = (geo, rec.tensors)
        side.wait_stream(torch.cuda.current_stream(dev))
        # the parameters are FROZEN for this predictor: folded BatchNorm and packed weight images are computed once in the
        # warm-up pass below (modules.frozen_parameters) and the captured graphs read them -- ~60 tiny launches fewer per
        # forward.  Build a new predictor after changing the model's parameters or running statistics.
        with modules.frozen_parameters() as cache:
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(max(1, warmup)):
                    self._forward(0)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            stagings = [m.start_staging for m in self._sa]
            if all(isinstance(st, ops.StartIndexStaging) for st in stagings):
                self._start_group = ops.StartIndexGroup(stagings)      # one host->device copy for the four levels' start draws
            self._frozen_tensors = list(cache.values())          # keep what the graphs will read alive with the predictor
            self._graphs, self._log_probs, self._labels = [], [], []
            for k in range(n_slots):
                self._capture(dev, k)
        self.graph, self.log_probs, self.labels = self._graphs[0], self._log_probs[0], self._labels[0]
        self._host_labels = torch.empty(batch_clouds, npoint, dtype=torch.int64).pin_memory()
        if self.pipeline:     # host results travel through a two-slot pinned ring and are handed out one call after their replay
            self._rb_host = [self._host_labels, torch.empty_like(self._host_labels).pin_memory()]
            self._rb_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._rb_rows, self._rb_slot = [None, None], 0
        torch.set_rng_state(rng_state)

    def _capture(self, dev, k):
        """graph k: feature path of slot k; pipelined: beside it the index pipeline of the batch waiting in slot 1-k"""
        from . import ops
        graph = torch.cuda.CUDAGraph()
        pool = {} if not self._graphs else {"pool": self._graphs[0].pool()}         # never replayed concurrently
        with _no_gc(), torch.no_grad(), torch.cuda.graph(graph, stream=_capture_stream(dev), **pool):
            if self.pipeline:
                main = torch.cuda.current_stream(dev)
                self._geo_stream.wait_stream(main)
                with torch.cuda.stream(self._geo_stream), ops.reuse_outputs(self._slots[1 - k][1]):
                    self.model.geometry_all(self._pts[1 - k].transpose(2, 1)[:, :3, :])
            budget = _fps_sm_budget(self._pts[k].shape[0]) if self.pipeline else None
            if budget and hasattr(self.model, "sa_sm_budget"):
                self.model.sa_sm_budget = budget
            try:
                labels, log_probs, _ = self._forward(k)                        # [B, npoint], [B, npoint, classes]
            finally:
                if budget and hasattr(self.model, "sa_sm_budget"):
                    self.model.sa_sm_budget = None
            if self.pipeline:
                main.wait_stream(self._geo_stream)
        self._graphs.append(graph)
        self._log_probs.append(log_probs)
        self._labels.append(labels)

    def _forward(self, k):
        """(labels, log-probabilities, l4 features) of slot k; the labels come out of the head kernel when it is fused"""
        x = self._pts[k].transpose(2, 1)
        geometry = self._slots[k][0] if self.pipeline else None
        if hasattr(self.model, "forward_labels"):
            return self.model.forward_labels(x, geometry=geometry)
        pred, feat = self.model(x) if geometry is None else self.model(x, geometry=geometry)
        return pred.argmax(dim=2), pred, feat

    def _replay(self, k=0):
        if self._start_group is not None:
            self._start_group.draw()
        else:
            for m in self._sa:
                m.start_staging.draw()
        self._graphs[k].replay()
        self.log_probs, self.labels = self._log_probs[k], self._labels[k]      # outputs of the replay just launched

    def forward_device(self, points):
        """points [b <= batch, npoint, C] on the device -> (log_probs, labels) views of the static outputs."""
        if self.pipeline:
            raise RuntimeError("pipelined predictor: use submit() / flush()")
        b = points.shape[0]
        self.points[:b].copy_(points, non_blocking=True)
        self._replay()
        return self.log_probs[:b], self.labels[:b]

    def predict_host(self, points_host):
        """points [b <= batch, npoint, C] on the host (pinned recommended) -> labels [b, npoint] on the host."""
        if self.pipeline:
            raise RuntimeError("pipelined predictor: use submit() / flush()")
        b = points_host.shape[0]
        self.points[:b].copy_(points_host, non_blocking=True)
        self._replay()
        self._host_labels.copy_(self.labels, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_labels[:b]

    def submit(self, points, to_host=True):
        """pipeline mode.  Hand in batch `points` [b <= batch, npoint, C] and get back the labels of an EARLIER batch
        ([b', npoint]; on the host when to_host, else a view of the static device output that the next call rewrites), or
        None while the pipeline fills.  Device input: its index pipeline runs next to the feature path of the batch handed
        in one call earlier (whose labels come back when to_host is False).  Host input (pinned recommended): additionally
        its host->device copy runs on a copy stream beside that replay -- one more stage.  to_host=True: the labels of a
        replay are read back by the NEXT call -- one more stage (device in / host out: 2 calls late; host in / host out: 3)."""
        if not self.pipeline:
            raise RuntimeError("submit() needs pipeline=True")
        if points.is_cuda:
            return self._advance(points, to_host)
        st = self._stager
        k = st.put(points)
        out = None
        if st.rows[k ^ 1] is not None:
            (staged,) = st.take(k ^ 1)
            out = self._advance(staged, to_host)
            st.release(k ^ 1)
        return out

    def _advance(self, points, to_host):
        from . import ops
        b = points.shape[0]
        k = self._parity
        self._parity = k ^ 1
        self._pts[k][:b].copy_(points, non_blocking=True)
        prev, self._pending = self._pending, b
        if prev is None:                           # first batch: only its index pipeline, eagerly
            with torch.no_grad(), ops.reuse_outputs(self._slots[k][1]):
                self.model.geometry_all(self._pts[k].transpose(2, 1)[:, :3, :])
            return None
        self._replay(k ^ 1)                        # feature path of slot k^1 (the batch before), index pipeline of slot k
        return self._read(prev, to_host)

    def flush_one(self, to_host=True):
        """pipeline mode: finish ONE more batch in flight and return its labels (like submit()), None once the pipeline
        is empty."""
        while True:
            pend = self._stager.pending()
            if pend:
                (staged,) = self._stager.take(pend[0])
                r = self._advance(staged, to_host)
                self._stager.release(pend[0])
            elif self._pending is not None:
                prev, self._pending = self._pending, None
                k = self._parity                   # the last batch sits in slot k^1; the index branch re-runs on the stale
                self._parity = k ^ 1               # slot k: harmless
                self._replay(k ^ 1)
                r = self._read(prev, to_host)
            else:
                for j in (self._rb_slot, self._rb_slot ^ 1):          # queued read-backs, older first
                    r = self._take_readback(j)
                    if r is not None:
                        return r
                return None
            if r is not None:
                return r

    def flush(self, to_host=True):
        """pipeline mode: finish every batch still in flight; returns the list of their label tensors (copies) in batch
        order, [] if nothing was pending."""
        out = []
        while True:
            r = self.flush_one(to_host)
            if r is None:
                return out
            out.append(r.clone())

    def _take_readback(self, j):
        if self._rb_rows[j] is None:
            return None
        self._rb_ev[j].synchronize()
        b, self._rb_rows[j] = self._rb_rows[j], None
        return self._rb_host[j][:b]

    def _read(self, b, to_host):
        if not to_host:
            return self.labels[:b]
        if not self.pipeline:
            self._host_labels.copy_(self.labels, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return self._host_labels[:b]
        # pipelined: queue the read-back of the labels the replay just launched will produce and hand out the one queued by
        # the call before (long finished), so the host never waits on the replay it has just launched.  The returned view
        # stays valid until the call after next.  (Use one of to_host=True / False consistently on a predictor.)
        j, self._rb_slot = self._rb_slot, self._rb_slot ^ 1
        out = self._take_readback(j ^ 1)
        self._rb_host[j].copy_(self.labels, non_blocking=True)
        self._rb_ev[j].record()
        self._rb_rows[j] = b
        return out


@torch.no_grad()
def predict_blocks(model, blocks_host, batch_size=32, rank=0, world=1, device="cuda", use_graph=True, pipeline=True):
    """sem_seg_testing-style inference (num_votes=1) over [nb, 4096, C] blocks held on the host:
    this rank labels its contiguous shard of blocks; no collective is needed (eval-mode BatchNorm
    uses running statistics, so blocks are independent).  Returns (lo, hi, labels [hi-lo, 4096] on host)."""
    model.eval()
    lo, hi = shard_range(blocks_host.shape[0], rank, world)
    out = torch.empty(hi - lo, blocks_host.shape[1], dtype=torch.int64)
    predictor = None
    if use_graph and hi - lo >= batch_size:
        predictor = SemSegPredictor(model, batch_size, blocks_host.shape[1], blocks_host.shape[2], device,
                                    pipeline=pipeline and hi - lo >= 2 * batch_size)
    if predictor is not None and predictor.pipeline:
        # consecutive batches overlap inside the graph: labels come back one call late
        done = lo
        for s in range(lo, hi, batch_size):
            prev = predictor.submit(blocks_host[s:min(hi, s + batch_size)].float())
            if prev is not None:
                out[done - lo:done - lo + prev.shape[0]] = prev
                done += prev.shape[0]
        for last in predictor.flush():
            out[done - lo:done - lo + last.shape[0]] = last
            done += last.shape[0]
        return lo, hi, out
    for s in range(lo, hi, batch_size):
        e = min(hi, s + batch_size)
        if predictor is not None:         # a short tail batch rides in the same fixed-shape graph
            out[s - lo:e - lo] = predictor.predict_host(blocks_host[s:e].float())
        else:
            x = blocks_host[s:e].to(device, non_blocking=True).float().transpose(2, 1)
            pred, _ = model(x)
            out[s - lo:e - lo] = pred.argmax(dim=2).cpu()
    return lo, hi, out


@torch.no_grad()
def predict_scene(model, blocks_host, point_idx_host, weight_host, num_points, num_classes, batch_size=32, rank=0, world=1,
                  device="cuda", pipeline=True, vote_pool=None, merge=True, predictor=None):
    """Whole-scene inference with one vote per block slot -- the body of the reference's per-scene test loop
    (/root/reference/localfunctions.py:385-405): batches of `blocks_host` [nb, npoint, C] go through the network, the
    arg-max labels vote into a [num_points, num_classes] pool through `point_idx_host` [nb, npoint] with the sample
    weights `weight_host` [nb, npoint] (pairs with weight 0 or inf do not vote, :341), and the scene labels are the
    arg-max of the pool.  Differences from the reference are where the work happens, not what it computes: the labels
    never leave the device between the network and the pool (ops.add_vote replaces the Python double loop), this rank
    only handles its contiguous shard of the blocks, and with merge=True the ranks' pools are summed with ONE
    all-reduce (torch.distributed initialised, world > 1) before the arg-max.  Returns (labels [num_points] int64 on the
    host, the int32 device pool).  Pass `vote_pool` to keep accumulating over several votes (num_votes > 1), and a
    ready `predictor` (SemSegPredictor of the same batch shape) to reuse its captured graph across scenes."""
    from . import ops
    model.eval()
    dev = torch.device(device)
    lo, hi = shard_range(blocks_host.shape[0], rank, world)
    pool = vote_pool if vote_pool is not None else ops.new_vote_pool(num_points, num_classes, dev)
    npoint = blocks_host.shape[1]

    def vote(s, e, labels):
        ops.add_vote(pool, point_idx_host[s:e], labels, None if weight_host is None else weight_host[s:e])

    if predictor is not None:
        batch_size = predictor.batch
    elif hi - lo >= batch_size:
        predictor = SemSegPredictor(model, batch_size, npoint, blocks_host.shape[2], dev,
                                    pipeline=pipeline and hi - lo >= 2 * batch_size)
    if predictor is not None and predictor.pipeline:
        done = lo
        for s in range(lo, hi, batch_size):
            prev = predictor.submit(blocks_host[s:min(hi, s + batch_size)], to_host=False)
            if prev is not None:          # device labels of the batch before; voted before the next replay rewrites them
                vote(done, done + prev.shape[0], prev)
                done += prev.shape[0]
        while True:                       # drain: staged batch (if any), then the one in flight -- vote after each replay
            rest = predictor.flush_one(to_host=False)
            if rest is None:
                break
            vote(done, done + rest.shape[0], rest)
            done += rest.shape[0]
    else:
        for s in range(lo, hi, batch_size):
            e = min(hi, s + batch_size)
            if predictor is not None:
                predictor.points[:e - s].copy_(blocks_host[s:e], non_blocking=True)
                predictor._replay()
                vote(s, e, predictor.labels[:e - s])
            else:
                pred, _ = model(blocks_host[s:e].to(dev, non_blocking=True).float().transpose(2, 1))
                vote(s, e, pred.argmax(dim=2))
    if merge and world > 1 and dist.is_available() and dist.is_initialized():
        dist.all_reduce(pool, op=dist.ReduceOp.SUM)
    return ops.vote_argmax(pool).cpu(), pool
