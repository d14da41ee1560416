"""Host-side drivers around the hot path: one training step and block-sharded inference.

``SemSegTrainer.step`` is the body of the reference's batch loop
(/root/reference/localfunctions.py:202-218: zero_grad, host->device copy, transpose,
forward, weighted NLL, backward, optimizer step) with the optimizer the reference builds
(/root/reference/sem_seg_training.py:576-582: Adam, betas (0.9, 0.999), eps 1e-8, weight
decay 1e-4).  Multi-GPU is pure data parallelism over point-cloud blocks, one process per
GPU: inference needs no collective; training all-reduces ONE flat fp32
gradient buffer (~3.9 MB) over NCCL/NVLink per step.  BatchNorm statistics stay per rank.
"""
import torch
import torch.distributed as dist

from .sem_seg import get_loss, get_model


def shard_range(n_items, rank, world):
    """Contiguous block partition: rank r owns [r*ceil(n/world), min(n, (r+1)*ceil(n/world)))."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class FlatGradients:
    """One flat fp32 buffer holding every parameter gradient, so data-parallel training needs a single
    all-reduce per step (no bucketing: the buffer is latency-, not bandwidth-bound).  The buffer is also the
    gradient SINK of the hot path (modules.set_grad_sink): backward kernels write weight / BatchNorm gradients
    straight into it and autograd adopts views of it as `.grad` -- no accumulate or fill kernels per parameter.
    Gradients produced elsewhere (the PyTorch head) are copied in by adopt()."""

    def __init__(self, params):
        from . import modules
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=self.params[0].device, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        if self.flat.is_cuda:
            modules.set_grad_sink({id(p): v for p, v in zip(self.params, self.views)})

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

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class SemSegTrainer:
    def __init__(self, num_classes=18, num_extra_features=3, lr=1e-3, weight_decay=1e-4, device="cuda",
                 class_weights=None, model=None, fused_optimizer=True):
        self.device = torch.device(device)
        self.num_classes = num_classes
        self.model = (model if model is not None else get_model(num_classes, num_extra_features)).to(self.device)
        self.criterion = get_loss()
        self.grads = FlatGradients(self.model.parameters())
        on_gpu = self.device.type == "cuda"
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-08,
                                          weight_decay=weight_decay, fused=bool(fused_optimizer and on_gpu),
                                          capturable=on_gpu)
        self.class_weights = (torch.ones(num_classes) if class_weights is None else class_weights).to(self.device)
        self._graph = None

    def _step_impl(self, points, target):
        self.grads.zero()
        pred, feat = self.model(points.transpose(2, 1))
        loss = self.criterion(pred.contiguous().view(-1, self.num_classes), target, feat, self.class_weights)
        loss.backward()
        self.grads.adopt()
        self.grads.all_reduce_mean()
        self.optimizer.step()
        return loss.detach()

    def enable_cuda_graph(self, batch_clouds, npoint, channels, warmup=3):
        """Capture the whole training step (forward, loss, backward, gradient all-reduce, Adam) into ONE
        CUDA graph replayed per step: the ~420 kernel launches of a step cost no host time any more.
        Inputs are copied into static buffers; the FPS start indices stay a fresh CPU-generator draw per
        step (drawn on the host before each replay into pinned buffers the graph's memcpy nodes read).
        Re-capture after changing BatchNorm momentum or the learning-rate schedule's Python state."""
        from .modules import PointNetSetAbstraction
        dev = self.device
        self.model.train()
        self._sa = [m for m in self.model.modules() if isinstance(m, PointNetSetAbstraction) and not m.group_all]
        for m in self._sa:
            m.use_static_start_buffers(True)
        self._g_points = torch.zeros(batch_clouds, npoint, channels, device=dev)
        self._g_target = torch.zeros(batch_clouds * npoint, dtype=torch.int64, device=dev)
        self._g_points.uniform_(-0.5, 0.5)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = [(p.detach().clone()) for p in self.model.state_dict().values()]
        opt_state = None
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self._g_points, self._g_target)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the warm-up steps must not count as training: restore parameters/buffers and optimizer moments
        with torch.no_grad():
            for t, s in zip(self.model.state_dict().values(), saved):
                t.copy_(s)
            for st in self.optimizer.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._g_loss = self._step_impl(self._g_points, self._g_target)
        with torch.no_grad():                         # capture does not execute, but keep the state pristine anyway
            for t, s in zip(self.model.state_dict().values(), saved):
                t.copy_(s)
        self._graph = graph
        return self

    def step_device(self, points, target):
        """points [B, N, C] (point-major, as the DataLoader yields it) and target [B*N], on the device."""
        self.model.train()
        if self._graph is None:
            return self._step_impl(points, target)
        self._g_points.copy_(points, non_blocking=True)
        self._g_target.copy_(target.view(-1), non_blocking=True)
        for m in self._sa:                            # the reference's per-forward randint draws, in module order
            m.start_staging.draw()
        self._graph.replay()
        return self._g_loss

    def step(self, points_host, target_host):
        """One training step from HOST buffers (pinned memory recommended); returns the loss as a float
        (a device->host read, like the reference's per-batch `seg_pred.cpu()`)."""
        if self._graph is not None:                  # host -> static device buffers directly
            self.model.train()
            self._g_points.copy_(points_host, non_blocking=True)
            self._g_target.copy_(target_host.view(-1), non_blocking=True)
            for m in self._sa:
                m.start_staging.draw()
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
    fresh CPU-generator draw per forward (pointnet2_utils.py:75), staged through pinned buffers."""

    def __init__(self, model, batch_clouds, npoint, channels, device="cuda", warmup=2):
        from .modules import PointNetSetAbstraction
        self.model = model.eval()
        self.device = dev = torch.device(device)
        self.batch = batch_clouds
        self._sa = [m for m in model.modules() if isinstance(m, PointNetSetAbstraction) and not m.group_all]
        for m in self._sa:
            m.use_static_start_buffers(True)
        self.points = torch.zeros(batch_clouds, npoint, channels, device=dev)
        self.points.uniform_(-0.5, 0.5)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.model(self.points.transpose(2, 1))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.log_probs, _ = self.model(self.points.transpose(2, 1))       # [B, npoint, classes]
            self.labels = self.log_probs.argmax(dim=2)
        self._host_labels = torch.empty(batch_clouds, npoint, dtype=torch.int64).pin_memory()

    def forward_device(self, points):
        """points [b <= batch, npoint, C] on the device -> (log_probs, labels) views of the static outputs."""
        b = points.shape[0]
        self.points[:b].copy_(points, non_blocking=True)
        for m in self._sa:
            m.start_staging.draw()
        self.graph.replay()
        return self.log_probs[:b], self.labels[:b]

    def predict_host(self, points_host):
        """points [b <= batch, npoint, C] on the host (pinned recommended) -> labels [b, npoint] on the host."""
        b = points_host.shape[0]
        self.points[:b].copy_(points_host, non_blocking=True)
        for m in self._sa:
            m.start_staging.draw()
        self.graph.replay()
        self._host_labels.copy_(self.labels, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_labels[:b]


@torch.no_grad()
def predict_blocks(model, blocks_host, batch_size=32, rank=0, world=1, device="cuda", use_graph=True):
    """sem_seg_testing-style inference (num_votes=1) over [nb, 4096, C] blocks held on the host:
    this rank labels its contiguous shard of blocks; no collective is needed (eval-mode BatchNorm
    uses running statistics, so blocks are independent).  Returns (lo, hi, labels [hi-lo, 4096] on host)."""
    model.eval()
    lo, hi = shard_range(blocks_host.shape[0], rank, world)
    out = torch.empty(hi - lo, blocks_host.shape[1], dtype=torch.int64)
    predictor = None
    if use_graph and hi - lo >= batch_size:
        predictor = SemSegPredictor(model, batch_size, blocks_host.shape[1], blocks_host.shape[2], device)
    for s in range(lo, hi, batch_size):
        e = min(hi, s + batch_size)
        if predictor is not None:         # a short tail batch rides in the same fixed-shape graph
            out[s - lo:e - lo] = predictor.predict_host(blocks_host[s:e].float())
        else:
            x = blocks_host[s:e].to(device, non_blocking=True).float().transpose(2, 1)
            pred, _ = model(x)
            out[s - lo:e - lo] = pred.argmax(dim=2).cpu()
    return lo, hi, out
