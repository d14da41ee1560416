"""Host side of the hot path: the reference's free functions over the C ABI.

Same names, argument meaning and tensor conventions as
/root/reference/models/pointnet2_utils.py (point-first ``[B,N,3]`` inputs, int64
indices), CUDA float32 only.  Every function below is a thin argument-checking
wrapper around one or two entry points of libpn2b200.so; torch supplies device
memory and the current stream, nothing else.
"""
import contextlib

import torch

from . import _lib
from ._lib import call, dt, ptr, require_cuda, stream

_PRECISION = {"rows": torch.float32}


def set_precision(mode):
    """'fp32': rows and MLP arithmetic in fp32 (FMA pipes) -- matches the reference's CPU fp32.
    'bf16': rows stored as bf16, MLP products on the tcgen05 tensor cores with fp32 accumulation."""
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16', got %r" % (mode,))
    _PRECISION["rows"] = torch.float32 if mode == "fp32" else torch.bfloat16
    # the PyTorch-owned head of the network (two 128-wide matrix products) follows the mode: strict fp32, or TF32
    # tensor cores next to the bf16 hot path (what the reference's own CUDA path does by default through cuDNN)
    torch.backends.cuda.matmul.allow_tf32 = mode == "bf16"


def get_precision():
    return "fp32" if _PRECISION["rows"] == torch.float32 else "bf16"


def rows_dtype():
    return _PRECISION["rows"]


def _xyz3(t, name):
    require_cuda(t, name)
    if t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("%s must be [B, N, 3], got %s" % (name, tuple(t.shape)))
    return t


def square_distance(src, dst):
    """pointnet2_utils.py:19-40 -- [B,N,3] x [B,M,3] -> [B,N,M] in the reference's rounding order."""
    _xyz3(src, "src"), _xyz3(dst, "dst")
    src, dst = src.contiguous(), dst.contiguous()
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty(B, N, M, device=src.device, dtype=torch.float32)
    call("pn2_square_distance", ptr(src), ptr(dst), B, N, M, ptr(out), stream())
    return out


class _IndexPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        B, N, C = points.shape
        flat = idx.reshape(B, -1).contiguous()
        J = flat.shape[1]
        out = torch.empty(B, J, C, device=points.device, dtype=torch.float32)
        sB, sN, sC = points.stride()
        call("pn2_index_points", ptr(points), sB, sN, sC, B, N, C, ptr(flat), J, ptr(out), stream())
        ctx.save_for_backward(flat)
        ctx.dims = (B, N, C, J)
        return out.view(*idx.shape, C)

    @staticmethod
    def backward(ctx, dout):
        (flat,) = ctx.saved_tensors
        B, N, C, J = ctx.dims
        dout = dout.contiguous().view(B, J, C)
        dpoints = torch.zeros(B, N, C, device=dout.device, dtype=torch.float32)
        call("pn2_index_points_bwd", ptr(dout), ptr(flat), B, N, C, J, ptr(dpoints), stream())
        return dpoints, None


def index_points(points, idx):
    """pointnet2_utils.py:43-60 -- points [B,N,C], idx [B,S] or [B,S,K] (int64) -> [B,S(,K),C]."""
    require_cuda(points, "points")
    require_cuda(idx, "idx", torch.int64)
    if points.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("index_points: points %s / idx %s" % (tuple(points.shape), tuple(idx.shape)))
    return _IndexPoints.apply(points, idx)


class _OutputArena:
    def __init__(self, tensors=None):
        self.replay = tensors is not None
        self.tensors = list(tensors) if self.replay else []
        self.i = 0


_ARENA = None


def _out(shape, dtype, device):
    """Output tensor of an index operator: a fresh one, or -- inside reuse_outputs() -- the next tensor of the arena."""
    a = _ARENA
    if a is None:
        return torch.empty(shape, device=device, dtype=dtype)
    if not a.replay:
        t = torch.empty(shape, device=device, dtype=dtype)
        a.tensors.append(t)
        return t
    if a.i >= len(a.tensors):
        raise RuntimeError("reuse_outputs: more operator outputs than recorded tensors")
    t = a.tensors[a.i]
    a.i += 1
    if tuple(t.shape) != tuple(shape) or t.dtype != dtype:
        raise RuntimeError("reuse_outputs: output %d is %s %s, recorded %s %s" % (a.i - 1, tuple(shape), dtype, tuple(t.shape), t.dtype))
    return t


@contextlib.contextmanager
def record_outputs():
    """Collect, in call order, every output tensor the index operators (farthest_point_sample, query_ball_point,
    three_nn) allocate inside the block: `with record_outputs() as rec: ...; rec.tensors`."""
    global _ARENA
    prev, _ARENA = _ARENA, _OutputArena()
    try:
        yield _ARENA
    finally:
        _ARENA = prev


@contextlib.contextmanager
def reuse_outputs(tensors):
    """Run the same sequence of index operators again, writing into `tensors` (from record_outputs) instead of new
    allocations -- the two geometry slots of a software-pipelined CUDA graph pair (trainer.py) are filled this way, so no
    copy is needed to hand a batch's indices from the index pipeline to the feature path."""
    global _ARENA
    prev, _ARENA = _ARENA, _OutputArena(tensors)
    try:
        yield _ARENA
        if _ARENA.i != len(_ARENA.tensors):
            raise RuntimeError("reuse_outputs: %d of %d recorded outputs were produced" % (_ARENA.i, len(_ARENA.tensors)))
    finally:
        _ARENA = prev


class StartIndexGroup:
    """The start-index stagings of several call sites (the four set-abstraction levels) behind ONE pinned ring and ONE
    device buffer: draw() makes the reference's draws in module order and sends them with a single host->device copy
    instead of one per level (each ~6 us of serial stream time ahead of a replay).  Built after the members exist and
    BEFORE the graph that reads their device buffers is captured (the members' buffers become views of the group's)."""
    SLOTS = 8

    def __init__(self, members):
        self.members = list(members)
        dev = self.members[0].dev.device
        total = sum(m.B for m in self.members)
        self.host = torch.empty(self.SLOTS, total, dtype=torch.int64).pin_memory()
        self.dev = torch.empty(total, dtype=torch.int64, device=dev)
        off = 0
        for m in self.members:
            self.dev[off:off + m.B].copy_(m.dev)
            m.dev = self.dev[off:off + m.B]
            m.group_off = off          # no back-reference to the group: a reference cycle would leave the pinned rings to the
            off += m.B                 # cyclic collector, which may run (and free pinned memory) in the middle of a capture
        self._events = [None] * self.SLOTS
        self._slot = 0

    def draw(self):
        j = self._slot
        self._slot = (j + 1) % self.SLOTS
        if self._events[j] is not None:
            self._events[j].synchronize()
        for m in self.members:          # the reference's per-forward draws, in module order (:75)
            self.host[j, m.group_off:m.group_off + m.B].copy_(torch.randint(0, m.N, (m.B,), dtype=torch.long))
        self.dev.copy_(self.host[j], non_blocking=True)
        ev = self._events[j] = self._events[j] or torch.cuda.Event()
        ev.record()


class StartIndexStaging:
    """Host ring + ONE static device buffer for the FPS start indices of one call site.  A captured CUDA graph reads the
    device buffer; before every replay draw() makes the reference's torch.randint draw (:75) into the next pinned ring
    slot and enqueues its host->device copy on the current stream -- ordinary stream work ahead of the replay, so the
    host may run several replays ahead without ever rewriting indices a queued replay has not consumed yet (a ring slot
    is only reused after its own copy has completed).  See trainer.SemSegTrainer.enable_cuda_graph."""
    SLOTS = 8

    def __init__(self, B, N, device):
        self.B, self.N = B, N
        self.host = torch.empty(self.SLOTS, B, dtype=torch.int64).pin_memory()
        self.dev = torch.empty(B, dtype=torch.int64, device=device)
        self._events = [None] * self.SLOTS
        self._slot = 0

    def draw(self):
        """One CPU-generator draw, exactly the reference's torch.randint(0, N, (B,), dtype=torch.long), sent to the device."""
        j = self._slot
        self._slot = (j + 1) % self.SLOTS
        if self._events[j] is not None:
            self._events[j].synchronize()
        self.host[j].copy_(torch.randint(0, self.N, (self.B,), dtype=torch.long))
        self.dev.copy_(self.host[j], non_blocking=True)
        ev = self._events[j] = self._events[j] or torch.cuda.Event()
        ev.record()

    def upload(self):
        return self.dev


def farthest_point_sample(xyz, npoint, start=None, return_xyz=False, staging=None):
    """pointnet2_utils.py:63-84 -- xyz [B,N,3] (any strides) -> int64 [B,npoint].

    The start index is drawn exactly as the reference does (:75): one
    ``torch.randint(0, N, (B,), dtype=torch.long)`` on the CPU default generator,
    so both stay in lock-step under ``torch.manual_seed``.  ``start`` overrides it.
    """
    _xyz3(xyz, "xyz")
    B, N, _ = xyz.shape
    npoint = int(npoint)
    if staging is not None and start is None:
        if not torch.cuda.is_current_stream_capturing():
            staging.draw()
        start = staging.upload()
    else:
        if start is None:
            start = torch.randint(0, N, (B,), dtype=torch.long)
        if start.shape != (B,) or start.dtype != torch.int64:
            raise ValueError("start must be int64 [B]")
        if not start.is_cuda and B and (int(start.min()) < 0 or int(start.max()) >= N):
            raise ValueError("start indices must lie in [0, %d)" % N)      # (the kernel reads xyz[start] unchecked)
        start = start.to(xyz.device, non_blocking=True)
    out = _out((B, npoint), torch.int64, xyz.device)
    new_xyz = _out((B, npoint, 3), torch.float32, xyz.device) if return_xyz else None
    sB, sN, sC = xyz.stride()
    call("pn2_farthest_point_sample", ptr(xyz), sB, sN, sC, B, N, npoint, ptr(start), ptr(out), ptr(new_xyz), stream())
    return (out, new_xyz) if return_xyz else out


def radius_sq(radius):
    """float32(radius ** 2) as `sqrdists > radius ** 2` (:102) sees it: a Python double squared,
    then rounded once to fp32 by the tensor/scalar comparison."""
    return float(torch.tensor(float(radius) ** 2, dtype=torch.float64).to(torch.float32).item())


BALL_GRID = True          # query_ball_point through the cell grid (csrc/ballgrid.cu) when the cloud is big enough to pay
_BALL_GRID_MIN_N = 256
_BALL_GRID_MAX_N = 16384     # beyond: per-query N-bit maps cost more than the scan's early exit saves (config 3: 13.7 vs 10.9 ms)


def query_ball_point(radius, nsample, xyz, new_xyz, return_count=False):
    """pointnet2_utils.py:87-107 -- first nsample in-radius indices in index order, padded with the first.
    Clouds of >= 256 points go through the cell grid (~100 distance evaluations per query instead of N), smaller ones
    through the index-order scan; the result is the same bit for bit (tests/test_gpu_ops.py)."""
    _xyz3(xyz, "xyz"), _xyz3(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    nsample = int(nsample)
    out = _out((B, S, nsample), torch.int64, xyz.device)
    cnt = _out((B, S), torch.int32, xyz.device) if return_count else None
    sB, sN, sC = xyz.stride()
    qB, qN, qC = new_xyz.stride()
    if BALL_GRID and _BALL_GRID_MIN_N <= N <= _BALL_GRID_MAX_N and float(radius) > 0.0:
        nbytes = _lib.load().pn2_ball_grid_workspace_bytes(B, N)
        ws = torch.empty(nbytes, device=xyz.device, dtype=torch.uint8)
        call("pn2_query_ball_point_grid", ptr(xyz), sB, sN, sC, ptr(new_xyz), qB, qN, qC, B, N, S, float(radius),
             radius_sq(radius), nsample, ptr(out), ptr(cnt), ptr(ws), nbytes, stream())
    else:
        call("pn2_query_ball_point", ptr(xyz), sB, sN, sC, ptr(new_xyz), qB, qN, qC, B, N, S, radius_sq(radius),
             nsample, ptr(out), ptr(cnt), stream())
    return (out, cnt) if return_count else out


# three_nn through the cell grid of csrc/ballgrid.cu (pn2_three_nn_grid): identical output (tests/test_gpu_ops.py), OFF by default --
# measured on B200, 32 facade clouds: 4096 fine / 1024 coarse 96.6 us vs 108.8 us for the index-order scan, 1024 / 256 60.8 vs 38.5 us;
# with smaller cells (43 instead of 92 candidates per query) 142.6 / 67.8 us: one thread per query reads its candidates with
# divergent 16-byte loads, which costs what the scan's one shared-memory broadcast per candidate saves, and a single query that
# falls back to the full scan holds its whole warp.  It needs queries processed in cell order to pay.
THREE_NN_GRID = False
_NN_GRID_MIN_S = 256


def three_nn(xyz1, xyz2, fallback_count=None):
    """pointnet2_utils.py:296-302 -- (idx [B,N,3] int64, weight [B,N,3]) of the three nearest xyz2 points.
    fallback_count: optional 1-element int32 CUDA tensor; the grid search adds the number of queries that took the full scan."""
    _xyz3(xyz1, "xyz1"), _xyz3(xyz2, "xyz2")
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = _out((B, N, 3), torch.int64, xyz1.device)
    w = _out((B, N, 3), torch.float32, xyz1.device)
    aB, aN, aC = xyz1.stride()
    cB, cN, cC = xyz2.stride()
    if THREE_NN_GRID and _NN_GRID_MIN_S <= S <= _BALL_GRID_MAX_N:
        nbytes = _lib.load().pn2_ball_grid_workspace_bytes(B, S)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz1.device)
        call("pn2_three_nn_grid", ptr(xyz1), aB, aN, aC, ptr(xyz2), cB, cN, cC, B, N, S, ptr(idx), ptr(w), ptr(ws), nbytes,
             ptr(fallback_count), stream())
        return idx, w
    call("pn2_three_nn", ptr(xyz1), aB, aN, aC, ptr(xyz2), cB, cN, cC, B, N, S, ptr(idx), ptr(w), stream())
    return idx, w


def as_rows(t):
    """[B, N, C] fp32 view with arbitrary strides -> contiguous point-major rows (copy only if needed)."""
    if t.is_contiguous():
        return t
    B, N, C = t.shape
    out = torch.empty(B, N, C, device=t.device, dtype=torch.float32)
    sB, sN, sC = t.stride()
    call("pn2_to_rows", ptr(t), sB, sN, sC, B, N, C, ptr(out), N * C, C, stream())
    return out


def group_rows(xyz, new_xyz, feats, idx, ld, dtype):
    """sample_and_group's gather (:127-132): rows [(b,s,k), ld] = [xyz[idx]-new_xyz | feats[idx] | 0]."""
    B, N, _ = xyz.shape
    S, K = idx.shape[1], idx.shape[2]
    D = 0 if feats is None else feats.shape[2]
    rows = torch.empty(B * S * K, ld, device=xyz.device, dtype=dtype)
    sB, sN, sC = xyz.stride()
    fB, fN, fD = (0, 0, 0) if feats is None else feats.stride()
    call("pn2_group_points", ptr(xyz), sB, sN, sC, ptr(new_xyz), ptr(feats), fB, fN, fD, ptr(idx), B, N, S, K, D,
         ptr(rows), ld, dt(rows), stream())
    return rows


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """pointnet2_utils.py:110-138 (forward only; the modules use the fused autograd path)."""
    _xyz3(xyz, "xyz")
    B, N, C = xyz.shape
    fps_idx, new_xyz = farthest_point_sample(xyz, npoint, return_xyz=True)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    feats = None if points is None else as_rows(require_cuda(points, "points"))
    D = 0 if feats is None else feats.shape[2]
    rows = group_rows(xyz, new_xyz, feats, idx, 3 + D, torch.float32)
    new_points = rows.view(B, npoint, nsample, 3 + D)
    if returnfps:
        return new_xyz, new_points, index_points(xyz, idx), fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet2_utils.py:141-158 -- one group holding every point; pure views/concat."""
    _xyz3(xyz, "xyz")
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device, dtype=xyz.dtype)
    grouped = xyz.reshape(B, 1, N, C)
    if points is not None:
        grouped = torch.cat([grouped, points.reshape(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped


def new_vote_pool(num_points, num_classes, device="cuda"):
    """The [P, NC] vote pool of /root/reference/localfunctions.py:385 (np.zeros((P, NUM_CLASSES))), as int32 counts in HBM."""
    return torch.zeros(int(num_points), int(num_classes), dtype=torch.int32, device=device)


def add_vote(vote_label_pool, point_idx, pred_label, weight=None):
    """localfunctions.py:336-343 -- vote_label_pool[point_idx[b,n], pred_label[b,n]] += 1 for every (b, n) whose weight is
    neither 0 nor inf; the reference's Python double loop as one kernel.  `vote_label_pool` is a CUDA int32 [P, NC] tensor
    (new_vote_pool) updated in place and returned; `point_idx` / `pred_label` are [B, N] integer (or integral float, as the
    reference's np.zeros-backed batch arrays are) tensors, `weight` [B, N] float32/float64 or None."""
    require_cuda(vote_label_pool, "vote_label_pool", dtype=None)
    if vote_label_pool.dtype != torch.int32 or vote_label_pool.dim() != 2 or not vote_label_pool.is_contiguous():
        raise TypeError("vote_label_pool must be a contiguous int32 [P, NC] tensor (ops.new_vote_pool)")
    dev = vote_label_pool.device
    if tuple(point_idx.shape) != tuple(pred_label.shape) or (weight is not None and tuple(weight.shape) != tuple(pred_label.shape)):
        raise ValueError("point_idx, pred_label and weight must have the same shape")
    pi = point_idx.to(dev, non_blocking=True).long().contiguous()
    pl = pred_label.to(dev, non_blocking=True).long().contiguous()
    w = None
    if weight is not None:
        w = weight.to(dev, non_blocking=True)
        if w.dtype not in (torch.float32, torch.float64):
            w = w.float()
        w = w.contiguous()
    P, NC = vote_label_pool.shape
    call("pn2_add_vote", ptr(pi), ptr(pl), ptr(w), int(w is not None and w.dtype == torch.float64), pi.numel(), P, NC,
         ptr(vote_label_pool), None, stream())
    return vote_label_pool


def vote_argmax(vote_label_pool, dtype=torch.int64):
    """localfunctions.py:405 -- np.argmax(vote_label_pool, 1): the first class with the most votes, per point."""
    require_cuda(vote_label_pool, "vote_label_pool", dtype=None)
    if vote_label_pool.dtype != torch.int32 or vote_label_pool.dim() != 2 or not vote_label_pool.is_contiguous():
        raise TypeError("vote_label_pool must be a contiguous int32 [P, NC] tensor (ops.new_vote_pool)")
    if dtype not in (torch.int64, torch.uint8):
        raise TypeError("labels come back as int64 or uint8")
    P, NC = vote_label_pool.shape
    labels = torch.empty(P, dtype=dtype, device=vote_label_pool.device)
    call("pn2_vote_argmax", ptr(vote_label_pool), P, NC, ptr(labels), int(dtype == torch.uint8), stream())
    return labels


class RotationStaging:
    """Host ring + one static device buffer for the per-cloud (cos, sin) of rotate_point_cloud_z_, same life cycle as
    StartIndexStaging: the host may queue several steps ahead without rewriting angles a queued kernel has not read."""
    SLOTS = 8

    def __init__(self, B, device):
        self.B = B
        self.host = torch.empty(self.SLOTS, B, 2, dtype=torch.float64).pin_memory()
        self.dev = torch.empty(B, 2, dtype=torch.float64, device=device)
        self._events = [None] * self.SLOTS
        self._slot = 0

    def send(self, cos_sin):
        j = self._slot
        self._slot = (j + 1) % self.SLOTS
        if self._events[j] is not None:
            self._events[j].synchronize()
        self.host[j].copy_(cos_sin)
        self.dev.copy_(self.host[j], non_blocking=True)
        ev = self._events[j] = self._events[j] or torch.cuda.Event()
        ev.record()
        return self.dev


def draw_rotation_angles(B):
    """The reference's draws (provider.py:76-77): one np.random.uniform() * 2 * np.pi per cloud, in cloud order, from numpy's
    global generator.  Returns a float64 [B, 2] tensor of (cos, sin)."""
    import numpy as np
    ang = np.array([np.random.uniform() * 2 * np.pi for _ in range(B)])
    return torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)], axis=1))


def rotate_point_cloud_z_(points, cos_sin=None, staging=None):
    """provider.rotate_point_cloud_z (/root/reference/provider.py:66-84) on the xyz channels of a CUDA fp32 batch
    `points` [B, N, C >= 3], IN PLACE (the reference's `points[:, :, :3] = provider.rotate_point_cloud_z(points[:, :, :3])`,
    localfunctions.py:205).  cos_sin: float64 [B, 2] per-cloud (cos, sin); None draws the angles exactly as the reference
    does (draw_rotation_angles).  Returns `points`."""
    require_cuda(points, "points")
    if points.dim() != 3 or points.shape[2] < 3:
        raise ValueError("points must be [B, N, C >= 3], got %s" % (tuple(points.shape),))
    B, N, _ = points.shape
    if cos_sin is None:
        cos_sin = draw_rotation_angles(B)
    if tuple(cos_sin.shape) != (B, 2) or cos_sin.dtype != torch.float64:
        raise ValueError("cos_sin must be float64 [B, 2]")
    if cos_sin.is_cuda:
        cs = cos_sin.contiguous()
    elif staging is not None:
        cs = staging.send(cos_sin)
    else:
        cs = cos_sin.to(points.device, non_blocking=True).contiguous()
    sB, sN, sC = points.stride()
    call("pn2_rotate_z", ptr(points), sB, sN, sC, ptr(cs), B, N, stream())
    return points


def slice_scene(points, labels=None, extra=None, extra_names=(), labelweights=None, block_size=1.0, stride=0.5,
                padding=0.001, block_points=4096, generator=None):
    """TestCustomDataset.__getitem__ (/root/reference/sem_seg_testing.py:182-254) for one scene, on the device.

    points [P, 3] float64 CUDA (scene coordinates as laspy yields them), labels [P] int64 or None, extra [E, P] float64 or
    None with `extra_names` (features named red / green / blue are divided by 255, :236-237), labelweights [NC] float32.
    Returns (data_room [nb, block_points, 6 + E] float32, label_room [nb, block_points] int64, sample_weight
    [nb, block_points] float32, index_room [nb, block_points] int64), all on the device: the reference's four arrays, the
    rows already rounded to float32 as `torch.Tensor(batch_data)` does before the forward (localfunctions.py:394).

    Same cells in the same (index_y, index_x) order, same member points per cell (the reference's float64 comparisons),
    same number of blocks per cell, bit-identical rows per (cell, point); WHICH members pad a cell and the order inside a
    cell are random in the reference as well (numpy's global generator) -- here they come from `generator` (a CUDA
    torch.Generator, default: the device's)."""
    import numpy as np
    require_cuda(points, "points", dtype=torch.float64)
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must be [P, 3] float64, got %s" % (tuple(points.shape),))
    dev, P = points.device, points.shape[0]
    E = 0 if extra is None else extra.shape[0]
    if E != len(extra_names):
        raise ValueError("extra has %d rows but %d names were given" % (E, len(extra_names)))
    C = 6 + E
    bp = int(block_points)

    def empty():
        return (torch.empty(0, bp, C, device=dev), torch.empty(0, bp, dtype=torch.int64, device=dev),
                torch.empty(0, bp, device=dev), torch.empty(0, bp, dtype=torch.int64, device=dev))

    if P == 0:
        return empty()
    cmin = points.amin(dim=0).cpu().numpy()            # np.amin / np.amax of :186 (exact)
    cmax = points.amax(dim=0).cpu().numpy()
    block_size, stride, padding = float(block_size), float(stride), float(padding)
    gx = int(np.ceil(float(cmax[0] - cmin[0] - block_size) / stride) + 1)     # :187-188
    gy = int(np.ceil(float(cmax[1] - cmin[1] - block_size) / stride) + 1)
    if gx <= 0 or gy <= 0:
        return empty()

    def axis(lo, hi, g):                               # :196-201 per column / row, the reference's own expressions
        los, his, ctr = np.empty(g), np.empty(g), np.empty(g)
        for i in range(g):
            s = lo + i * stride
            e = min(s + block_size, hi)
            s = e - block_size
            los[i], his[i], ctr[i] = s - padding, e + padding, s + block_size / 2.0
        return los, his, ctr

    lo_x, hi_x, cx = axis(cmin[0], cmax[0], gx)
    lo_y, hi_y, cy = axis(cmin[1], cmax[1], gy)
    bounds = torch.from_numpy(np.concatenate([lo_x, hi_x, cx, lo_y, hi_y, cy])).to(dev)
    d_lo_x, d_hi_x, d_cx = bounds[0:gx], bounds[gx:2 * gx], bounds[2 * gx:3 * gx]
    d_lo_y, d_hi_y, d_cy = bounds[3 * gx:3 * gx + gy], bounds[3 * gx + gy:3 * gx + 2 * gy], bounds[3 * gx + 2 * gy:]
    sP, sC = points.stride()
    cells = gx * gy
    counts = torch.zeros(cells, dtype=torch.int32, device=dev)

    def walk(cell_offset, slot_point, slot_cell):
        call("pn2_slice_cells", ptr(points), sP, sC, P, ptr(d_lo_x), ptr(d_hi_x), gx, ptr(d_lo_y), ptr(d_hi_y), gy,
             float(cmin[0]), float(cmin[1]), stride, block_size, padding, ptr(counts), ptr(cell_offset), ptr(slot_point),
             ptr(slot_cell), stream())

    walk(None, None, None)                             # pass 1: members per cell
    n = counts.long()
    padded = (n + bp - 1) // bp * bp                    # :204-205 (empty cells are skipped, :202)
    cell_offset = torch.cumsum(padded, 0) - padded
    S = int(padded.sum())
    if S == 0:
        return empty()
    slot_point = torch.empty(S, dtype=torch.int64, device=dev)
    slot_cell = torch.empty(S, dtype=torch.int32, device=dev)
    counts.zero_()
    walk(cell_offset, slot_point, slot_cell)           # pass 2: member lists (counts is the per-cell cursor)
    cell_of_slot = torch.repeat_interleave(torch.arange(cells, device=dev), padded, output_size=S)
    rank = torch.arange(S, device=dev) - cell_offset[cell_of_slot]
    is_pad = rank >= n[cell_of_slot]

    def rand(k):
        return torch.randint(0, 2 ** 31, (k,), device=dev, dtype=torch.int64, generator=generator)

    # a random permutation of every cell's members (padding slots stay behind them)
    order = torch.argsort((cell_of_slot << 33) | (is_pad.long() << 32) | rand(S))
    slot_point = slot_point[order]
    pad_slots = torch.nonzero(is_pad).squeeze(1)
    if pad_slots.numel():
        pad_cell = cell_of_slot[pad_slots].int()
        pad_rank = (rank[pad_slots] - n[cell_of_slot[pad_slots]]).contiguous()
        call("pn2_slice_pad", ptr(counts), ptr(cell_offset), ptr(pad_cell), ptr(pad_rank), ptr(rand(pad_slots.numel())),
             pad_slots.numel(), bp, ptr(slot_point), ptr(slot_cell), stream())
    slot_cell = cell_of_slot.int()
    # np.random.shuffle of the padded list (:209)
    order = torch.argsort((cell_of_slot << 32) | rand(S))
    slot_point = slot_point[order].contiguous()
    rows = torch.empty(S, C, device=dev, dtype=torch.float32)
    out_label = torch.empty(S, dtype=torch.int64, device=dev)
    out_weight = torch.empty(S, dtype=torch.float32, device=dev)
    div = None
    eE = eP = 0
    if E:
        require_cuda(extra, "extra", dtype=torch.float64)
        div = torch.tensor([255.0 if nm in ("red", "green", "blue") else 1.0 for nm in extra_names], dtype=torch.float64, device=dev)
        eE, eP = extra.stride()
    lw = None if labelweights is None else labelweights.to(dev, torch.float32).contiguous()
    lab = None if labels is None else require_cuda(labels, "labels", dtype=torch.int64).contiguous()
    call("pn2_slice_rows", ptr(points), sP, sC, ptr(lab), ptr(extra), eE, eP, ptr(div), E, ptr(lw), ptr(slot_point),
         ptr(slot_cell), ptr(d_cx), ptr(d_cy), gx, float(cmax[0]), float(cmax[1]), float(cmax[2]), S, ptr(rows), ptr(out_label),
         ptr(out_weight), stream())
    nb = S // bp
    return rows.view(nb, bp, C), out_label.view(nb, bp), out_weight.view(nb, bp), slot_point.view(nb, bp)


def sample_training_crops(points, labels, batch, extra=None, extra_names=(), coord_max=None, num_point=4096, block_size=1.0,
                          min_points=1024, generator=None, max_rounds=50):
    """`batch` items of TrainCustomDataset.__getitem__ (/root/reference/sem_seg_training.py:200-259) for one room, on the
    device: random centre point, the block_size x block_size column around it, redrawn until it holds more than
    `min_points` points (:206-215), `num_point` of them chosen without replacement when there are enough, with replacement
    otherwise (:217-220), rows [x - cx, y - cy, z, x/max, y/max, z/max, extras (/255 for colours)] (:223-252).

    points [N, 3] float64 CUDA, labels [N] int64 CUDA, extra [E, N] float64 CUDA.  Returns (features [batch, num_point,
    6 + E] float32, labels [batch, num_point] int64, centre point index [batch] int64, selected point indices [batch,
    num_point] int64), on the device.  Membership and rows are the reference's float64 arithmetic; all the choices are
    random in the reference too (numpy's generator there, `generator` -- a CUDA torch.Generator -- here)."""
    import ctypes
    import numpy as np
    require_cuda(points, "points", dtype=torch.float64)
    require_cuda(labels, "labels", dtype=torch.int64)
    dev, N = points.device, points.shape[0]
    if points.dim() != 2 or points.shape[1] != 3 or N == 0:
        raise ValueError("points must be a non-empty [N, 3] float64 tensor")
    E = 0 if extra is None else extra.shape[0]
    if E != len(extra_names):
        raise ValueError("extra has %d rows but %d names were given" % (E, len(extra_names)))
    cmax = points.amax(dim=0).cpu().numpy() if coord_max is None else np.asarray(coord_max, dtype=np.float64)
    sP, sC = points.stride()
    half = float(block_size) / 2.0
    centre_idx = torch.empty(batch, dtype=torch.int64, device=dev)
    centres = torch.empty(batch, 3, dtype=torch.float64, device=dev)
    counts_all = torch.zeros(batch, dtype=torch.int64, device=dev)
    todo = list(range(batch))

    def boxes_of(c):                                   # :208-209 centre -+ [block_size / 2, block_size / 2, 0] in float64
        c = c.cpu().numpy()
        return np.ascontiguousarray(np.stack([c[:, 0] - half, c[:, 0] + half, c[:, 1] - half, c[:, 1] + half], axis=1))

    def members(boxes, counts, offset, slot_point, slot_crop):
        for k0 in range(0, boxes.shape[0], 64):        # <= 64 boxes per launch
            bx = np.ascontiguousarray(boxes[k0:k0 + 64])
            call("pn2_crop_members", ptr(points), sP, sC, N, bx.ctypes.data_as(ctypes.c_void_p), bx.shape[0], k0,
                 ptr(counts[k0:]), ptr(None if offset is None else offset[k0:]), ptr(slot_point), ptr(slot_crop), stream())

    for _ in range(max_rounds):                        # the reference's `while True` (:206), all pending crops per round
        k = len(todo)
        draw = torch.randint(0, N, (k,), device=dev, generator=generator)       # np.random.choice(N_points)
        c = points[draw]
        counts = torch.zeros(k, dtype=torch.int32, device=dev)
        members(boxes_of(c), counts, None, None, None)
        ok = (counts > min_points).cpu().numpy()
        rows = torch.tensor(todo, device=dev)
        good = torch.from_numpy(ok).to(dev)
        centre_idx[rows[good]] = draw[good]
        centres[rows[good]] = c[good]
        counts_all[rows[good]] = counts[good].long()
        todo = [t for t, g in zip(todo, ok) if not g]
        if not todo:
            break
    else:
        raise RuntimeError("no block with more than %d points found after %d rounds" % (min_points, max_rounds))
    offset = torch.cumsum(counts_all, 0) - counts_all
    total = int(counts_all.sum())
    slot_point = torch.empty(total, dtype=torch.int64, device=dev)
    slot_crop = torch.empty(total, dtype=torch.int32, device=dev)
    cursor = torch.zeros(batch, dtype=torch.int32, device=dev)
    members(boxes_of(centres), cursor, offset, slot_point, slot_crop)
    # a random permutation of every crop's members; the first num_point of it = np.random.choice(..., replace=False)
    key = torch.randint(0, 2 ** 31, (total,), device=dev, dtype=torch.int64, generator=generator)
    order = torch.argsort((slot_crop.long() << 32) | key)
    slot_point = slot_point[order]
    j = torch.arange(num_point, device=dev).unsqueeze(0).expand(batch, num_point)
    with_repl = torch.randint(0, 2 ** 62, (batch, num_point), device=dev, dtype=torch.int64, generator=generator) % counts_all.unsqueeze(1)
    pick = torch.where((counts_all >= num_point).unsqueeze(1), j, with_repl)      # :217-220
    sel = slot_point[(offset.unsqueeze(1) + pick).reshape(-1)].contiguous()
    S = batch * num_point
    C = 6 + E
    feats = torch.empty(S, C, device=dev, dtype=torch.float32)
    out_label = torch.empty(S, dtype=torch.int64, device=dev)
    crop_of_slot = torch.arange(batch, device=dev, dtype=torch.int32).repeat_interleave(num_point)
    div, eE, eP = None, 0, 0
    if E:
        require_cuda(extra, "extra", dtype=torch.float64)
        div = torch.tensor([255.0 if nm in ("red", "green", "blue") else 1.0 for nm in extra_names], dtype=torch.float64, device=dev)
        eE, eP = extra.stride()
    cx, cy = centres[:, 0].contiguous(), centres[:, 1].contiguous()
    call("pn2_slice_rows", ptr(points), sP, sC, ptr(labels.contiguous()), ptr(extra), eE, eP, ptr(div), E, None, ptr(sel),
         ptr(crop_of_slot), ptr(cx), ptr(cy), 0, float(cmax[0]), float(cmax[1]), float(cmax[2]), S, ptr(feats), ptr(out_label),
         None, stream())
    return feats.view(batch, num_point, C), out_label.view(batch, num_point), centre_idx, sel.view(batch, num_point)
