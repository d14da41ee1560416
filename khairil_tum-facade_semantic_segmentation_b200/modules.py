"""The reference's set-abstraction / feature-propagation modules over libpn2b200.

Same constructor and ``forward`` signatures, same channel-first tensor
conventions and -- because ``mlp_convs`` / ``mlp_bns`` stay ``nn.ModuleList``s of
real ``nn.Conv2d/Conv1d`` and ``nn.BatchNorm2d/1d`` -- the same ``state_dict``
as /root/reference/models/pointnet2_utils.py:161-315.  The convolution and
batch-norm modules only OWN the parameters; their arithmetic is done by the
CUDA kernels (weights, running statistics, momentum, eps and train/eval mode are
read from them at call time, running statistics are written back).

One autograd.Function per module: forward = FPS -> ball query -> gather -> MLP
(-> max over nsample); backward is hand written on the same kernels.
"""
import contextlib
import ctypes
import os
import weakref

import torch
import torch.nn as nn

from . import ops
from ._lib import BnFinalize, call, dt, load, ptr, require_cuda, stream


def _round_up(v, m):
    return (v + m - 1) // m * m


def _row_ld(k, dtype):
    # 16-byte aligned rows for the bf16 (tensor-core) path; fp32 rows are packed
    return k if dtype == torch.float32 else _round_up(k, 8)


class _LayerState:
    __slots__ = ("W", "K", "N", "Z", "scale", "shift", "mean", "invstd", "train", "has_bias", "wpack_bwd", "dgamma", "dbeta")


_GRAD_MODE = [True]       # grad mode at the module's forward() (inside autograd.Function.forward it always reads False)


def _note_grad_mode():
    _GRAD_MODE[0] = torch.is_grad_enabled()


def _want_backward(ctx):
    """Will this forward be differentiated?  (ctx.needs_input_grad ignores torch.no_grad())"""
    return _GRAD_MODE[0] and any(ctx.needs_input_grad)


# Inference with FROZEN parameters (opt-in: `with frozen_parameters():`, used by trainer.SemSegPredictor): folded BatchNorm
# (scale, shift) and packed weight images are a function of the parameters and running statistics only, so inside the
# context they are computed once and reused.  It is opt-in because nothing can invalidate such a cache reliably: CUDA-graph
# replays of a training step and this library's own kernels update parameters / running statistics without moving torch's
# version counters.  Entering the (outermost) context starts from an empty cache; the version check below only catches
# ordinary in-place updates made while the context is active.
_EVAL_CACHE = {}
_FROZEN = [0]


@contextlib.contextmanager
def frozen_parameters():
    if _FROZEN[0] == 0:
        _EVAL_CACHE.clear()
    _FROZEN[0] += 1
    try:
        yield _EVAL_CACHE
    finally:
        _FROZEN[0] -= 1


def _eval_key_versions(convs, bns):
    ts = []
    for conv, bn in zip(convs, bns):
        ts += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    return tuple((None if t is None else (t.data_ptr(), t._version, t.device.index)) for t in ts) + tuple(float(bn.eps) for bn in bns)


def _eval_cache_get(kind, convs, bns):
    if not _FROZEN[0]:
        return None
    ent = _EVAL_CACHE.get((kind,) + tuple(id(c) for c in convs))
    if ent is not None and ent[0] == _eval_key_versions(convs, bns):
        return ent[1]
    return None


def _eval_cache_put(kind, convs, bns, payload):
    # (tensors born inside a capture belong to that graph's pool: never cached)
    if _FROZEN[0] and not torch.cuda.is_current_stream_capturing():
        _EVAL_CACHE[(kind,) + tuple(id(c) for c in convs)] = (_eval_key_versions(convs, bns), payload)


def _bn_trains(bn):
    return bn.training or bn.running_mean is None


_ACCUM = {}
_ACCUM_COLS = 4096
_ACCUM_REPLICAS = 4          # PN2_STAT_REPLICAS of include/pn2b200.h


def _scratch(table, dev, make):
    """One zero-initialised scratch tensor per (device, stream) for eager launches, plus ONE per device shared by every
    CUDA-graph capture (allocated eagerly -- outside any graph pool -- the first time an eager launch needs scratch, which
    the warm-up passes before a capture always do).  The kernels leave these buffers zeroed, so neither eager steps nor
    graph replays need memset launches; captured graphs that use them must not be replayed concurrently with each other."""
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        buf = table.get((dev.index, "capture"))
        # no eager launch preceded this capture: a buffer created now lives in the graph's pool, zeroed by a memset node
        return buf if buf is not None else make()
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = table.get(key)
    if buf is None:
        buf = table[key] = make()
        if (dev.index, "capture") not in table:
            table[(dev.index, "capture")] = make()
    return buf


def _stat_accum(dev):
    """[PN2_STAT_REPLICAS][2][C <= 4096] fp64 accumulator of the column-sum epilogues.  It is zero whenever no
    producer/finalize pair is in flight on its stream: allocated zeroed, and every *_finalize entry point zeroes what it
    consumed (so a step needs no extra memset launches)."""
    return _scratch(_ACCUM, dev, lambda: torch.zeros(_ACCUM_REPLICAS * 2 * _ACCUM_COLS, device=dev, dtype=torch.float64))


_TICKET = {}


def _ticket(dev):
    """Zero-initialised uint32 the "last block finalizes" kernels count on; they leave it zero, so one serves every layer
    (same life cycle as _stat_accum)."""
    return _scratch(_TICKET, dev, lambda: torch.zeros(4, device=dev, dtype=torch.int32))


def _pack_weights(Ws, Ks, Ns, transposed, dev):
    """bf16 tensor-core images of several weights in ONE launch; returns the uint8 image tensors."""
    lib = load()
    n = len(Ws)
    imgs = [torch.empty(lib.pn2_linear_wpack_bytes(*((N, K) if t else (K, N))), device=dev, dtype=torch.uint8)
            for K, N, t in zip(Ks, Ns, transposed)]
    vp, ci = ctypes.c_void_p * n, ctypes.c_int * n
    call("pn2_pack_weights", n, vp(*[w.data_ptr() for w in Ws]), ci(*Ks), ci(*Ns), ci(*transposed),
         vp(*[i.data_ptr() for i in imgs]), stream())
    return imgs


_MOMENTUM_DEV = weakref.WeakKeyDictionary()      # BatchNorm module -> 1-element fp32 CUDA tensor holding its momentum


def set_momentum_buffers(mapping):
    """mapping: {BatchNorm module: 1-element fp32 CUDA tensor}.  The fused BatchNorm finalize of a train-mode layer then
    reads the momentum from that tensor at run time instead of baking `bn.momentum` into the launch, so a captured
    training step follows the reference's per-epoch momentum schedule (localfunctions.py:191-195) without re-capture.
    The owner (trainer.SemSegTrainer) keeps the tensors in step with `bn.momentum`."""
    for bn, t in mapping.items():
        _MOMENTUM_DEV[bn] = t


_STEP_IMAGES = {}          # tuple(id(conv) for conv in an MLP) -> (forward images, data-gradient images) packed by prepack_mlps


def _mlp_weights(convs, K0=None):
    """([W [N, K] fp32 views], [K per layer], [N per layer]) of a conv chain; K0 checks the first layer's input width."""
    Ws, Ks = [], []
    K = K0
    for conv in convs:
        N = conv.out_channels
        W = conv.weight.detach().reshape(N, -1)
        if not W.is_contiguous():
            W = W.contiguous()
        if K is None:
            K = W.shape[1]
        if W.shape[1] != K or W.dtype != torch.float32:
            raise ValueError("conv weight %s does not match %d input channels (fp32)" % (tuple(conv.weight.shape), K))
        Ws.append(W)
        Ks.append(K)
        K = N
    return Ws, Ks, Ks[1:] + [K]


def prepack_mlps(chains):
    """The bf16 tensor-core weight images (both orientations) of SEVERAL MLPs in ONE launch, ahead of a training step:
    `chains` lists the conv chains the coming forward will run (get_model.training_chains()).  mlp_forward picks its
    images up from here instead of packing per MLP (8 launches of 2-6 us each on the feature path's critical path).
    Images are valid for ONE forward: call again after the parameters changed (every optimizer step)."""
    _STEP_IMAGES.clear()
    if ops.rows_dtype() != torch.bfloat16:
        return
    Ws, Ks, Ns, tr, spans = [], [], [], [], []
    for convs in chains:
        convs = list(convs)
        if not convs or not convs[0].weight.is_cuda:
            continue
        w, k, n = _mlp_weights(convs)
        spans.append((tuple(id(c) for c in convs), len(Ws), len(w)))
        Ws += w + w
        Ks += k + k
        Ns += n + n
        tr += [0] * len(w) + [1] * len(w)
    if not Ws:
        return
    imgs = _pack_weights(Ws, Ks, Ns, tr, Ws[0].device)
    for key, first, n_l in spans:
        _STEP_IMAGES[key] = (imgs[first:first + n_l], imgs[first + n_l:first + 2 * n_l])


def mlp_forward(x0, K0, M, convs, bns, want_bwd=True):
    """relu(bn(conv(.))) chain on rows (pointnet2_utils.py:196-198 / :311-314) up to the last
    layer's pre-BN product; the caller applies the last BN+ReLU fused with its tail."""
    dev, dtype = x0.device, x0.dtype
    x, ldx, K = x0, x0.shape[1], K0
    in_scale = in_shift = None
    layers = []
    Ws, Ks, Ns = _mlp_weights(convs, K0)
    K = K0
    fwd_img = bwd_img = None
    # pure inference (every BatchNorm on running statistics, nothing to differentiate): folded (scale, shift) and the packed
    # weight images depend on the parameters only -- reuse them until a parameter / running statistic changes
    frozen = not want_bwd and not any(_bn_trains(bn) for bn in bns)
    cached = _eval_cache_get(("mlp", dtype), convs, bns) if frozen else None
    new_fold = []
    if cached is not None:
        fwd_img = cached[0]
    elif dtype == torch.bfloat16 and want_bwd and tuple(id(c) for c in convs) in _STEP_IMAGES:
        fwd_img, bwd_img = _STEP_IMAGES.pop(tuple(id(c) for c in convs))      # packed for the whole network by prepack_mlps
    elif dtype == torch.bfloat16:
        # every layer's weight image (and, for a backward pass, its transposed image for the data gradient) in one launch;
        # the first layer's data gradient is only needed when the MLP input wants a gradient -- pack it anyway, it is tiny
        n_l = len(Ws)
        tr = [0] * n_l + ([1] * n_l if want_bwd else [])
        imgs = _pack_weights(Ws * (2 if want_bwd else 1), Ks * (2 if want_bwd else 1), Ns * (2 if want_bwd else 1), tr, dev)
        fwd_img, bwd_img = imgs[:n_l], (imgs[n_l:] if want_bwd else None)
    for l, (conv, bn) in enumerate(zip(convs, bns)):
        st = _LayerState()
        N, W = Ns[l], Ws[l]
        bias = None if conv.bias is None else conv.bias.detach()
        gamma = None if bn.weight is None else bn.weight.detach()
        beta = None if bn.bias is None else bn.bias.detach()
        ldz = _row_ld(N, dtype)
        z = torch.empty(M, ldz, device=dev, dtype=dtype)
        stats = torch.empty(4, N, device=dev, dtype=torch.float32)
        st.W, st.K, st.N, st.Z, st.has_bias = W, K, N, z, bias is not None
        st.scale, st.shift, st.mean, st.invstd = stats[0], stats[1], stats[2], stats[3]
        st.train = _bn_trains(bn)
        st.wpack_bwd = None if bwd_img is None else bwd_img[l]
        wpack = None if fwd_img is None else fwd_img[l]
        if st.train:
            accum = _stat_accum(dev)
            momentum = bn.momentum
            nbt = bn.num_batches_tracked if (bn.num_batches_tracked is not None and bn.training) else None
            if momentum is None:      # cumulative moving average (nn.BatchNorm semantics): needs the counter's value
                if nbt is not None:
                    nbt.add_(1)
                momentum = 1.0 / float(nbt.item()) if nbt is not None else 0.0
                nbt = None
            update = bn.training and bn.running_mean is not None
            if wpack is not None:     # tensor-core rows: statistics -> scale/shift fused into the layer kernel
                mom_dev = _MOMENTUM_DEV.get(bn) if (update and bn.momentum is not None) else None
                fin = BnFinalize(ptr(_ticket(dev)), ptr(gamma), ptr(beta), ptr(bias), float(bn.eps), float(momentum),
                                 ptr(bn.running_mean) if update else None, ptr(bn.running_var) if update else None,
                                 ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd), ptr(nbt), ptr(mom_dev))
                call("pn2_linear_fwd_prepacked", ptr(x), ldx, dt(x), ptr(in_scale), ptr(in_shift), ptr(W), None, M, K, N,
                     ptr(z), ldz, dt(z), ptr(accum), ptr(wpack), ctypes.addressof(fin), stream())
            else:
                call("pn2_linear_fwd", ptr(x), ldx, dt(x), ptr(in_scale), ptr(in_shift), ptr(W), None, M, K, N,
                     ptr(z), ldz, dt(z), ptr(accum), None, stream())
                call("pn2_bn_train_finalize", ptr(accum), M, N, ptr(gamma), ptr(beta), ptr(bias), float(bn.eps),
                     float(momentum), ptr(bn.running_mean) if update else None, ptr(bn.running_var) if update else None,
                     ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd), ptr(nbt), stream())
        else:
            if cached is not None:
                st.scale, st.shift = cached[1][l]
            else:
                call("pn2_bn_eval_fold", ptr(gamma), ptr(beta), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps), N,
                     ptr(st.scale), ptr(st.shift), stream())
                new_fold.append((st.scale, st.shift))
            if wpack is not None:
                call("pn2_linear_fwd_prepacked", ptr(x), ldx, dt(x), ptr(in_scale), ptr(in_shift), ptr(W), ptr(bias), M, K,
                     N, ptr(z), ldz, dt(z), None, ptr(wpack), None, stream())
            else:
                call("pn2_linear_fwd", ptr(x), ldx, dt(x), ptr(in_scale), ptr(in_shift), ptr(W), ptr(bias), M, K, N,
                     ptr(z), ldz, dt(z), None, None, stream())
            if want_bwd:
                # for a backward pass through frozen statistics (Z already holds the bias here):
                # zhat = (z - running_mean) * invstd_running
                st.mean = bn.running_mean.detach()
                st.invstd = torch.rsqrt(bn.running_var.detach() + bn.eps)
        layers.append(st)
        x, ldx, K = z, ldz, N
        in_scale, in_shift = st.scale, st.shift
    if frozen and cached is None:
        _eval_cache_put(("mlp", dtype), convs, bns, (fwd_img, new_fold))
    return layers


WGRAD_ACCUMULATE = os.environ.get("PN2_WGRAD_DETERMINISTIC", "0") != "1" and os.environ.get("PN2_DISABLE_TC", "0") != "1"
# bf16 rows: weight gradients are added straight into the (zeroed) gradient buffer with L2 reductions -- no fp32 partials,
# no reduce launch; PN2_WGRAD_DETERMINISTIC=1 keeps the fixed-order partial + reduce path (pn2_linear_bwd_weight)


def _weight_grad(dZ, xin, ldx, sc, sh, M, K, N, param, dev, cuda_stream, keep):
    """dW [N, K] of one layer on `cuda_stream`: into the parameter's gradient-sink view when there is one."""
    lib = load()
    dW = _sink(param, (N, K))
    # K % 4 != 0 (the 3 + D wide first layers: 67, 131, 259): rows of dW are not 16-byte aligned, the reductions would be
    # scalar -- 4x the L2 operations on the same addresses (measured 107 us vs 27 us for sa3.1) -- so those keep the partials
    if (WGRAD_ACCUMULATE and K % 4 == 0 and dZ.dtype == torch.bfloat16 and xin.dtype == torch.bfloat16
            and dZ.shape[1] % 8 == 0 and ldx % 8 == 0):
        if dW is None:
            dW = torch.zeros(N, K, device=dev, dtype=torch.float32)      # a sink is zeroed once per step by its owner
        call("pn2_linear_bwd_weight_accum", ptr(dZ), dZ.shape[1], dt(dZ), ptr(xin), ldx, dt(xin), ptr(sc), ptr(sh), M, K, N,
             ptr(dW), cuda_stream)
        return dW
    if dW is None:
        dW = torch.empty(N, K, device=dev, dtype=torch.float32)
    scratch = torch.empty(lib.pn2_linear_wgrad_scratch_bytes(M, K, N), device=dev, dtype=torch.uint8)
    call("pn2_linear_bwd_weight", ptr(dZ), dZ.shape[1], dt(dZ), ptr(xin), ldx, dt(xin), ptr(sc), ptr(sh), M, K, N, ptr(dW),
         ptr(scratch), cuda_stream)
    keep.append(scratch)
    return dW


OVERLAP_WGRAD = True      # weight gradients on a side stream, concurrently with the data-gradient / BatchNorm chain
_SIDE = {}
_GRAD_SINK = {}           # id(parameter) -> (weak reference to it, flat fp32 view its gradient is written into)


def set_grad_sink(params, views):
    """The gradient sink of the hot path (trainer.FlatGradients): backward passes write the weight / BatchNorm gradients of
    `params` straight into `views` (preallocated fp32 tensors of the parameters' shapes, zeroed once per step by their
    owner) and hand autograd views of them -- no accumulate kernels.  The latest owner of a parameter wins; entries of
    parameters that no longer exist are dropped here and can never be hit: a sink is looked up by id() AND checked
    against a weak reference, because Python hands the id of a dead parameter to the next object allocated there -- a
    new model would otherwise write its gradients into a finished trainer's buffer, on top of that trainer's last step."""
    for k in [k for k, (ref, _) in _GRAD_SINK.items() if ref() is None]:
        del _GRAD_SINK[k]
    for p, v in zip(params, views):
        _GRAD_SINK[id(p)] = (weakref.ref(p), v)


def clear_grad_sink(params=None):
    """Forget the sink of `params` (None: every sink): their gradients go the ordinary autograd way again."""
    if params is None:
        _GRAD_SINK.clear()
        return
    for p in params:
        ent = _GRAD_SINK.get(id(p))
        if ent is not None and ent[0]() is p:
            del _GRAD_SINK[id(p)]


def _sink(param, shape=None):
    if param is None or not isinstance(param, nn.Parameter):
        return None
    ent = _GRAD_SINK.get(id(param))
    if ent is None or ent[0]() is not param:
        return None
    v = ent[1]
    if not v.is_cuda or v.dtype != torch.float32 or v.numel() != param.numel() or not v.is_contiguous():
        return None
    return v.view(shape if shape is not None else param.shape)      # a fresh tensor object every time


def _side_stream(dev):
    s = _SIDE.get(dev.index)
    if s is None:
        s = _SIDE[dev.index] = torch.cuda.Stream(device=dev)
    return s


FUSED_BWD = True          # bf16 rows: one fused kernel per layer (csrc/bwd_fused.cu) wherever the library takes the layer
FUSED_BWD_POOLED = False  # ... the pooled top layer too (da_mode 2).  Measured on B200 (config 2): the extra transform work costs more
                          # (+0.13 ms in the fused kernels, whose transform group is the bottleneck) than the dense
                          # pn2_pool_bn_relu_bwd_dz pass it replaces (0.12 ms at 4.3 TB/s): 2.66 vs 2.58 ms per step -- off


def _fused_bwd_layer(st, prev, x0, M, cur, mode, conv, bn_prev, want_dx, dev, dtype, keep, pooled=None):
    """One launch of pn2_mlp_bwd_layer for layer `st` (see include/pn2b200.h), or None when the library does not take it.
    cur / mode: the incoming gradient -- "dA" (dense, unmasked), "dA_masked", "dZ", or "pooled" (cur None, pooled = (dOut
    [G, N] fp32, arg [G, N] int32, nsample)).  Returns (dW, dA_prev, dgamma_prev, dbeta_prev)."""
    from ._lib import BwdLayer
    lib = load()
    if not FUSED_BWD or not WGRAD_ACCUMULATE or dtype != torch.bfloat16 or (cur is not None and cur.dtype != torch.bfloat16) or (
            st.wpack_bwd is None and want_dx):
        return None          # (PN2_WGRAD_DETERMINISTIC=1 keeps the fixed-order per-step kernels: the fused kernel adds dW with L2 reductions)
    if mode == "pooled" and (pooled[2] != 32 or M % 32 != 0 or pooled[0].dtype != torch.float32 or not pooled[0].is_contiguous()):
        return None
    xin = x0 if prev is None else prev.Z
    ldx = xin.shape[1]
    ldd = (_row_ld(st.K, dtype) if prev is not None else x0.shape[1]) if want_dx else 0
    da_mode = {"dA": 0, "dA_masked": 1, "pooled": 2, "dZ": 3}[mode]
    if xin.dtype != torch.bfloat16 or not lib.pn2_mlp_bwd_layer_supported(M, st.K, st.N, ldx, ldd, da_mode, int(prev is not None),
                                                                          int(want_dx), 1):
        return None
    param = getattr(conv, "weight", None)
    dW = _sink(param, (st.N, st.K))
    scratch = None
    nbytes = lib.pn2_mlp_bwd_layer_scratch_bytes(M, st.K, st.N)
    if nbytes:                                   # K % 4 != 0: fixed-order partials, the reduce launch WRITES dW
        scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        if dW is None:
            dW = torch.empty(st.N, st.K, device=dev, dtype=torch.float32)
    elif dW is None:
        dW = torch.zeros(st.N, st.K, device=dev, dtype=torch.float32)      # (a sink is zeroed once per step by its owner)
    dA_prev = torch.empty(M, ldd, device=dev, dtype=dtype) if want_dx else None
    dgamma_prev = dbeta_prev = None
    if prev is not None:
        dgamma_prev = _sink(bn_prev.weight) if bn_prev is not None else None
        dbeta_prev = _sink(bn_prev.bias) if bn_prev is not None else None
        if dgamma_prev is None or dbeta_prev is None:
            dgb = torch.empty(2, st.K, device=dev, dtype=torch.float32)
            dgamma_prev, dbeta_prev = dgb[0], dgb[1]
    L = BwdLayer()
    L.dA, L.ldda, L.da_mode = ptr(cur), (0 if cur is None else cur.shape[1]), da_mode
    if mode == "pooled":
        L.dOut, L.arg, L.nsample = ptr(pooled[0]), ptr(pooled[1]), pooled[2]
    if da_mode != 3:
        L.Z, L.ldz = ptr(st.Z), st.Z.shape[1]
        L.scale, L.shift = ptr(st.scale), ptr(st.shift)
        if st.train:
            L.mean, L.invstd, L.dgamma, L.dbeta = ptr(st.mean), ptr(st.invstd), ptr(st.dgamma), ptr(st.dbeta)
    L.wpack_t = ptr(st.wpack_bwd) if want_dx else None
    L.X, L.ldx = ptr(xin), ldx
    if prev is not None:
        L.prev_scale, L.prev_shift, L.prev_mean, L.prev_invstd = ptr(prev.scale), ptr(prev.shift), ptr(prev.mean), ptr(prev.invstd)
        L.stat_accum, L.ticket = ptr(_stat_accum(dev)), ptr(_ticket(dev))
        L.dgamma_prev, L.dbeta_prev = ptr(dgamma_prev), ptr(dbeta_prev)
    L.dX, L.lddx = ptr(dA_prev), ldd
    L.dW, L.scratch = ptr(dW), ptr(scratch)
    L.M, L.K, L.N = M, st.K, st.N
    call("pn2_mlp_bwd_layer", ctypes.addressof(L), stream())
    keep.append(scratch)
    return dW, dA_prev, dgamma_prev, dbeta_prev


def mlp_backward(layers, x0, K0, M, dout, arg, nsample, need_dx0, convs=None, bns=None, zmax=None):
    """Backward of mlp_forward + tail.  zmax: [G, ld] pre-BatchNorm values of the pooled winners (pn2_bn_relu_max_keep).  dout: [G, C] fp32 pooled gradient with arg-max map `arg`
    (set abstraction) or [M, C] dense gradient (feature propagation: fp32; head: bf16), arg None.
    Returns (per-layer (dW, dbias, dgamma, dbeta), dX0 or None).

    Per layer, from the top: the gradient w.r.t. the layer's activation arrives as "dA" (dense), "dA_masked" (dense, already
    multiplied by the layer's ReLU mask, its BatchNorm gradients known) or "dZ" (gradient w.r.t. the pre-BatchNorm product).
    bf16 rows: ONE fused launch per layer (csrc/bwd_fused.cu: dZ on the fly, data gradient, weight gradient and the
    BatchNorm gradients of the layer below) wherever the library takes the layer; otherwise the per-step kernels
    (reduce / dz / data gradient / weight gradient on a side stream)."""
    lib = load()
    dev, dtype = x0.device, x0.dtype
    L = len(layers)
    grads = [None] * L
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if OVERLAP_WGRAD else None
    side_used = False
    keep = []                 # operands of side-stream kernels stay alive until the streams are joined

    def bn_params(l):
        """(dgamma, dbeta) targets of layer l: the parameters' gradient-sink views when there are any"""
        dgamma = _sink(bns[l].weight) if bns is not None else None
        dbeta = _sink(bns[l].bias) if bns is not None else None
        if dgamma is None or dbeta is None:
            dgb = torch.empty(2, layers[l].N, device=dev, dtype=torch.float32)
            dgamma, dbeta = dgb[0], dgb[1]
        return dgamma, dbeta

    def bn_reduce(l, dA, ldda, pooled):
        """dgamma / dbeta of layer l from the gradient w.r.t. its activation"""
        st = layers[l]
        st.dgamma, st.dbeta = bn_params(l)
        if pooled and zmax is not None:
            # the winners' pre-BatchNorm values were kept by the forward tail: a dense [G, C] column reduction
            call("pn2_bn_relu_bwd_reduce_finalize", ptr(dout), dout.shape[1], dt(dout), ptr(zmax), zmax.shape[1], dt(zmax),
                 ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd), M // nsample, st.N, ptr(_stat_accum(dev)),
                 ptr(_ticket(dev)), ptr(st.dgamma), ptr(st.dbeta), stream())
        elif pooled:
            call("pn2_pool_bn_relu_bwd_reduce_finalize", ptr(dout), ptr(arg), ptr(st.Z), st.Z.shape[1], dt(st.Z),
                 ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd), M // nsample, nsample, st.N, ptr(_stat_accum(dev)),
                 ptr(_ticket(dev)), ptr(st.dgamma), ptr(st.dbeta), stream())
        else:
            call("pn2_bn_relu_bwd_reduce_finalize", ptr(dA), ldda, dt(dA), ptr(st.Z), st.Z.shape[1], dt(st.Z),
                 ptr(st.scale), ptr(st.shift), ptr(st.mean), ptr(st.invstd), M, st.N, ptr(_stat_accum(dev)), ptr(_ticket(dev)),
                 ptr(st.dgamma), ptr(st.dbeta), stream())

    def bn_dz(l, dA, ldda, pooled):
        """dZ of layer l (its dgamma / dbeta known)"""
        st = layers[l]
        C = st.N
        mean_train = ptr(st.mean) if st.train else None
        if pooled:
            dZ = torch.empty(M, _row_ld(C, dtype), device=dev, dtype=dtype)
            call("pn2_pool_bn_relu_bwd_dz", ptr(dout), ptr(arg), ptr(st.Z), st.Z.shape[1], dt(st.Z), ptr(st.scale),
                 ptr(st.shift), mean_train, ptr(st.invstd), ptr(st.dgamma), ptr(st.dbeta), M // nsample, nsample, C, ptr(dZ),
                 dZ.shape[1], dt(dZ), stream())
            return dZ
        if dA is not dout and dA.dtype == dtype and ldda == _row_ld(C, dtype):
            dZ = dA                                   # element-wise update in place (never on autograd's grad)
        else:
            dZ = torch.empty(M, _row_ld(C, dtype), device=dev, dtype=dtype)
        call("pn2_bn_relu_bwd_dz", ptr(dA), ldda, dt(dA), ptr(st.Z), st.Z.shape[1], dt(st.Z), ptr(st.scale),
             ptr(st.shift), mean_train, ptr(st.invstd), ptr(st.dgamma), ptr(st.dbeta), M, C, ptr(dZ), dZ.shape[1],
             dt(dZ), stream())
        return dZ

    # ---- the top layer: its BatchNorm gradients always come from a reduction over the incoming gradient ----
    top = L - 1
    pooled = arg is not None
    bn_reduce(top, dout, dout.shape[1], pooled)
    pool_args = (dout, arg, nsample) if pooled else None
    if not pooled and dout.dtype == torch.bfloat16 and dtype == torch.bfloat16:
        cur, mode = dout, "dA"                        # (the head's bf16 gradient: dZ is formed inside the fused layer kernel)
    elif pooled and FUSED_BWD and FUSED_BWD_POOLED and dtype == torch.bfloat16 and nsample == 32 and M % 32 == 0 and lib.pn2_mlp_bwd_layer_supported(
            M, layers[top].K, layers[top].N, (layers[top - 1].Z.shape[1] if top > 0 else x0.shape[1]),
            (_row_ld(layers[top].K, dtype) if top > 0 else x0.shape[1]) if (top > 0 or need_dx0) else 0, 2, int(top > 0),
            int(top > 0 or need_dx0), 1):
        cur, mode = None, "pooled"                    # dZ of the pooled layer is formed inside the fused kernel from (dOut, arg, Z)
    else:
        cur, mode = bn_dz(top, dout, dout.shape[1], pooled), "dZ"
    dx0 = None
    for l in range(L - 1, -1, -1):
        st = layers[l]
        prev = layers[l - 1] if l > 0 else None
        conv = convs[l] if convs is not None else None
        want_dx = l > 0 or need_dx0
        fused = _fused_bwd_layer(st, prev, x0, M, cur, mode, conv, bns[l - 1] if (bns is not None and l > 0) else None,
                                 want_dx, dev, dtype, keep, pooled=pool_args)
        if fused is None and mode == "pooled":        # (the library declined after all: the per-step kernels)
            cur, mode = bn_dz(l, dout, dout.shape[1], True), "dZ"
        if fused is not None:
            dW, dA, dgp, dbp = fused
            if prev is not None:
                prev.dgamma, prev.dbeta = dgp, dbp
            next_mode = "dA_masked"
        else:
            if mode == "dA_masked":
                cur = bn_dz(l, cur, cur.shape[1], False)
            elif mode == "dA":
                cur = bn_dz(l, cur, cur.shape[1], False)     # (its reduction ran when the gradient was produced / at the top)
            dZ = cur
            if l == 0:
                xin, ldx, sc, sh = x0, x0.shape[1], None, None
            else:
                xin, ldx, sc, sh = prev.Z, prev.Z.shape[1], prev.scale, prev.shift
            if side is not None:          # dZ is complete on the main stream here; the side stream picks it up
                side.wait_stream(main)
                with torch.cuda.stream(side):     # (allocations of this call belong to the side stream)
                    dW = _weight_grad(dZ, xin, ldx, sc, sh, M, st.K, st.N, getattr(conv, "weight", None), dev, side.cuda_stream, keep)
                keep += [dZ, xin, sc, sh]
                side_used = True
            else:
                dW = _weight_grad(dZ, xin, ldx, sc, sh, M, st.K, st.N, getattr(conv, "weight", None), dev, stream(), keep)
            dA = None
            if want_dx:
                ldd = _row_ld(st.K, dtype) if l > 0 else x0.shape[1]
                dA = torch.empty(M, ldd, device=dev, dtype=dtype)
                # (padding columns st.K..ldd of the first layer's input gradient are never read: its consumers --
                # pn2_group_points_bwd, pn2_rows_to_f32, pn2_interp_bwd, _GroupAllFn -- address columns < st.K only)
                if st.wpack_bwd is not None:
                    call("pn2_linear_bwd_data_prepacked", ptr(dZ), dZ.shape[1], dt(dZ), ptr(st.W), M, st.K, st.N, ptr(dA), ldd,
                         dt(dA), ptr(st.wpack_bwd), stream())
                else:
                    wpack = None
                    if dtype == torch.bfloat16:
                        wpack = torch.empty(lib.pn2_linear_wpack_bytes(st.N, st.K), device=dev, dtype=torch.uint8)
                    call("pn2_linear_bwd_data", ptr(dZ), dZ.shape[1], dt(dZ), ptr(st.W), M, st.K, st.N, ptr(dA), ldd, dt(dA),
                         ptr(wpack), stream())
                if l > 0:
                    bn_reduce(l - 1, dA, ldd, False)
            next_mode = "dA"
        if not st.has_bias:
            dbias = None
        elif st.train:      # batch-norm's mean subtraction cancels the conv bias exactly
            dbias = _sink(getattr(conv, "bias", None))        # the sink is zeroed once per step by its owner
            if dbias is None:
                dbias = torch.zeros(st.N, device=dev, dtype=torch.float32)
        else:               # frozen statistics: sum_m dz = scale * dbeta
            dbias = st.scale * st.dbeta
        grads[l] = (dW, dbias, st.dgamma, st.dbeta)
        if l > 0:
            cur, mode = dA, next_mode
        else:
            dx0 = dA
    if side_used:
        main.wait_stream(side)
    del keep
    return grads, dx0


FUSED_EVAL = True         # inference: one fused kernel per set-abstraction level (csrc/sa_fused.cu) when it applies


def _fused_eval_applies(convs, bns, nsample, tensors, D=0):
    if not FUSED_EVAL or ops.rows_dtype() != torch.bfloat16 or nsample != 32 or not 1 <= len(convs) <= 4:
        return False
    if any(_bn_trains(bn) for bn in bns):
        return False
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        return False                      # somebody wants gradients: the per-layer path keeps what backward needs
    widths = [c.out_channels for c in convs]
    # the library's own plan decides (widths, shared / tensor memory): 0 bytes = not supported -> per-layer path
    return load().pn2_sa_fused_eval_workspace_bytes(D, len(widths), (ctypes.c_int * len(widths))(*widths)) > 0


def sa_fused_eval(idx, convs, bns, new_xyz, xyz_r, pts_r):
    """Inference forward of one set-abstraction level in one kernel: gather -> L x relu(bn(conv)) -> max over nsample."""
    lib = load()
    dev = xyz_r.device
    B, N, _ = xyz_r.shape
    S, nsample = idx.shape[1], idx.shape[2]
    feats = None if pts_r is None else ops.as_rows(pts_r)
    D = 0 if feats is None else feats.shape[2]
    L = len(convs)
    widths = [c.out_channels for c in convs]
    c_widths = (ctypes.c_int * L)(*widths)
    ws_bytes = lib.pn2_sa_fused_eval_workspace_bytes(D, L, c_widths)
    if ws_bytes == 0:
        raise ValueError("sa_fused_eval: unsupported level D=%d widths=%s" % (D, widths))
    cached = _eval_cache_get("sa_fused", convs, bns)
    if cached is not None:
        ws, fold = cached
    else:
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        fold = torch.empty(L, 2, max(widths), device=dev, dtype=torch.float32)
    Ws, keep = [], []
    K = 3 + D
    for l, (conv, bn) in enumerate(zip(convs, bns)):
        W = conv.weight.detach().reshape(widths[l], -1)
        W = W if W.is_contiguous() else W.contiguous()
        if W.shape[1] != K or W.dtype != torch.float32:
            raise ValueError("conv weight %s does not match %d input channels (fp32)" % (tuple(conv.weight.shape), K))
        gamma = None if bn.weight is None else bn.weight.detach()
        beta = None if bn.bias is None else bn.bias.detach()
        if cached is None:
            call("pn2_bn_eval_fold", ptr(gamma), ptr(beta), ptr(bn.running_mean), ptr(bn.running_var), float(bn.eps),
                 widths[l], ptr(fold[l, 0]), ptr(fold[l, 1]), stream())
        Ws.append(W)
        keep.append(None if conv.bias is None else conv.bias.detach())
        K = widths[l]
    arr = ctypes.c_void_p * L
    out = torch.empty(B, S, widths[-1], device=dev, dtype=torch.float32)
    sB, sN, sC = xyz_r.stride()
    fB, fN = (0, 0) if feats is None else feats.stride()[:2]
    call("pn2_sa_fused_eval", ptr(xyz_r), sB, sN, sC, ptr(new_xyz), ptr(feats), fB, fN, ptr(idx), B, N, S, nsample, D, L,
         c_widths, None if cached is not None else arr(*[w.data_ptr() for w in Ws]),       # NULL: ws holds the packed images
         arr(*[None if b is None else b.data_ptr() for b in keep]),
         arr(*[fold[l, 0].data_ptr() for l in range(L)]), arr(*[fold[l, 1].data_ptr() for l in range(L)]),
         ptr(out), ptr(ws), stream())
    if cached is None:
        _eval_cache_put("sa_fused", convs, bns, (ws, fold))
    return out


def _flat_params(convs, bns):
    out = []
    for conv, bn in zip(convs, bns):
        out += [conv.weight, conv.bias, bn.weight, bn.bias]
    return out


def _param_grads(grads, convs, needs):
    out = []
    for (dW, dbias, dgamma, dbeta), conv in zip(grads, convs):
        out += [dW.view_as(conv.weight), dbias, dgamma, dbeta]
    return tuple(g if need else None for g, need in zip(out, needs))


class _SetAbstractionFn(torch.autograd.Function):
    """sample_and_group + MLP + max (pointnet2_utils.py:185-200).  perm: optional column
    permutation of the first conv's input channels (the Msg variant concatenates features
    first, :248, while the gather kernel writes [dxyz | feats])."""

    @staticmethod
    def forward(ctx, idx, convs, bns, radius, nsample, new_xyz, xyz_r, pts_r, *params):
        B, N, _ = xyz_r.shape
        S = new_xyz.shape[1]
        dtype = ops.rows_dtype()
        if idx is None:
            idx = ops.query_ball_point(radius, nsample, xyz_r, new_xyz)
        feats = None if pts_r is None else ops.as_rows(pts_r)
        D = 0 if feats is None else feats.shape[2]
        K0 = 3 + D
        x0 = ops.group_rows(xyz_r, new_xyz, feats, idx, _row_ld(K0, dtype), dtype)
        M = B * S * nsample
        layers = mlp_forward(x0, K0, M, convs, bns, _want_backward(ctx))
        last = layers[-1]
        out = torch.empty(B, S, last.N, device=xyz_r.device, dtype=torch.float32)
        arg = torch.empty(B, S, last.N, device=xyz_r.device, dtype=torch.int32)
        zmax = None
        if _want_backward(ctx):
            zmax = torch.empty(B * S, last.Z.shape[1], device=xyz_r.device, dtype=dtype)
            call("pn2_bn_relu_max_keep", ptr(last.Z), last.Z.shape[1], dt(last.Z), ptr(last.scale), ptr(last.shift), B * S,
                 nsample, last.N, ptr(out), ptr(arg), ptr(zmax), zmax.shape[1], stream())
        else:
            call("pn2_bn_relu_max", ptr(last.Z), last.Z.shape[1], dt(last.Z), ptr(last.scale), ptr(last.shift), B * S,
                 nsample, last.N, ptr(out), ptr(arg), stream())
        ctx.state = (layers, x0, K0, M, arg, idx, nsample, (B, N, S, D), convs, bns, zmax)
        return out

    @staticmethod
    def backward(ctx, dout):
        layers, x0, K0, M, arg, idx, nsample, (B, N, S, D), convs, bns, zmax = ctx.state
        ctx.state = None
        need_pts = ctx.needs_input_grad[7] and D > 0
        dout = dout.contiguous().view(B * S, -1)
        grads, dx0 = mlp_backward(layers, x0, K0, M, dout, arg.view(B * S, -1), nsample, need_pts, convs, bns, zmax=zmax)
        dpts = None
        if need_pts:
            dpts = torch.zeros(B, N, D, device=dout.device, dtype=torch.float32)
            call("pn2_group_points_bwd", ptr(dx0), dx0.shape[1], dt(dx0), ptr(idx), B, N, S, nsample, D, ptr(dpts),
                 stream())
        return (None,) * 7 + (dpts,) + _param_grads(grads, convs, ctx.needs_input_grad[8:])


class _GroupAllFn(torch.autograd.Function):
    """group_all=True branch (:141-158, :189-190): one group of all N points, no centring."""

    @staticmethod
    def forward(ctx, convs, bns, x0_f32, *params):
        B, N, K0 = x0_f32.shape
        dtype = ops.rows_dtype()
        ld = _row_ld(K0, dtype)
        x0 = torch.zeros(B * N, ld, device=x0_f32.device, dtype=dtype)
        x0[:, :K0] = x0_f32.reshape(B * N, K0)
        layers = mlp_forward(x0, K0, B * N, convs, bns, _want_backward(ctx))
        last = layers[-1]
        out = torch.empty(B, 1, last.N, device=x0.device, dtype=torch.float32)
        arg = torch.empty(B, 1, last.N, device=x0.device, dtype=torch.int32)
        call("pn2_bn_relu_max", ptr(last.Z), last.Z.shape[1], dt(last.Z), ptr(last.scale), ptr(last.shift), B, N,
             last.N, ptr(out), ptr(arg), stream())
        ctx.state = (layers, x0, K0, arg, (B, N), convs, bns)
        return out

    @staticmethod
    def backward(ctx, dout):
        layers, x0, K0, arg, (B, N), convs, bns = ctx.state
        ctx.state = None
        need = ctx.needs_input_grad[2]
        grads, dx0 = mlp_backward(layers, x0, K0, B * N, dout.contiguous().view(B, -1), arg.view(B, -1), N, need, convs, bns)
        dx = dx0[:, :K0].float().view(B, N, K0) if need else None
        return (None, None, dx) + _param_grads(grads, convs, ctx.needs_input_grad[3:])


class _FeaturePropagationFn(torch.autograd.Function):
    """3-NN inverse-distance interpolation + concat + MLP (pointnet2_utils.py:285-314)."""

    @staticmethod
    def forward(ctx, nn3, convs, bns, xyz1_r, xyz2_r, p1_r, p2_r, *params):
        B, N, _ = xyz1_r.shape
        S = xyz2_r.shape[1]
        dtype = ops.rows_dtype()
        p2 = ops.as_rows(p2_r)
        D2 = p2.shape[2]
        D1 = 0 if p1_r is None else p1_r.shape[2]
        # S == 1 degenerates to weight 1 on the only point (:293-294)
        idx3, w3 = nn3 if nn3 is not None else ops.three_nn(xyz1_r, xyz2_r)
        K0 = D1 + D2
        M = B * N
        x0 = torch.empty(M, _row_ld(K0, dtype), device=p2.device, dtype=dtype)
        pB, pN, pD = (0, 0, 0) if p1_r is None else p1_r.stride()
        call("pn2_interp_concat", ptr(p1_r), pB, pN, pD, ptr(p2), S * D2, D2, 1, ptr(idx3), ptr(w3), B, N, S, D1, D2,
             ptr(x0), x0.shape[1], dt(x0), stream())
        layers = mlp_forward(x0, K0, M, convs, bns, _want_backward(ctx))
        last = layers[-1]
        out = torch.empty(B, N, last.N, device=p2.device, dtype=torch.float32)
        call("pn2_bn_relu", ptr(last.Z), last.Z.shape[1], dt(last.Z), ptr(last.scale), ptr(last.shift), M, last.N,
             ptr(out), stream())
        ctx.state = (layers, x0, K0, M, idx3, w3, (B, N, S, D1, D2), convs, bns)
        return out

    @staticmethod
    def backward(ctx, dout):
        layers, x0, K0, M, idx3, w3, (B, N, S, D1, D2), convs, bns = ctx.state
        ctx.state = None
        need1 = ctx.needs_input_grad[5] and D1 > 0
        need2 = ctx.needs_input_grad[6]
        dout = dout.contiguous().view(M, -1)
        grads, dx0 = mlp_backward(layers, x0, K0, M, dout, None, 1, need1 or need2, convs, bns)
        dp1 = dp2 = None
        if need1:
            dp1 = torch.empty(B, N, D1, device=dout.device, dtype=torch.float32)
            call("pn2_rows_to_f32", ptr(dx0), dx0.shape[1], dt(dx0), M, 0, D1, ptr(dp1), stream())
        if need2:
            dp2 = torch.zeros(B, S, D2, device=dout.device, dtype=torch.float32)
            call("pn2_interp_bwd", ptr(dx0), dx0.shape[1], dt(dx0), ptr(idx3), ptr(w3), B, N, S, D1, D2, ptr(dp2),
                 stream())
        return (None, None, None, None, None, dp1, dp2) + _param_grads(grads, convs, ctx.needs_input_grad[7:])


class _FeaturePropagationHeadFn(torch.autograd.Function):
    """The last feature-propagation level followed by the segmentation head (pointnet2_sem_seg.py:35-39) as ONE chain
    of point-major rows: conv1 + bn1 ride as one more layer of the level's MLP (same tensor-core kernels, same fused
    BatchNorm), then csrc/head.cu does ReLU -> dropout -> conv2 -> log_softmax per point.  Returns log-probabilities
    [B, N, classes].  SURVEY.md 8(f) n2; the modules keep owning every parameter."""

    @staticmethod
    def forward(ctx, nn3, convs, bns, conv2, drop_p, seed, extras, xyz1_r, xyz2_r, p1_r, p2_r, *params):
        """extras: None or a dict.  "loss": (target [B*N] int64, class_weight [NC] fp32 or None) -- the weighted NLL of
        pointnet2_sem_seg.py:47-48 is evaluated inside the head kernels and (logp, loss) is returned; logp is not
        differentiable on that path (the gradient enters through the loss).  "labels": an int64 [B, N] tensor that
        receives argmax(logp, dim=2) from the same kernel (inference; localfunctions.py:400)."""
        loss_args = None if extras is None else extras.get("loss")
        labels = None if extras is None else extras.get("labels")
        B, N, _ = xyz1_r.shape
        S = xyz2_r.shape[1]
        dtype = ops.rows_dtype()
        p2 = ops.as_rows(p2_r)
        D2 = p2.shape[2]
        D1 = 0 if p1_r is None else p1_r.shape[2]
        idx3, w3 = nn3 if nn3 is not None else ops.three_nn(xyz1_r, xyz2_r)
        K0 = D1 + D2
        M = B * N
        x0 = torch.empty(M, _row_ld(K0, dtype), device=p2.device, dtype=dtype)
        pB, pN, pD = (0, 0, 0) if p1_r is None else p1_r.stride()
        call("pn2_interp_concat", ptr(p1_r), pB, pN, pD, ptr(p2), S * D2, D2, 1, ptr(idx3), ptr(w3), B, N, S, D1, D2,
             ptr(x0), x0.shape[1], dt(x0), stream())
        want_bwd = _want_backward(ctx)
        layers = mlp_forward(x0, K0, M, convs, bns, want_bwd)
        last = layers[-1]
        NC = conv2.out_channels
        W2 = conv2.weight.detach().reshape(NC, -1)
        W2 = W2 if W2.is_contiguous() else W2.contiguous()
        b2 = None if conv2.bias is None else conv2.bias.detach()
        logp = torch.empty(B, N, NC, device=p2.device, dtype=torch.float32)
        act = torch.empty(M, last.Z.shape[1], device=p2.device, dtype=torch.bfloat16) if want_bwd else None
        loss = None
        if loss_args is None:
            call("pn2_head_tail_fwd", ptr(last.Z), last.Z.shape[1], ptr(last.scale), ptr(last.shift), ptr(W2), ptr(b2), M,
                 last.N, NC, float(drop_p), ptr(seed), ptr(logp), ptr(act), 0 if act is None else act.shape[1], ptr(labels),
                 stream())
        else:
            target, cw = loss_args
            loss = torch.empty(2, device=p2.device, dtype=torch.float32)          # [loss, sum of the target weights]
            call("pn2_head_tail_loss_fwd", ptr(last.Z), last.Z.shape[1], ptr(last.scale), ptr(last.shift), ptr(W2), ptr(b2),
                 M, last.N, NC, float(drop_p), ptr(seed), ptr(target), ptr(cw), ptr(logp), ptr(act),
                 0 if act is None else act.shape[1], ptr(_stat_accum(p2.device)), ptr(loss), stream())
        ctx.state = (layers, x0, K0, M, idx3, w3, (B, N, S, D1, D2), convs, bns, conv2, W2, float(drop_p), seed, logp, act,
                     loss_args, loss)
        if loss_args is None:
            return logp
        ctx.mark_non_differentiable(logp)
        ctx.set_materialize_grads(False)          # no zero-filled [B, N, classes] gradient for the detached log-probabilities
        return logp, loss[0]

    @staticmethod
    def backward(ctx, dlogp, dloss=None):
        (layers, x0, K0, M, idx3, w3, (B, N, S, D1, D2), convs, bns, conv2, W2, drop_p, seed, logp, act, loss_args,
         loss) = ctx.state
        ctx.state = None
        dev = logp.device
        need1 = ctx.needs_input_grad[9] and D1 > 0
        need2 = ctx.needs_input_grad[10]
        NC, C = W2.shape
        dA = torch.empty(M, act.shape[1], device=dev, dtype=torch.bfloat16)
        if dA.shape[1] != C:
            dA[:, C:].zero_()
        lddl = _round_up(_round_up(NC, 4), 8)
        dl_rows = torch.empty(M, lddl, device=dev, dtype=torch.bfloat16)
        db2 = _sink(conv2.bias)
        if db2 is None:
            db2 = torch.empty(NC, device=dev, dtype=torch.float32)
        if loss_args is None:
            dlogp = dlogp.contiguous().view(M, NC)
            call("pn2_head_tail_bwd", ptr(dlogp), ptr(logp), ptr(W2), M, C, NC, drop_p, ptr(seed), ptr(dA), dA.shape[1],
                 ptr(dl_rows), lddl, ptr(_stat_accum(dev)), ptr(db2), stream())
        else:
            target, cw = loss_args
            if dloss is None:
                raise RuntimeError("fused head loss: backward reached the head without a gradient for the loss")
            dloss = dloss.contiguous().float()
            call("pn2_head_tail_loss_bwd", ptr(logp), ptr(target), ptr(cw), ptr(loss), ptr(dloss), ptr(W2), M, C, NC, drop_p,
                 ptr(seed), ptr(dA), dA.shape[1], ptr(dl_rows), lddl, ptr(_stat_accum(dev)), ptr(db2), stream())
        dW2 = _weight_grad(dl_rows, act, act.shape[1], None, None, M, C, NC, conv2.weight, dev, stream(), [])
        grads, dx0 = mlp_backward(layers, x0, K0, M, dA, None, 1, need1 or need2, convs, bns)
        dp1 = dp2 = None
        if need1:
            dp1 = torch.empty(B, N, D1, device=dev, dtype=torch.float32)
            call("pn2_rows_to_f32", ptr(dx0), dx0.shape[1], dt(dx0), M, 0, D1, ptr(dp1), stream())
        if need2:
            dp2 = torch.zeros(B, S, D2, device=dev, dtype=torch.float32)
            call("pn2_interp_bwd", ptr(dx0), dx0.shape[1], dt(dx0), ptr(idx3), ptr(w3), B, N, S, D1, D2, ptr(dp2),
                 stream())
        n_mlp = 4 * len(convs)
        tail_needs = ctx.needs_input_grad[11 + n_mlp:]
        tail = (dW2.view_as(conv2.weight) if tail_needs[0] else None, db2 if len(tail_needs) > 1 and tail_needs[1] else None)
        return (None,) * 9 + (dp1, dp2) + _param_grads(grads, convs, ctx.needs_input_grad[11:11 + n_mlp]) + tail[:len(tail_needs)]


def _check_module_inputs(xyz, points):
    require_cuda(xyz, "xyz")
    if xyz.dim() != 3 or xyz.shape[1] != 3:
        raise ValueError("xyz must be [B, 3, N], got %s" % (tuple(xyz.shape),))
    if points is not None:
        require_cuda(points, "points")
        if points.dim() != 3 or points.shape[0] != xyz.shape[0] or points.shape[2] != xyz.shape[2]:
            raise ValueError("points must be [B, D, N] matching xyz, got %s" % (tuple(points.shape),))


def _build_mlp(conv_cls, bn_cls, in_channel, widths):
    convs, bns = nn.ModuleList(), nn.ModuleList()
    last = in_channel
    for width in widths:
        convs.append(conv_cls(last, width, 1))
        bns.append(bn_cls(width))
        last = width
    return convs, bns


class PointNetSetAbstraction(nn.Module):
    """Drop-in for pointnet2_utils.py:161-202."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs, self.mlp_bns = _build_mlp(nn.Conv2d, nn.BatchNorm2d, in_channel, mlp)
        self.start_staging = None      # ops.StartIndexStaging when the caller captures CUDA graphs

    def use_static_start_buffers(self, enable=True):
        """Route the FPS start-index draw through pinned staging buffers (CUDA-graph capture)."""
        self.start_staging = "pending" if enable else None

    def _staging(self, B, N, device):
        st = self.start_staging
        if st is None:
            return None
        if st == "pending" or st.B != B or st.N != N:
            st = self.start_staging = ops.StartIndexStaging(B, N, device)
        return st

    def geometry(self, xyz):
        """The index half of sample_and_group (:124-126): FPS centroids and ball-query indices.  They depend on
        xyz only, so a caller may compute them ahead of (and concurrently with) the feature path and hand
        them to forward().  Returns (new_xyz [B,S,3], idx [B,S,nsample] int64)."""
        _check_module_inputs(xyz, None)
        xyz_r = xyz.permute(0, 2, 1)
        _, new_xyz = ops.farthest_point_sample(xyz_r, self.npoint, return_xyz=True,
                                               staging=self._staging(xyz.shape[0], xyz.shape[2], xyz.device))
        idx = ops.query_ball_point(self.radius, self.nsample, xyz_r, new_xyz)
        return new_xyz, idx

    def sample(self, xyz):
        """FPS half of geometry(): xyz [B,3,N] -> new_xyz [B,S,3] (:124-125)."""
        _check_module_inputs(xyz, None)
        _, new_xyz = ops.farthest_point_sample(xyz.permute(0, 2, 1), self.npoint, return_xyz=True,
                                               staging=self._staging(xyz.shape[0], xyz.shape[2], xyz.device))
        return new_xyz

    def ball_query(self, xyz, new_xyz):
        """ball-query half of geometry(): xyz [B,3,N], new_xyz [B,S,3] -> idx [B,S,nsample] int64 (:126)."""
        return ops.query_ball_point(self.radius, self.nsample, xyz.permute(0, 2, 1), new_xyz)

    def forward(self, xyz, points, geometry=None):
        """xyz [B,3,N], points [B,D,N] or None -> new_xyz [B,3,S], new_points [B,D',S]."""
        _note_grad_mode()
        _check_module_inputs(xyz, points)
        xyz_r = xyz.permute(0, 2, 1)
        pts_r = None if points is None else points.permute(0, 2, 1)
        params = _flat_params(self.mlp_convs, self.mlp_bns)
        if self.group_all:
            new_xyz = torch.zeros(xyz.shape[0], 1, 3, device=xyz.device, dtype=xyz.dtype)
            x0 = xyz_r if pts_r is None else torch.cat([xyz_r, pts_r], dim=-1)
            out = _GroupAllFn.apply(self.mlp_convs, self.mlp_bns, x0, *params)
        else:
            if geometry is None:
                _, new_xyz = ops.farthest_point_sample(xyz_r, self.npoint, return_xyz=True,
                                                       staging=self._staging(xyz.shape[0], xyz.shape[2], xyz.device))
                idx = None
            else:
                new_xyz, idx = geometry
            if _fused_eval_applies(self.mlp_convs, self.mlp_bns, self.nsample, [points] + params,
                                   0 if points is None else points.shape[1]):
                if idx is None:
                    idx = ops.query_ball_point(self.radius, self.nsample, xyz_r, new_xyz)
                out = sa_fused_eval(idx, self.mlp_convs, self.mlp_bns, new_xyz, xyz_r, pts_r)
            else:
                out = _SetAbstractionFn.apply(idx, self.mlp_convs, self.mlp_bns, self.radius, self.nsample, new_xyz,
                                              xyz_r, pts_r, *params)
        return new_xyz.permute(0, 2, 1), out.permute(0, 2, 1)


class _PermutedConv(nn.Module):
    """View of a Conv2d whose input channels are read in a different order (no parameters of its own)."""

    def __init__(self, conv, perm):
        super().__init__()
        self.__dict__["_conv"] = conv          # not registered: the owner keeps the parameters
        self.perm = perm
        self.out_channels = conv.out_channels

    @property
    def weight(self):
        return self._conv.weight[:, self.perm]

    @property
    def bias(self):
        return self._conv.bias


class PointNetSetAbstractionMsg(nn.Module):
    """Drop-in for pointnet2_utils.py:205-262 (multi-scale grouping; not used by SSG sem-seg).
    Composed from the same kernels; the reference's feature-first concat order (:248) is
    honoured by reading the first conv's weight columns in permuted order."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for widths in mlp_list:
            convs, bns = _build_mlp(nn.Conv2d, nn.BatchNorm2d, in_channel + 3, widths)
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        _note_grad_mode()
        _check_module_inputs(xyz, points)
        xyz_r = xyz.permute(0, 2, 1)
        pts_r = None if points is None else points.permute(0, 2, 1)
        D = 0 if points is None else points.shape[1]
        _, new_xyz = ops.farthest_point_sample(xyz_r, self.npoint, return_xyz=True)
        outs = []
        for radius, k, convs, bns in zip(self.radius_list, self.nsample_list, self.conv_blocks, self.bn_blocks):
            if D:
                perm = torch.cat([torch.arange(D, D + 3), torch.arange(0, D)]).to(xyz.device)
                first = _PermutedConv(convs[0], perm)
                conv_list = [first] + list(convs)[1:]
                params = [first.weight, first.bias, bns[0].weight, bns[0].bias] + _flat_params(convs, bns)[4:]
            else:
                conv_list, params = list(convs), _flat_params(convs, bns)
            outs.append(_SetAbstractionFn.apply(None, conv_list, list(bns), radius, k, new_xyz, xyz_r, pts_r, *params))
        return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=2).permute(0, 2, 1)


class PointNetFeaturePropagation(nn.Module):
    """Drop-in for pointnet2_utils.py:265-315."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = _build_mlp(nn.Conv1d, nn.BatchNorm1d, in_channel, mlp)

    @staticmethod
    def neighbours(xyz1, xyz2):
        """The index half of forward (:296-302): the three nearest xyz2 points of every xyz1 point and their
        normalised inverse-distance weights; depends on coordinates only (see PointNetSetAbstraction.geometry)."""
        _check_module_inputs(xyz1, None)
        _check_module_inputs(xyz2, None)
        return ops.three_nn(xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1))

    def head_applies(self, conv1, bn1, conv2, points2):
        """Can forward_with_head run this head?  (bf16 rows on CUDA, <= 32 classes, a 1x1 conv of <= 256 channels)"""
        return (ops.rows_dtype() == torch.bfloat16 and points2.is_cuda and isinstance(conv1, nn.Conv1d)
                and isinstance(conv2, nn.Conv1d) and isinstance(bn1, nn.BatchNorm1d) and conv1.kernel_size == (1,)
                and conv2.kernel_size == (1,) and conv1.in_channels == self.mlp_convs[-1].out_channels
                and conv1.out_channels % 32 == 0 and conv1.out_channels <= 256 and conv2.in_channels == conv1.out_channels
                and conv2.out_channels <= 32 and bn1.num_features == conv1.out_channels)

    def forward_with_head(self, xyz1, xyz2, points1, points2, conv1, bn1, dropout, conv2, neighbours=None, loss_target=None,
                          loss_weight=None, labels_out=None):
        """This level followed by `log_softmax(conv2(dropout(relu(bn1(conv1(.))))))` (pointnet2_sem_seg.py:36-38) in one
        chain of rows; returns the log-probabilities [B, N, classes] (already permuted as :39 does).
        loss_target [B*N] int64 (+ loss_weight [classes] fp32 or None): also evaluates F.nll_loss(pred, target, weight)
        (pointnet2_sem_seg.py:47-48) inside the head kernels and returns (log-probabilities, loss); the log-probabilities
        are then detached (the gradient enters through the loss).
        labels_out: a contiguous int64 [B, N] CUDA tensor that receives argmax(log-probabilities, dim=2) from the same kernel
        (not together with loss_target)."""
        _note_grad_mode()
        _check_module_inputs(xyz1, points1)
        _check_module_inputs(xyz2, points2)
        p = float(dropout.p) if (dropout is not None and dropout.training) else 0.0
        seed = None
        if p > 0.0:      # drawn on the device (CUDA generator), read by the kernels at run time: graph replays get fresh masks
            seed = torch.randint(0, 2 ** 31 - 1, (1,), device=points2.device, dtype=torch.int64)
        convs, bns = list(self.mlp_convs) + [conv1], list(self.mlp_bns) + [bn1]
        params = _flat_params(convs, bns) + [conv2.weight] + ([conv2.bias] if conv2.bias is not None else [])
        extras = None
        if labels_out is not None:
            require_cuda(labels_out, "labels_out", torch.int64)
            if loss_target is not None or not labels_out.is_contiguous() or labels_out.numel() != xyz1.shape[0] * xyz1.shape[2]:
                raise ValueError("labels_out must be a contiguous int64 [B, N] tensor (and excludes loss_target)")
            extras = {"labels": labels_out}
        if loss_target is not None:
            require_cuda(loss_target, "loss_target", torch.int64)
            if loss_target.numel() != xyz1.shape[0] * xyz1.shape[2]:
                raise ValueError("loss_target must hold one label per point (%d), got %d" % (xyz1.shape[0] * xyz1.shape[2],
                                                                                         loss_target.numel()))
            if loss_weight is not None:
                require_cuda(loss_weight, "loss_weight")
                if loss_weight.numel() != conv2.out_channels:
                    raise ValueError("loss_weight must hold one weight per class")
                loss_weight = loss_weight.contiguous()
            extras = {"loss": (loss_target.contiguous().view(-1), loss_weight)}
        return _FeaturePropagationHeadFn.apply(
            neighbours, convs, bns, conv2, p, seed, extras, xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1),
            None if points1 is None else points1.permute(0, 2, 1), points2.permute(0, 2, 1), *params)

    def forward(self, xyz1, xyz2, points1, points2, neighbours=None):
        """xyz1 [B,3,N], xyz2 [B,3,S], points1 [B,D1,N] or None, points2 [B,D2,S] -> [B,D',N]."""
        _note_grad_mode()
        _check_module_inputs(xyz1, points1)
        _check_module_inputs(xyz2, points2)
        if points2 is None:
            raise ValueError("points2 is required")
        out = _FeaturePropagationFn.apply(
            neighbours, self.mlp_convs, self.mlp_bns, xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1),
            None if points1 is None else points1.permute(0, 2, 1), points2.permute(0, 2, 1),
            *_flat_params(self.mlp_convs, self.mlp_bns))
        return out.permute(0, 2, 1)
