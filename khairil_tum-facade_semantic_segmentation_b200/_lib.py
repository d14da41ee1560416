"""ctypes binding of libpn2b200.so (the C ABI declared in include/pn2b200.h).

This is the only place that touches the shared library.  There is no CPU or
PyTorch fallback: if the library is missing or a call fails, an exception is
raised.  PyTorch is used by the callers of this module for device memory and
streams only; every pointer handed to the library is a raw ``data_ptr()``.
"""
import contextlib
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpn2b200.so")

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}


class Pn2Error(RuntimeError):
    """A libpn2b200 call returned a negative status."""


_i, _l, _f, _p, _z, _d = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double

# name -> (restype, argtypes); mirrors include/pn2b200.h one to one
_SIGNATURES = {
    "pn2_version": (_i, []),
    "pn2_last_error": (ctypes.c_char_p, []),
    "pn2_launch_count": (ctypes.c_ulonglong, []),
    "pn2_set_sm_budget": (ctypes.c_int, [ctypes.c_int]),
    "pn2_square_distance": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "pn2_index_points": (_i, [_p, _l, _l, _l, _i, _i, _i, _p, _l, _p, _p]),
    "pn2_index_points_bwd": (_i, [_p, _p, _i, _i, _i, _l, _p, _p]),
    "pn2_farthest_point_sample": (_i, [_p, _l, _l, _l, _i, _i, _i, _p, _p, _p, _p]),
    "pn2_query_ball_point": (_i, [_p, _l, _l, _l, _p, _l, _l, _l, _i, _i, _i, _f, _i, _p, _p, _p]),
    "pn2_ball_grid_workspace_bytes": (_z, [_i, _i]),
    "pn2_query_ball_point_grid": (_i, [_p, _l, _l, _l, _p, _l, _l, _l, _i, _i, _i, _f, _f, _i, _p, _p, _p, _z, _p]),
    "pn2_group_points": (_i, [_p, _l, _l, _l, _p, _p, _l, _l, _l, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "pn2_group_points_bwd": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p]),
    "pn2_linear_wpack_bytes": (_z, [_i, _i]),
    "pn2_pack_weights": (_i, [_i, _p, _p, _p, _p, _p, _p]),
    "pn2_linear_fwd_prepacked": (_i, [_p, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _i, _i, _p, _p, _p, _p]),
    "pn2_linear_bwd_data_prepacked": (_i, [_p, _i, _i, _p, _l, _i, _i, _p, _i, _i, _p, _p]),
    "pn2_linear_fwd": (_i, [_p, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _i, _i, _p, _p, _p]),
    "pn2_linear_bwd_data": (_i, [_p, _i, _i, _p, _l, _i, _i, _p, _i, _i, _p, _p]),
    "pn2_linear_wgrad_scratch_bytes": (_z, [_l, _i, _i]),
    "pn2_linear_bwd_weight": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _l, _i, _i, _p, _p, _p]),
    "pn2_slice_cells": (_i, [_p, _l, _l, _l, _p, _p, _i, _p, _p, _i, _d, _d, _d, _d, _d, _p, _p, _p, _p, _p]),
    "pn2_crop_members": (_i, [_p, _l, _l, _l, _p, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_slice_pad": (_i, [_p, _p, _p, _p, _p, _l, _i, _p, _p, _p]),
    "pn2_slice_rows": (_i, [_p, _l, _l, _p, _p, _l, _l, _p, _i, _p, _p, _p, _p, _p, _i, _d, _d, _d, _l, _p, _p, _p, _p]),
    "pn2_rotate_z": (_i, [_p, _l, _l, _l, _p, _i, _i, _p]),
    "pn2_add_vote": (_i, [_p, _p, _p, _i, _l, _l, _i, _p, _p, _p]),
    "pn2_vote_argmax": (_i, [_p, _l, _i, _p, _i, _p]),
    "pn2_linear_bwd_weight_accum": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _l, _i, _i, _p, _p]),
    "pn2_mlp_bwd_layer_supported": (_i, [_l, _i, _i, _i, _i, _i, _i, _i, _i]),
    "pn2_mlp_bwd_layer_scratch_bytes": (_z, [_l, _i, _i]),
    "pn2_mlp_bwd_layer": (_i, [_p, _p]),
    "pn2_bn_train_finalize": (_i, [_p, _l, _i, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_bn_eval_fold": (_i, [_p, _p, _p, _p, _f, _i, _p, _p, _p]),
    "pn2_bn_relu_max": (_i, [_p, _i, _i, _p, _p, _l, _i, _i, _p, _p, _p]),
    "pn2_bn_relu_max_keep": (_i, [_p, _i, _i, _p, _p, _l, _i, _i, _p, _p, _p, _i, _p]),
    "pn2_bn_relu": (_i, [_p, _i, _i, _p, _p, _l, _i, _p, _p]),
    "pn2_sa_fused_eval_workspace_bytes": (_z, [_i, _i, _p]),
    "pn2_sa_fused_eval": (_i, [_p, _l, _l, _l, _p, _p, _l, _l, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_bn_relu_bwd_reduce": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _l, _i, _p, _p]),
    "pn2_pool_bn_relu_bwd_reduce": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _p]),
    "pn2_bn_bwd_finalize": (_i, [_p, _i, _p, _p, _p]),
    "pn2_bn_relu_bwd_reduce_finalize": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _l, _i, _p, _p, _p, _p, _p]),
    "pn2_pool_bn_relu_bwd_reduce_finalize": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p]),
    "pn2_bn_relu_bwd_dz": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _l, _i, _p, _i, _i, _p]),
    "pn2_pool_bn_relu_bwd_dz": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _l, _i, _i, _p, _i, _i, _p]),
    "pn2_three_nn": (_i, [_p, _l, _l, _l, _p, _l, _l, _l, _i, _i, _i, _p, _p, _p]),
    "pn2_three_nn_grid": (_i, [_p, _l, _l, _l, _p, _l, _l, _l, _i, _i, _i, _p, _p, _p, _z, _p, _p]),
    "pn2_interp_concat": (_i, [_p, _l, _l, _l, _p, _l, _l, _l, _p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "pn2_interp_bwd": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "pn2_head_tail_fwd": (_i, [_p, _i, _p, _p, _p, _p, _l, _i, _i, _f, _p, _p, _p, _i, _p, _p]),
    "pn2_head_tail_bwd": (_i, [_p, _p, _p, _l, _i, _i, _f, _p, _p, _i, _p, _i, _p, _p, _p]),
    "pn2_head_tail_loss_fwd": (_i, [_p, _i, _p, _p, _p, _p, _l, _i, _i, _f, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    "pn2_head_tail_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _l, _i, _i, _f, _p, _p, _i, _p, _i, _p, _p, _p]),
    "pn2_adam_step": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "pn2_to_rows": (_i, [_p, _l, _l, _l, _i, _l, _i, _p, _l, _i, _p]),
    "pn2_rows_to_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
}



class BnFinalize(ctypes.Structure):
    """pn2_bn_finalize of include/pn2b200.h (a HOST struct of device pointers)."""
    _fields_ = [("ticket", _p), ("gamma", _p), ("beta", _p), ("conv_bias", _p), ("eps", _f), ("momentum", _f),
                ("running_mean", _p), ("running_var", _p), ("scale", _p), ("shift", _p), ("save_mean", _p),
                ("save_invstd", _p), ("num_batches_tracked", _p), ("momentum_dev", _p)]


class BwdLayer(ctypes.Structure):
    """pn2_bwd_layer of include/pn2b200.h (a HOST struct of device pointers)."""
    _fields_ = [("dA", _p), ("ldda", _i), ("da_mode", _i), ("dOut", _p), ("arg", _p), ("nsample", _i), ("Z", _p), ("ldz", _i),
                ("scale", _p), ("shift", _p), ("mean", _p), ("invstd", _p), ("dgamma", _p), ("dbeta", _p),
                ("wpack_t", _p), ("X", _p), ("ldx", _i),
                ("prev_scale", _p), ("prev_shift", _p), ("prev_mean", _p), ("prev_invstd", _p),
                ("dX", _p), ("lddx", _i), ("dW", _p), ("scratch", _p),
                ("stat_accum", _p), ("ticket", _p), ("dgamma_prev", _p), ("dbeta_prev", _p),
                ("M", _l), ("K", _i), ("N", _i)]


EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load libpn2b200.so and declare the signatures.  Fails loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise Pn2Error(
                "libpn2b200.so is not built (%s).  Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or khairil_tum-facade_semantic_segmentation_b200/csrc/build.sh.  There is no CPU fallback." % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


_TIMED = {"name": None, "events": [], "calls": []}


def time_entry_point(name):
    """Bracket every later call of entry point `name` ("*" = every entry point) with CUDA events on the
    stream the call is launched on; bench.py uses this to measure the kernels live.  None switches it off."""
    _TIMED["name"], _TIMED["events"], _TIMED["calls"] = name, [], []


def timed_durations_ms():
    """Per-launch durations (ms) recorded since time_entry_point(); synchronises."""
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in _TIMED["events"]]


def timed_calls():
    """[(entry point, argument tuple, ms)] recorded since time_entry_point(); synchronises."""
    return [(n, a, ms) for (n, a), ms in zip(_TIMED["calls"], timed_durations_ms())]


def call(name, *args):
    """Invoke an int-returning entry point; raise Pn2Error with the library's message on failure.
    The last argument of every such entry point is the stream."""
    lib = load()
    if _TIMED["name"] is not None and (_TIMED["name"] == name or _TIMED["name"] == "*"):
        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        rc = getattr(lib, name)(*args)
        e.record(st)
        _TIMED["events"].append((s, e))
        if name == "pn2_mlp_bwd_layer":      # (its arguments travel in a host struct: keep what the byte model needs)
            L = ctypes.cast(args[0], ctypes.POINTER(BwdLayer)).contents
            args = (int(L.M), int(L.K), int(L.N), int(L.da_mode), int(L.ldda), int(L.ldz), int(L.ldx), int(L.lddx), bool(L.dX), bool(L.X))
        _TIMED["calls"].append((name, args))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise Pn2Error("%s failed (%d): %s" % (name, rc, lib.pn2_last_error().decode()))


def launch_count():
    return int(load().pn2_launch_count())


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt(t):
    return _DT[t.dtype]


def require_cuda(t, name, dtype=torch.float32):
    """The product path is CUDA only; wrong device/dtype is an error, never a fallback."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (pn2-b200 has no CPU path), got device %s" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


@contextlib.contextmanager
def sm_budget(sms):
    """Inside: the persistent kernels size their grids for `sms` SMs (pn2_set_sm_budget); None / 0: no change."""
    if not sms:
        yield
        return
    lib = load()
    prev = lib.pn2_set_sm_budget(int(sms))
    try:
        yield
    finally:
        lib.pn2_set_sm_budget(prev)
