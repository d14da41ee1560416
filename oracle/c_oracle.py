"""ctypes loader for oracle/pn2_oracle.c -- TEST INFRASTRUCTURE ONLY.

The C file is the machine-independent authority for the distance/index
arithmetic (it spells out every rounding step, see its header).  Functions take
and return numpy arrays.  Reference lines: /root/reference/models/pointnet2_utils.py
:19-40 (square_distance), :63-84 (FPS), :87-107 (ball query), :296-303 (3-NN interp).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpn2oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pn2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def radius_sq(radius):
    """float32(radius ** 2) exactly as `sqrdists > radius ** 2` sees it (:102)."""
    return float(np.float32(float(radius) ** 2))


def square_distance(src, dst):
    src, dst = _f32(src), _f32(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().pn2o_square_distance(_p(src), _p(dst), B, N, M, _p(out))
    return out


def fps(xyz, npoint, start):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    start = np.ascontiguousarray(start, dtype=np.int64)
    out = np.empty((B, npoint), np.int64)
    lib().pn2o_fps(_p(xyz), B, N, int(npoint), _p(start), _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz, return_count=False):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    cnt = np.empty((B, S), np.int32)
    lib().pn2o_ball_query(_p(xyz), _p(new_xyz), B, N, S, ctypes.c_float(radius_sq(radius)),
                          int(nsample), _p(out), _p(cnt))
    return (out, cnt) if return_count else out


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    K3 = min(3, S)
    idx = np.empty((B, N, K3), np.int64)
    dist = np.empty((B, N, K3), np.float32)
    w = np.empty((B, N, K3), np.float32)
    lib().pn2o_three_nn(_p(xyz1), _p(xyz2), B, N, S, _p(idx), _p(dist), _p(w))
    return idx, dist, w


def interpolate(points2, idx3, w3):
    points2 = _f32(points2)
    idx3 = np.ascontiguousarray(idx3, dtype=np.int64)
    w3 = _f32(w3)
    B, S, D = points2.shape
    N, K3 = idx3.shape[1], idx3.shape[2]
    out = np.empty((B, N, D), np.float32)
    lib().pn2o_interpolate(_p(points2), _p(idx3), _p(w3), B, N, S, D, K3, _p(out))
    return out
