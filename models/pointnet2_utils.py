"""Drop-in replacement for the reference's models/pointnet2_utils.py.

Put this repo's root ahead of the reference's on ``sys.path`` (or copy this file
over the reference's) and ``models/pointnet2_sem_seg.py``, ``sem_seg_training.py``
and ``sem_seg_testing.py`` run unchanged: the same names with the same signatures
(/root/reference/models/pointnet2_utils.py :19, :43, :63, :87, :110, :141, :161,
:205, :265), now executed by the sm_100a kernels of libpn2b200.so.  CUDA float32
tensors only; there is no CPU fallback.
"""
import importlib
import os
import sys
from time import time

import numpy as np

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")

square_distance = _pn2.square_distance
index_points = _pn2.index_points
farthest_point_sample = _pn2.farthest_point_sample
query_ball_point = _pn2.query_ball_point
sample_and_group = _pn2.sample_and_group
sample_and_group_all = _pn2.sample_and_group_all
PointNetSetAbstraction = _pn2.PointNetSetAbstraction
PointNetSetAbstractionMsg = _pn2.PointNetSetAbstractionMsg
PointNetFeaturePropagation = _pn2.PointNetFeaturePropagation


def timeit(tag, t):
    """Wall-clock helper kept for API parity (reference :7-9; unused there too)."""
    now = time()
    print("{}: {}s".format(tag, now - t))
    return now


def pc_normalize(pc):
    """Centre a numpy cloud and scale it into the unit sphere (reference :11-17)."""
    centred = pc - np.mean(pc, axis=0)
    return centred / np.max(np.linalg.norm(centred, axis=1))
