"""pn2-b200: B200-native (sm_100a) PointNet++ SSG set-abstraction / feature-propagation hot path.

Public surface == the names of /root/reference/models/pointnet2_utils.py.  The
directory name contains a hyphen, so import it with
``importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")`` or through
the repo-root alias ``import pn2b200``; ``models/pointnet2_utils.py`` re-exports the same
names for the reference's unchanged ``models/pointnet2_sem_seg.py``.
"""
from ._lib import EXPORTED_SYMBOLS, SO_PATH, Pn2Error, launch_count, load
from .modules import PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg
from .ops import (add_vote, draw_rotation_angles, rotate_point_cloud_z_, farthest_point_sample, get_precision, index_points, new_vote_pool, query_ball_point,
                  sample_and_group, sample_and_group_all, sample_training_crops, set_precision, slice_scene, square_distance, three_nn, vote_argmax)
from .sem_seg import get_loss, get_model
from .trainer import FlatAdam, FlatGradients, SemSegPredictor, SemSegTrainer, predict_blocks, predict_scene, shard_range

__all__ = [
    "PointNetSetAbstraction", "PointNetSetAbstractionMsg", "PointNetFeaturePropagation",
    "square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
    "sample_and_group_all", "three_nn", "add_vote", "vote_argmax", "slice_scene", "sample_training_crops", "rotate_point_cloud_z_", "draw_rotation_angles", "new_vote_pool", "predict_scene", "set_precision", "get_precision", "get_model", "get_loss",
    "SemSegTrainer", "SemSegPredictor", "FlatGradients", "FlatAdam", "predict_blocks", "shard_range", "load", "launch_count", "Pn2Error", "SO_PATH", "EXPORTED_SYMBOLS",
]
