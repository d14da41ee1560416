"""The reference's own models/pointnet2_sem_seg.py, UNCHANGED, on top of this repo's
models/pointnet2_utils.py.  Needs /root/reference (authoring container only); the GPU half
additionally needs a device, so it only runs where both exist."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _import_reference_model():
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("models", "pointnet2_sem_seg")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path[:0] = [ROOT, os.path.join(REF, "models")]     # OUR models/pointnet2_utils.py, THEIR pointnet2_sem_seg.py
    try:
        mod = importlib.import_module("pointnet2_sem_seg")
        utils = importlib.import_module("models.pointnet2_utils")
    finally:
        sys.path[:] = saved_path
    return mod, utils


def test_unchanged_reference_model_builds_on_our_operators(pn2, golden):
    mod, utils = _import_reference_model()
    assert mod.__file__.startswith(REF) and utils.__file__.startswith(ROOT)
    net = mod.get_model(18, 3)
    assert isinstance(net.sa1, pn2.PointNetSetAbstraction) and isinstance(net.fp1, pn2.PointNetFeaturePropagation)
    assert sorted(net.state_dict().keys()) == list(golden("model")["state_keys"])
    ours = pn2.get_model(18, 3)
    ours.load_state_dict(net.state_dict())                  # checkpoints are interchangeable
    for name in ("timeit", "pc_normalize", "square_distance", "index_points", "farthest_point_sample",
                 "query_ball_point", "sample_and_group", "sample_and_group_all", "PointNetSetAbstractionMsg"):
        assert hasattr(utils, name), name


@pytest.mark.gpu
def test_unchanged_reference_model_runs_on_gpu(pn2, golden):
    import numpy as np
    import _inputs as I
    mod, _ = _import_reference_model()
    net = I.randomize_module_(mod.get_model(18, 3), 61).cuda().eval()
    g = golden("model")
    x = I.facade_batch(2, 2048, 9, int(g["facade_seed"])).cuda().transpose(2, 1)
    torch.manual_seed(71)
    with torch.no_grad():
        pred, _ = net(x)
    assert np.abs(pred.cpu().numpy() - g["facade_eval_pred"]).max() < 1e-3
