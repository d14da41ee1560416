"""The C-ABI library loads, exports every symbol include/pn2b200.h declares, and the ctypes
signatures agree with the header (CPU only: no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    text = open(os.path.join(ROOT, "include", "pn2b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(pn2_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_header_declares_the_hot_path():
    d = _header_decls()
    for name in ("pn2_farthest_point_sample", "pn2_query_ball_point", "pn2_square_distance", "pn2_index_points",
                 "pn2_group_points", "pn2_linear_fwd", "pn2_linear_bwd_data", "pn2_linear_bwd_weight", "pn2_linear_bwd_weight_accum", "pn2_add_vote", "pn2_vote_argmax", "pn2_rotate_z", "pn2_slice_cells", "pn2_crop_members", "pn2_slice_pad", "pn2_slice_rows",
                 "pn2_bn_relu_max", "pn2_three_nn", "pn2_interp_concat", "pn2_interp_bwd", "pn2_last_error"):
        assert name in d


def test_library_exports_every_declared_symbol(pn2):
    from importlib import import_module
    lib_mod = import_module(pn2.__name__ + "._lib")
    assert os.path.exists(pn2.SO_PATH), "libpn2b200.so not built: run __graft_entry__.build()"
    raw = ctypes.CDLL(pn2.SO_PATH)
    decls = _header_decls()
    assert set(decls) == set(lib_mod._SIGNATURES), set(decls) ^ set(lib_mod._SIGNATURES)
    for name, nargs in decls.items():
        assert hasattr(raw, name), name
        assert len(lib_mod._SIGNATURES[name][1]) == nargs, (name, nargs, len(lib_mod._SIGNATURES[name][1]))
    lib = pn2.load()
    assert lib.pn2_version() == 100
    assert lib.pn2_last_error() is not None
    assert lib.pn2_linear_wgrad_scratch_bytes(1 << 20, 12, 32) > 0


def test_sm_budget_is_plain_library_state(pn2):
    """pn2_set_sm_budget: returns the previous value, 0 restores the default, negative values count as 0; fused-backward planning
    and the grid workspace sizes answer without a GPU"""
    lib = pn2.load()
    assert lib.pn2_set_sm_budget(0) in (0,)                       # default: no budget
    assert lib.pn2_set_sm_budget(116) == 0
    assert lib.pn2_set_sm_budget(-5) == 116
    assert lib.pn2_set_sm_budget(0) == 0
    from importlib import import_module
    lib_mod = import_module(pn2.__name__ + "._lib")
    with lib_mod.sm_budget(100):
        assert lib.pn2_set_sm_budget(100) == 100                   # inside: the budget is set
    assert lib.pn2_set_sm_budget(0) == 0                           # ... and restored on exit
    # the layers the fused backward takes (M, K, N, ldx, lddx, da_mode, has_prev, want_dx, want_dw)
    assert lib.pn2_mlp_bwd_layer_supported(1 << 20, 32, 32, 32, 32, 1, 1, 1, 1) == 1
    assert lib.pn2_mlp_bwd_layer_supported(131072, 128, 128, 128, 128, 1, 1, 1, 1) == 1
    assert lib.pn2_mlp_bwd_layer_supported(65536, 131, 128, 136, 136, 1, 0, 1, 1) == 0      # K_ld > 128
    assert lib.pn2_mlp_bwd_layer_supported(16384, 256, 256, 256, 256, 1, 1, 1, 1) == 0
    assert lib.pn2_ball_grid_workspace_bytes(32, 1024) > 32 * 1024 * 20


def test_argument_errors_do_not_need_a_gpu(pn2):
    from importlib import import_module
    lib_mod = import_module(pn2.__name__ + "._lib")
    with pytest.raises(pn2.Pn2Error, match="null pointer"):
        lib_mod.call("pn2_farthest_point_sample", None, 0, 0, 0, 1, 8, 2, None, None, None, None)
    with pytest.raises(pn2.Pn2Error, match="bad sizes"):
        lib_mod.call("pn2_linear_fwd", 1, 4, 0, None, None, 1, None, 8, 8, 8, 1, 8, 0, None, None, None)


def test_product_path_has_no_cpu_fallback(pn2):
    import torch
    with pytest.raises(ValueError, match="CUDA"):
        pn2.farthest_point_sample(torch.rand(1, 16, 3), 4)
    with pytest.raises(ValueError, match="CUDA"):
        pn2.PointNetSetAbstraction(4, 0.2, 4, 6, [8], False)(torch.rand(1, 3, 16), torch.rand(1, 3, 16))
    with pytest.raises(ValueError, match="CUDA"):
        pn2.PointNetFeaturePropagation(8, [8])(torch.rand(1, 3, 16), torch.rand(1, 3, 4), None, torch.rand(1, 8, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "khairil_tum-facade_semantic_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
    assert "oracle" not in open(os.path.join(ROOT, "models", "pointnet2_utils.py")).read()


def test_state_dict_layout_matches_reference_fixture(pn2, golden):
    g = golden("model")
    net = pn2.get_model(18, 3)
    keys = sorted(net.state_dict().keys())
    assert keys == list(g["state_keys"])
    assert [str(tuple(net.state_dict()[k].shape)) for k in keys] == list(g["state_shapes"])
    import torch.nn as nn
    assert isinstance(net.sa1.mlp_convs[0], nn.Conv2d) and isinstance(net.sa1.mlp_bns[0], nn.BatchNorm2d)
    assert isinstance(net.fp1.mlp_convs[0], nn.Conv1d) and isinstance(net.fp1.mlp_bns[0], nn.BatchNorm1d)
