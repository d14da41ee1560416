"""Parity of the CUDA operators (through the C ABI) with the golden fixtures of the unmodified
reference and with the C oracle.  Index outputs are bit-exact; 3-NN weights and interpolated
features are bit-exact too (same fp32 operation order)."""
import hashlib

import numpy as np
import importlib

import pytest
import torch

import _inputs as I
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gather_xyz(xyz, idx):
    return np.take_along_axis(xyz, idx[:, :, None].repeat(3, 2), 1)


@pytest.mark.parametrize("tag", ["cube", "facade"])
def test_small_ops_match_reference_fixture(pn2, golden, tag):
    g = golden("ops_small")
    xyz_h = torch.from_numpy(g[tag + "_xyz"])
    xyz = xyz_h.to(DEV)
    fps, new_xyz = pn2.farthest_point_sample(xyz, 64, start=I.start_indices(2, 256, 11), return_xyz=True)
    assert np.array_equal(fps.cpu().numpy(), g[tag + "_fps"].astype(np.int64))
    assert np.array_equal(new_xyz.cpu().numpy(), g[tag + "_new_xyz"])
    assert np.array_equal(pn2.index_points(xyz, fps).cpu().numpy(), g[tag + "_new_xyz"])
    for r in (0.1, 0.2, 0.4):
        for k in (8, 32):
            got = pn2.query_ball_point(r, k, xyz, new_xyz).cpu().numpy()
            assert np.array_equal(got, g["%s_ball_r%g_k%d" % (tag, r, k)].astype(np.int64)), (r, k)
    assert np.array_equal(pn2.square_distance(new_xyz, xyz).cpu().numpy(), g[tag + "_sqdist"])
    idx, w = pn2.three_nn(xyz, new_xyz)
    assert np.array_equal(idx.cpu().numpy(), g[tag + "_nn_idx"].astype(np.int64))
    assert np.array_equal(w.cpu().numpy(), g[tag + "_nn_w"])
    # sample_and_group consumes the CPU generator exactly like the reference (:75)
    feats = torch.from_numpy(g[tag + "_feats"]).to(DEV)
    torch.manual_seed(12)
    nx, grouped = pn2.sample_and_group(32, 0.3, 8, xyz, feats)
    assert np.array_equal(nx.cpu().numpy(), g[tag + "_sg_new_xyz"])
    assert np.array_equal(grouped.cpu().numpy(), g[tag + "_sg_grouped"])


@pytest.mark.parametrize("tag,B", [("facade", 2), ("cube", 1)])
def test_network_levels_match_reference_fixture(pn2, golden, tag, B):
    g = golden("ops_levels")
    if tag == "facade":
        batch = I.facade_batch(2, 4096, 9, 1).to(DEV)
        xyz = batch[:, :, :3]                      # strided view, as the model sees level 0
    else:
        xyz = I.cube_xyz(1, 4096, 0).to(DEV)
    levels = []
    for lvl, (S, r) in enumerate(((1024, 0.1), (256, 0.2), (64, 0.4), (16, 0.8)), 1):
        torch.manual_seed(20 + lvl)
        fps, new_xyz = pn2.farthest_point_sample(xyz, S, return_xyz=True)
        assert np.array_equal(fps.cpu().numpy(), g["%s_l%d_fps" % (tag, lvl)].astype(np.int64)), lvl
        ball, cnt = pn2.query_ball_point(r, 32, xyz, new_xyz, return_count=True)
        assert np.array_equal(ball.cpu().numpy(), g["%s_l%d_ball" % (tag, lvl)].astype(np.int64)), lvl
        assert int(cnt.min()) >= 1 and int(cnt.max()) <= 32
        levels.append((xyz, new_xyz))
        xyz = new_xyz
    for lvl, (fine, coarse) in enumerate(levels, 1):
        idx, w = pn2.three_nn(fine, coarse)
        assert np.array_equal(idx.cpu().numpy(), g["%s_l%d_nn_idx" % (tag, lvl)].astype(np.int64)), lvl
        assert np.array_equal(w.cpu().numpy(), g["%s_l%d_nn_w" % (tag, lvl)]), lvl


@pytest.mark.parametrize("B,N,npoint", [(1, 1, 1), (2, 7, 7), (2, 7, 12), (3, 33, 20), (1, 64, 16), (16, 64, 64), (3, 100, 37), (16, 256, 64), (2, 1000, 333),
                                         (16, 1024, 256), (16, 4096, 1024), (2, 5000, 500), (2, 8192, 256)])
@pytest.mark.parametrize("kind", ["cube", "facade"])
def test_fps_matches_c_oracle(pn2, B, N, npoint, kind):
    xyz = (I.cube_xyz(B, N, 3) if kind == "cube" else I.facade_xyz(B, N, 4))
    start = I.start_indices(B, N, 5)
    want = C.fps(xyz.numpy(), npoint, start.numpy())
    got, new_xyz = pn2.farthest_point_sample(xyz.to(DEV), npoint, start=start, return_xyz=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(new_xyz.cpu().numpy(), _gather_xyz(xyz.numpy(), want))


def test_fps_degenerate_clouds(pn2):
    # all points identical: every distance is 0, torch.max returns index 0 after the start
    xyz = torch.ones(2, 128, 3)
    start = torch.tensor([5, 77])
    want = C.fps(xyz.numpy(), 8, start.numpy())
    got = pn2.farthest_point_sample(xyz.to(DEV), 8, start=start)
    assert np.array_equal(got.cpu().numpy(), want)
    assert want[0, 0] == 5 and (want[:, 1:] == 0).all()
    assert pn2.farthest_point_sample(xyz.to(DEV), 0, start=start).shape == (2, 0)


def test_fps_cluster_path_config3_shape(pn2, golden):
    """N = 65536 -> 16384: points spread over an 8-CTA cluster, winners exchanged through DSMEM."""
    g = golden("ops_large")
    xyz = I.cube_xyz(1, 65536, 0)
    assert I.checksum(xyz) == float(g["xyz_checksum"])
    torch.manual_seed(31)
    fps, new_xyz = pn2.farthest_point_sample(xyz.to(DEV), 16384, return_xyz=True)
    assert np.array_equal(fps.cpu().numpy(), g["fps"].astype(np.int64))
    ball = pn2.query_ball_point(0.1, 32, xyz.to(DEV), new_xyz).cpu().numpy()
    assert np.array_equal(ball[:, :256], g["ball_head"].astype(np.int64))
    assert np.array_equal(ball[:, -256:], g["ball_tail"].astype(np.int64))
    assert hashlib.sha256(ball.astype(np.uint16).tobytes()).hexdigest() == str(g["ball_sha256"])


@pytest.mark.parametrize("B,N,npoint", [(3, 20000, 700), (2, 9000, 300), (1, 131072, 64)])
def test_fps_cluster_path_ragged_sizes(pn2, B, N, npoint):
    xyz = I.facade_xyz(B, N, 9)
    start = I.start_indices(B, N, 6)
    want = C.fps(xyz.numpy(), npoint, start.numpy())
    got = pn2.farthest_point_sample(xyz.to(DEV), npoint, start=start)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("B,N,S,r,k", [(2, 64, 16, 0.8, 32), (16, 1024, 256, 0.2, 32), (2, 3000, 130, 0.15, 64),
                                        (1, 4096, 1024, 0.1, 32), (2, 500, 50, 0.05, 5)])
def test_ball_query_matches_c_oracle(pn2, B, N, S, r, k):
    xyz = I.facade_xyz(B, N, 8)
    new_xyz = xyz[:, torch.randperm(N, generator=torch.Generator().manual_seed(1))[:S]].contiguous()
    want, wcnt = C.ball_query(r, k, xyz.numpy(), new_xyz.numpy(), return_count=True)
    got, cnt = pn2.query_ball_point(r, k, xyz.to(DEV), new_xyz.to(DEV), return_count=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(cnt.cpu().numpy(), wcnt)


def test_ball_query_empty_ball_yields_N(pn2):
    xyz = I.cube_xyz(1, 200, 1)
    far = torch.full((1, 3, 3), 50.0)
    got = pn2.query_ball_point(0.1, 4, xyz.to(DEV), far.to(DEV)).cpu().numpy()
    assert (got == 200).all()                       # what the reference's sort leaves behind
    assert np.array_equal(got, C.ball_query(0.1, 4, xyz.numpy(), far.numpy()))


def _ball_both_paths(pn2, r, k, xyz, new_xyz):
    """(cell-grid result, index-order-scan result) of query_ball_point, with counts"""
    ops = __import__("importlib").import_module("khairil_tum-facade_semantic_segmentation_b200.ops")
    saved, saved_max = ops.BALL_GRID, ops._BALL_GRID_MAX_N
    try:
        ops.BALL_GRID, ops._BALL_GRID_MAX_N = True, 409600
        g = pn2.query_ball_point(r, k, xyz.to(DEV), new_xyz.to(DEV), return_count=True)
        ops.BALL_GRID = False
        s = pn2.query_ball_point(r, k, xyz.to(DEV), new_xyz.to(DEV), return_count=True)
    finally:
        ops.BALL_GRID, ops._BALL_GRID_MAX_N = saved, saved_max
    return g, s


@pytest.mark.parametrize("kind", ["facade", "cube", "lattice_at_r", "duplicates", "far_queries", "flat"])
@pytest.mark.parametrize("r,k", [(0.1, 32), (0.2, 16), (0.4, 64)])
def test_ball_query_cell_grid_equals_index_order_scan(pn2, kind, r, k):
    """csrc/ballgrid.cu must reproduce the radius scan (= the reference's mask + sort, pointnet2_utils.py:96-106) bit for
    bit, including the cases DESIGN.md named: a lattice with neighbours at EXACTLY the radius (the fp32 expanded-form
    distance decides, not geometry), duplicated points, queries far outside the cloud (empty balls), degenerate boxes."""
    g = torch.Generator().manual_seed(5)
    B, N, S = 3, 3000, 700
    if kind == "facade":
        xyz = I.facade_xyz(B, N, 8)
    elif kind == "cube":
        xyz = I.cube_xyz(B, N, 3)
    elif kind == "lattice_at_r":
        xyz = torch.round(torch.rand(B, N, 3, generator=g) * 10) * r                  # spacing exactly r: ties at the radius
    elif kind == "duplicates":
        xyz = I.cube_xyz(B, N, 4)
        xyz[:, N // 2:] = xyz[:, :N - N // 2]                                          # every point twice
    elif kind == "far_queries":
        xyz = I.cube_xyz(B, N, 6)
    else:
        xyz = I.cube_xyz(B, N, 7)
        xyz[:, :, 1] = 0.25                                                            # a plane: one cell thick
    new_xyz = xyz[:, torch.randperm(N, generator=torch.Generator().manual_seed(1))[:S]].contiguous()
    if kind == "far_queries":
        new_xyz = new_xyz.clone()
        new_xyz[:, ::3] += 5.0
    (gi, gc), (si, sc) = _ball_both_paths(pn2, r, k, xyz, new_xyz)
    assert torch.equal(gi, si) and torch.equal(gc, sc)
    want, wcnt = C.ball_query(r, k, xyz.numpy(), new_xyz.numpy(), return_count=True)
    assert np.array_equal(gi.cpu().numpy(), want) and np.array_equal(gc.cpu().numpy(), wcnt)


def test_ball_query_cell_grid_falls_back_where_rounding_could_reach_the_margin(pn2):
    """un-centred coordinates (z ~ 512, as raw LAS heights would be): the reference's fp32 distance is mostly rounding
    noise at r = 0.1 and accepts points far outside the ball; NaN coordinates count as inside (:102 compares false).
    The grid path must hand such clouds to the index-order scan and still match it exactly."""
    xyz = I.facade_xyz(2, 2048, 9)
    xyz[0, :, 2] += 512.0
    xyz[1, 7, 0] = float("nan")
    new_xyz = xyz[:, ::8].contiguous()
    (gi, gc), (si, sc) = _ball_both_paths(pn2, 0.1, 32, xyz, new_xyz)
    assert torch.equal(gi, si) and torch.equal(gc, sc)
    want = C.ball_query(0.1, 32, xyz[:1].numpy(), new_xyz[:1].numpy())
    assert np.array_equal(gi[:1].cpu().numpy(), want)


def test_ball_query_cell_grid_config3_shape(pn2, golden):
    """BASELINE.json configs[2]: 65536 points, r = 0.1, nsample = 32 against the fixture of the unmodified reference"""
    g = golden("ops_large")
    xyz = I.cube_xyz(1, 65536, 0)
    idx = torch.from_numpy(g["fps"].astype(np.int64))
    new_xyz = torch.gather(xyz, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    (gi, _), (si, _) = _ball_both_paths(pn2, 0.1, 32, xyz, new_xyz)
    assert torch.equal(gi, si)
    assert np.array_equal(gi[:, :256].cpu().numpy(), g["ball_head"].astype(np.int64))
    assert np.array_equal(gi[:, -256:].cpu().numpy(), g["ball_tail"].astype(np.int64))


@pytest.mark.parametrize("B,N,S", [(2, 300, 1), (2, 300, 2), (2, 300, 3), (4, 1024, 256), (1, 4096, 1024), (2, 100, 2500)])
def test_three_nn_matches_c_oracle(pn2, B, N, S):
    fine, coarse = I.facade_xyz(B, N, 2), I.facade_xyz(B, S, 3, dup_frac=0.3 if S > 8 else 0.0)
    widx, _, ww = C.three_nn(fine.numpy(), coarse.numpy())
    idx, w = pn2.three_nn(fine.to(DEV), coarse.to(DEV))
    k3 = min(3, S)
    assert np.array_equal(idx.cpu().numpy()[:, :, :k3], widx)
    assert np.array_equal(w.cpu().numpy()[:, :, :k3], ww)
    assert (w.cpu().numpy()[:, :, k3:] == 0).all()


def test_three_nn_exact_ties_take_the_lower_index(pn2):
    """Equal distances: the reference's torch.sort(stable=False) leaves their order unspecified
    (its CPU path is an unstable vectorised sort, its CUDA path a stable radix sort).  This
    implementation and the C oracle define it as the stable order: lower index first."""
    fine = torch.zeros(1, 1, 3)
    coarse = torch.tensor([[[0., 0., 2.], [1., 0., 0.], [-1., 0., 0.], [0., 1., 0.], [0., -1., 0.]]])
    idx, w = pn2.three_nn(fine.to(DEV), coarse.to(DEV))
    assert idx.cpu().tolist() == [[[1, 2, 3]]]
    np.testing.assert_allclose(w.cpu().numpy(), 1.0 / 3.0, rtol=1e-6)
    widx, _, ww = C.three_nn(fine.numpy(), coarse.numpy())
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(w.cpu().numpy(), ww)


def _nn3_both_paths(pn2, fine, coarse):
    ops = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.ops")
    count = torch.zeros(1, dtype=torch.int32, device=DEV)
    try:
        ops.THREE_NN_GRID = True
        gi, gw = ops.three_nn(fine, coarse, fallback_count=count)
    finally:
        ops.THREE_NN_GRID = False
    si, sw = ops.three_nn(fine, coarse)
    return (gi, gw), (si, sw), int(count.item())


@pytest.mark.parametrize("case", ["facade_levels", "cube", "plane", "queries_outside", "lattice_ties", "duplicates", "offset_512", "nan"])
def test_three_nn_grid_equals_the_index_order_scan(pn2, case):
    """csrc/ballgrid.cu nn3_grid_kernel vs three_nn_kernel: indices AND weights identical, bit for bit, whatever the cloud --
    the grid search only accepts a list it can prove final and scans everything otherwise"""
    g = torch.Generator().manual_seed(5)
    B = 3
    if case == "facade_levels":            # the network's own pairs: coarse = FPS samples of fine (fp1: 4096 / 1024, fp2: 1024 / 256)
        fine = I.facade_xyz(B, 4096, 7).to(DEV)
        idx = pn2.farthest_point_sample(fine, 1024, start=I.start_indices(B, 4096, 1).to(DEV))
        coarse = torch.gather(fine, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    elif case == "cube":
        fine, coarse = I.cube_xyz(B, 3000, 1).to(DEV), I.cube_xyz(B, 700, 2).to(DEV)
    elif case == "plane":                  # a degenerate box: the density estimate must not break anything
        fine, coarse = torch.rand(B, 2000, 3, generator=g), torch.rand(B, 512, 3, generator=g)
        fine[:, :, 2] = 0.25
        coarse[:, :, 2] = 0.25
        fine, coarse = fine.to(DEV), coarse.to(DEV)
    elif case == "queries_outside":        # fine points far outside the coarse cloud's box (clamped cells)
        coarse = (torch.rand(B, 600, 3, generator=g) * 0.2 + 0.4).to(DEV)
        fine = (torch.rand(B, 2000, 3, generator=g) * 3.0 - 1.0).to(DEV)
    elif case == "lattice_ties":           # a regular lattice: many exactly equal distances, the lower index must win
        ax = torch.arange(8, dtype=torch.float32) * 0.125
        lat = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, 512, 3)
        coarse = lat[:, torch.randperm(512, generator=g)].expand(B, -1, -1).contiguous().to(DEV)
        fine = (lat + 0.0625).expand(B, -1, -1).contiguous().to(DEV)          # cell centres: 8 equidistant corners each
    elif case == "duplicates":
        fine, coarse = I.facade_xyz(B, 1500, 2).to(DEV), I.facade_xyz(B, 400, 3, dup_frac=0.5).to(DEV)
    elif case == "offset_512":             # un-centred coordinates: the build refuses the grid (rounding error > cell margin)
        fine, coarse = I.facade_xyz(B, 1024, 2).to(DEV) + 512.0, I.facade_xyz(B, 300, 3).to(DEV) + 512.0
    else:                                   # NaN in the coarse cloud of one batch element
        fine, coarse = I.facade_xyz(B, 1024, 2).to(DEV), I.facade_xyz(B, 300, 3).to(DEV)
        coarse[1, 17, 1] = float("nan")
    (gi, gw), (si, sw), fallbacks = _nn3_both_paths(pn2, fine, coarse)
    assert torch.equal(gi, si), case
    assert torch.equal(torch.nan_to_num(gw, nan=-7.0), torch.nan_to_num(sw, nan=-7.0)), case
    total = fine.shape[0] * fine.shape[1]
    if case in ("facade_levels", "cube", "duplicates"):
        assert fallbacks < 0.25 * total, (case, fallbacks, total)       # the grid is doing the work
    if case == "offset_512":
        assert fallbacks == total
    print("three_nn grid [%s]: %d of %d queries took the full scan" % (case, fallbacks, total))


def test_index_points_forward_backward(pn2):
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(3, 50, 7, generator=g)
    idx = torch.randint(0, 50, (3, 20, 4), generator=g)
    a = pts.clone().to(DEV).requires_grad_(True)
    out = pn2.index_points(a, idx.to(DEV))
    ref = pts.clone().requires_grad_(True)
    want = ref[torch.arange(3).view(3, 1, 1), idx]
    assert np.array_equal(out.detach().cpu().numpy(), want.detach().numpy())
    wsel = torch.rand(want.shape, generator=g)
    (out * wsel.to(DEV)).sum().backward()
    (want * wsel).sum().backward()
    np.testing.assert_allclose(a.grad.cpu().numpy(), ref.grad.numpy(), rtol=1e-6, atol=1e-6)
    # strided (channel-first view) source
    cf = pts.permute(0, 2, 1).contiguous().to(DEV).permute(0, 2, 1)
    assert np.array_equal(pn2.index_points(cf, idx.to(DEV)).cpu().numpy(), want.detach().numpy())


def test_wrong_dtype_or_shape_raises(pn2):
    with pytest.raises(TypeError):
        pn2.farthest_point_sample(torch.rand(1, 16, 3, device=DEV).double(), 4)
    with pytest.raises(ValueError):
        pn2.farthest_point_sample(torch.rand(1, 3, 16, device=DEV), 4)
    with pytest.raises(TypeError):
        pn2.index_points(torch.rand(1, 16, 3, device=DEV), torch.zeros(1, 4, device=DEV, dtype=torch.int32))
