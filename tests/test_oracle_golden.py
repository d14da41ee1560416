"""Pins the CPU oracle (oracle/pn2_oracle.c and oracle/pn2_oracle.py) to fixtures produced by
the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

import _inputs as I
from oracle import c_oracle as C
from oracle import pn2_oracle as O

RADII = (0.1, 0.2, 0.4)


@pytest.mark.parametrize("tag", ["cube", "facade"])
def test_c_oracle_small_ops(golden, tag):
    g = golden("ops_small")
    xyz = g[tag + "_xyz"]
    start = I.start_indices(2, 256, 11).numpy()
    fps = C.fps(xyz, 64, start)
    assert np.array_equal(fps, g[tag + "_fps"].astype(np.int64))
    new_xyz = g[tag + "_new_xyz"]
    assert np.array_equal(np.take_along_axis(xyz, fps[:, :, None].repeat(3, 2), 1), new_xyz)
    for r in RADII:
        for k in (8, 32):
            want = g["%s_ball_r%g_k%d" % (tag, r, k)].astype(np.int64)
            assert np.array_equal(C.ball_query(r, k, xyz, new_xyz), want), (r, k)
    assert np.array_equal(C.square_distance(new_xyz, xyz), g[tag + "_sqdist"])      # bit exact
    idx, _, w = C.three_nn(xyz, new_xyz)
    assert np.array_equal(idx, g[tag + "_nn_idx"].astype(np.int64))
    assert np.array_equal(w, g[tag + "_nn_w"])
    assert np.array_equal(C.interpolate(g[tag + "_p2"], idx, w), g[tag + "_interp"])


@pytest.mark.parametrize("tag", ["cube", "facade"])
def test_torch_port_small_ops(golden, tag):
    g = golden("ops_small")
    xyz = torch.from_numpy(g[tag + "_xyz"])
    fps = O.fps(xyz, 64, I.start_indices(2, 256, 11))
    assert np.array_equal(fps.numpy(), g[tag + "_fps"].astype(np.int64))
    new_xyz = O.take_points(xyz, fps)
    for r in RADII:
        for k in (8, 32):
            want = g["%s_ball_r%g_k%d" % (tag, r, k)].astype(np.int64)
            assert np.array_equal(O.ball_query(r, k, xyz, new_xyz).numpy(), want)
    order, w = O.three_nn_weights(xyz, new_xyz)
    assert np.array_equal(order.numpy(), g[tag + "_nn_idx"].astype(np.int64))
    assert np.array_equal(w.numpy(), g[tag + "_nn_w"])
    feats = torch.from_numpy(g[tag + "_feats"])
    torch.manual_seed(12)
    nx, grouped = O.group(32, 0.3, 8, xyz, feats)
    assert np.array_equal(nx.numpy(), g[tag + "_sg_new_xyz"])
    assert np.array_equal(grouped.numpy(), g[tag + "_sg_grouped"])


@pytest.mark.parametrize("tag,B", [("facade", 2), ("cube", 1)])
def test_c_oracle_network_levels(golden, tag, B):
    g = golden("ops_levels")
    xyz = (I.facade_batch(2, 4096, 9, 1)[:, :, :3].contiguous() if tag == "facade" else I.cube_xyz(1, 4096, 0)).numpy()
    assert I.checksum(xyz) == float(g[tag + "_xyz_checksum"])
    levels = []
    for lvl, (S, r) in enumerate(((1024, 0.1), (256, 0.2), (64, 0.4), (16, 0.8)), 1):
        start = I.start_indices(B, xyz.shape[1], 20 + lvl).numpy()
        fps = C.fps(xyz, S, start)
        assert np.array_equal(fps, g["%s_l%d_fps" % (tag, lvl)].astype(np.int64)), lvl
        new_xyz = np.take_along_axis(xyz, fps[:, :, None].repeat(3, 2), 1)
        assert np.array_equal(C.ball_query(r, 32, xyz, new_xyz), g["%s_l%d_ball" % (tag, lvl)].astype(np.int64)), lvl
        levels.append((xyz, new_xyz))
        xyz = new_xyz
    for lvl, (fine, coarse) in enumerate(levels, 1):
        idx, _, w = C.three_nn(fine, coarse)
        assert np.array_equal(idx, g["%s_l%d_nn_idx" % (tag, lvl)].astype(np.int64)), lvl
        assert np.array_equal(w, g["%s_l%d_nn_w" % (tag, lvl)]), lvl


def test_c_oracle_config3_shape(golden):
    """FPS 65536 -> 16384 and ball query r=0.1 k=32 on one cloud (BASELINE.json config 3)."""
    g = golden("ops_large")
    xyz = I.cube_xyz(1, 65536, 0).numpy()
    assert I.checksum(xyz) == float(g["xyz_checksum"])
    fps = C.fps(xyz, 16384, I.start_indices(1, 65536, 31).numpy())
    assert np.array_equal(fps, g["fps"].astype(np.int64))
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].repeat(3, 2), 1)
    ball = C.ball_query(0.1, 32, xyz, new_xyz)
    assert np.array_equal(ball[:, :256], g["ball_head"].astype(np.int64))
    assert np.array_equal(ball[:, -256:], g["ball_tail"].astype(np.int64))
    assert hashlib.sha256(ball.astype(np.uint16).tobytes()).hexdigest() == str(g["ball_sha256"])


def _load_state(module, seed):
    return I.randomize_module_(module, seed)


def test_torch_port_modules(golden):
    g = golden("modules")
    x = I.facade_batch(2, 256, 9, 5).transpose(2, 1)
    xyz = x[:, :3, :]
    sa = _load_state(O.OracleSA(64, 0.3, 16, 12, [16, 16, 32], False), 41).train()
    pts = x.clone().requires_grad_(True)
    torch.manual_seed(51)
    nx, out = sa(xyz, pts)
    wsel = torch.rand(out.shape, generator=torch.Generator().manual_seed(6))
    (out * wsel).sum().backward()
    assert np.array_equal(nx.detach().numpy(), g["sa_train_new_xyz"])
    np.testing.assert_allclose(out.detach().numpy(), g["sa_train_out"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(pts.grad.numpy(), g["sa_train_dpoints"], rtol=0, atol=1e-5)
    for n, p in sa.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g["grad/sa." + n], rtol=1e-4, atol=1e-5, err_msg=n)
    for n, b in sa.named_buffers():
        np.testing.assert_allclose(b.numpy(), g["sa_buf_after/" + n], rtol=1e-6, atol=1e-7, err_msg=n)
    sa.eval()
    torch.manual_seed(52)
    with torch.no_grad():
        _, out = sa(xyz, x)
    np.testing.assert_allclose(out.numpy(), g["sa_eval_out"], rtol=0, atol=1e-6)

    coarse = torch.from_numpy(g["fp_coarse_xyz"])
    p1 = torch.rand(2, 7, 256, generator=torch.Generator().manual_seed(8)).requires_grad_(True)
    p2 = torch.rand(2, 32, 64, generator=torch.Generator().manual_seed(9)).requires_grad_(True)
    fp = _load_state(O.OracleFP(39, [24, 16]), 42).train()
    y = fp(xyz, coarse, p1, p2)
    wsel = torch.rand(y.shape, generator=torch.Generator().manual_seed(10))
    (y * wsel).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g["fp_train_out"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(p1.grad.numpy(), g["fp_train_dp1"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(p2.grad.numpy(), g["fp_train_dp2"], rtol=0, atol=1e-5)


def test_torch_port_model(golden):
    g = golden("model")
    net = _load_state(O.OracleSemSeg(18, 3), 61)
    net.drop1.p = 0.0
    assert sorted(net.state_dict().keys()) == list(g["state_keys"])
    assert [str(tuple(net.state_dict()[k].shape)) for k in sorted(net.state_dict())] == list(g["state_shapes"])
    got = sum(I.checksum(v) for v in net.state_dict().values() if v.is_floating_point())
    assert abs(got - float(g["param_checksum"])) < 1e-6 * float(g["param_checksum"])
    net.eval()
    x = I.facade_batch(2, 2048, 9, int(g["facade_seed"])).transpose(2, 1)
    torch.manual_seed(71)
    with torch.no_grad():
        pred, l4 = net(x)
    np.testing.assert_allclose(pred.numpy(), g["facade_eval_pred"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(l4.numpy(), g["facade_eval_l4"], rtol=0, atol=2e-5)
    net.train()
    torch.manual_seed(72)
    pred, _ = net(x)
    loss = O.nll(pred.contiguous().view(-1, 18), I.labels(2, 2048, 18, 7), torch.linspace(0.5, 1.5, 18))
    assert abs(loss.item() - float(g["train_loss"])) < 1e-5


@pytest.mark.parametrize("tag,NC", [("small", 18), ("wide", 5), ("dense", 18)])
def test_vote_oracle_matches_reference_loop(golden, tag, NC):
    """oracle add_vote / vote_argmax (numpy) == the reference's Python double loop (localfunctions.py:336-343, :405)."""
    v = golden("votes")
    pool = np.zeros((v[tag + "_pool"].shape[0], NC))
    for it in range(v[tag + "_idx"].shape[0]):
        pool = O.add_vote(pool, v[tag + "_idx"][it].astype(np.float64), v[tag + "_lab"][it], v[tag + "_w"][it])
    assert np.array_equal(pool, v[tag + "_pool"].astype(np.float64))
    assert np.array_equal(O.vote_argmax(pool), v[tag + "_labels"].astype(np.int64))


@pytest.mark.parametrize("tag", ["small", "batch"])
def test_rotation_oracle_matches_reference(golden, tag):
    """oracle rotate_z == provider.rotate_point_cloud_z (provider.py:66-84) on the angles the reference drew."""
    v = golden("rotation")
    assert np.array_equal(O.rotate_z(v[tag + "_xyz"], v[tag + "_angles"]), v[tag + "_rotated"])


@pytest.mark.parametrize("tag", ["wall", "sparse"])
def test_slicer_oracle_reproduces_reference(golden, tag):
    """oracle slice_scene == TestCustomDataset.__getitem__ (sem_seg_testing.py:182-254) under the same numpy seed: the four
    arrays exactly (rows after the float32 conversion the test loop applies)."""
    v = golden("slicer")
    bp, seed = int(v[tag + "_bp"][0]), int(v[tag + "_bp"][1])
    np.random.seed(seed)
    data, lab, w, idx, cells = O.slice_scene(v[tag + "_points"].copy(), v[tag + "_labels"], list(v[tag + "_extra"]),
                                             ["red", "blue", "green", "planarity"], v[tag + "_labelweights"], block_points=bp)
    assert np.array_equal(torch.Tensor(data).numpy(), v[tag + "_data32"])
    assert np.array_equal(lab, v[tag + "_label"]) and np.array_equal(idx, v[tag + "_index"])
    assert np.array_equal(w.astype(np.float32), v[tag + "_weight"])
    assert sum(c[2] for c in cells) == data.shape[0] and len(cells) > 4


@pytest.mark.parametrize("tag", ["dense", "thin"])
def test_train_crop_oracle_reproduces_reference(golden, tag):
    """oracle train_crop == TrainCustomDataset.__getitem__ (sem_seg_training.py:200-259) under the same numpy seed."""
    v = golden("crops")
    npnt, seed = int(v[tag + "_meta"][0]), int(v[tag + "_meta"][1])
    pts, labels, extra = v[tag + "_points"], v[tag + "_labels"], list(v[tag + "_extra"])
    np.random.seed(seed)
    for i in range(v[tag + "_features"].shape[0]):
        f, l, c, sel = O.train_crop(pts, labels, extra, ["red", "blue", "green", "planarity"], np.amax(pts, axis=0), num_point=npnt)
        assert np.array_equal(f, v[tag + "_features"][i]) and np.array_equal(l, v[tag + "_item_labels"][i])
