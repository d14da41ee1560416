"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, mounted read-only in the authoring container) on the seeded
synthetic inputs of tests/_inputs.py.  The fixtures travel to the GPU box; the
reference does not.  Re-run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path = [REF, os.path.join(REF, "models")] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
sys.path.append(os.path.dirname(HERE))          # tests/ for _inputs

import _inputs as I                              # noqa: E402
import models.pointnet2_utils as R              # noqa: E402
import pointnet2_sem_seg as RM                   # noqa: E402

assert R.__file__.startswith(REF), R.__file__
torch.set_num_threads(8)


def i16(t):
    a = t.numpy()
    assert a.min() >= 0 and a.max() < 65536
    return a.astype(np.uint16)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **d)
    print("%-22s %8.1f KB  %d arrays" % (name, os.path.getsize(path) / 1024, len(d)))


def ops_small():
    out = {}
    for tag, xyz in (("cube", I.cube_xyz(2, 256, 0)), ("facade", I.facade_xyz(2, 256, 1))):
        torch.manual_seed(11)
        idx = R.farthest_point_sample(xyz, 64)
        new_xyz = R.index_points(xyz, idx)
        out[tag + "_xyz"] = xyz.numpy()
        out[tag + "_fps"] = i16(idx)
        out[tag + "_new_xyz"] = new_xyz.numpy()
        for r in (0.1, 0.2, 0.4):
            for ns in (8, 32):
                out["%s_ball_r%g_k%d" % (tag, r, ns)] = i16(R.query_ball_point(r, ns, xyz, new_xyz))
        out[tag + "_sqdist"] = R.square_distance(new_xyz, xyz).numpy()
        d, order = R.square_distance(xyz, new_xyz).sort(dim=-1)
        d, order = d[:, :, :3], order[:, :, :3]
        rec = 1.0 / (d + 1e-8)
        w = rec / torch.sum(rec, dim=2, keepdim=True)
        p2 = torch.rand(2, 64, 16, generator=torch.Generator().manual_seed(3))
        out[tag + "_nn_idx"] = i16(order)
        out[tag + "_nn_w"] = w.numpy()
        out[tag + "_p2"] = p2.numpy()
        out[tag + "_interp"] = torch.sum(R.index_points(p2, order) * w.view(2, 256, 3, 1), dim=2).numpy()
        # grouping (sample_and_group) with features
        feats = torch.rand(2, 256, 5, generator=torch.Generator().manual_seed(4))
        torch.manual_seed(12)
        nx, grouped = R.sample_and_group(32, 0.3, 8, xyz, feats)
        out[tag + "_feats"] = feats.numpy()
        out[tag + "_sg_new_xyz"] = nx.numpy()
        out[tag + "_sg_grouped"] = grouped.numpy()
    save("ops_small.npz", out)


def ops_levels():
    """The four (N -> S, radius) levels of the SSG network on a facade batch and a cube cloud."""
    out = {}
    for tag, xyz0 in (("facade", I.facade_batch(2, 4096, 9, 1)[:, :, :3].contiguous()),
                      ("cube", I.cube_xyz(1, 4096, 0))):
        xyz = xyz0
        levels = []
        for lvl, (S, r) in enumerate(((1024, 0.1), (256, 0.2), (64, 0.4), (16, 0.8)), 1):
            torch.manual_seed(20 + lvl)
            idx = R.farthest_point_sample(xyz, S)
            new_xyz = R.index_points(xyz, idx)
            out["%s_l%d_fps" % (tag, lvl)] = i16(idx)
            out["%s_l%d_ball" % (tag, lvl)] = i16(R.query_ball_point(r, 32, xyz, new_xyz))
            levels.append((xyz, new_xyz))
            xyz = new_xyz
        for lvl, (fine, coarse) in enumerate(levels, 1):
            d, order = R.square_distance(fine, coarse).sort(dim=-1)
            d, order = d[:, :, :3], order[:, :, :3]
            rec = 1.0 / (d + 1e-8)
            out["%s_l%d_nn_idx" % (tag, lvl)] = i16(order)
            out["%s_l%d_nn_w" % (tag, lvl)] = (rec / torch.sum(rec, dim=2, keepdim=True)).numpy()
        out[tag + "_xyz_checksum"] = np.float64(I.checksum(xyz0))
    save("ops_levels.npz", out)


def ops_large():
    """Config-3 shape, one cloud: FPS 65536 -> 16384 and ball query r=0.1 k=32 (S-chunked)."""
    xyz = I.cube_xyz(1, 65536, 0)
    t = time.time()
    torch.manual_seed(31)
    idx = R.farthest_point_sample(xyz, 16384)
    print("  reference FPS 65536->16384: %.1f s" % (time.time() - t))
    new_xyz = R.index_points(xyz, idx)
    t = time.time()
    chunks = [R.query_ball_point(0.1, 32, xyz, new_xyz[:, s:s + 512]) for s in range(0, 16384, 512)]
    ball = torch.cat(chunks, 1)
    print("  reference ball query: %.1f s" % (time.time() - t))
    save("ops_large.npz", {"fps": i16(idx), "ball_sha256": np.array(sha(i16(ball))),
                           "ball_head": i16(ball[:, :256]), "ball_tail": i16(ball[:, -256:]),
                           "xyz_checksum": np.float64(I.checksum(xyz))})


def _grads(named, out):
    for n, p in named:
        g = p.grad
        out["grad_stat/" + n] = np.array([g.double().sum().item(), g.double().abs().sum().item(),
                                          g.double().pow(2).sum().sqrt().item()])
        if g.numel() <= 8192:
            out["grad/" + n] = g.numpy().copy()


def modules():
    out = {}
    batch = I.facade_batch(2, 256, 9, 5)                       # [B,N,9]
    x = batch.transpose(2, 1)                                  # [B,9,N] strided like localfunctions.py:209
    xyz = x[:, :3, :]
    # ---- set abstraction -------------------------------------------------
    sa = I.randomize_module_(R.PointNetSetAbstraction(64, 0.3, 16, 9 + 3, [16, 16, 32], False), 41)
    pts = x.clone().requires_grad_(True)
    sa.train()
    torch.manual_seed(51)
    nx, np_ = sa(xyz, pts)
    wsel = torch.rand(np_.shape, generator=torch.Generator().manual_seed(6))
    (np_ * wsel).sum().backward()
    out["sa_train_new_xyz"] = nx.detach().numpy()
    out["sa_train_out"] = np_.detach().numpy()
    out["sa_train_dpoints"] = pts.grad.numpy().copy()
    _grads([("sa." + n, p) for n, p in sa.named_parameters()], out)
    for n, b in sa.named_buffers():
        out["sa_buf_after/" + n] = b.numpy().copy()
    sa.eval()
    torch.manual_seed(52)
    with torch.no_grad():
        nx, np_ = sa(xyz, x)
    out["sa_eval_out"] = np_.numpy()
    # ---- feature propagation --------------------------------------------
    coarse_xyz = nx                                             # [B,3,64]
    p1 = torch.rand(2, 7, 256, generator=torch.Generator().manual_seed(8)).requires_grad_(True)
    p2 = torch.rand(2, 32, 64, generator=torch.Generator().manual_seed(9)).requires_grad_(True)
    fp = I.randomize_module_(R.PointNetFeaturePropagation(7 + 32, [24, 16]), 42)
    fp.train()
    y = fp(xyz, coarse_xyz, p1, p2)
    wsel = torch.rand(y.shape, generator=torch.Generator().manual_seed(10))
    (y * wsel).sum().backward()
    out["fp_coarse_xyz"] = coarse_xyz.numpy()
    out["fp_train_out"] = y.detach().numpy()
    out["fp_train_dp1"] = p1.grad.numpy().copy()
    out["fp_train_dp2"] = p2.grad.numpy().copy()
    _grads([("fp." + n, p) for n, p in fp.named_parameters()], out)
    for n, b in fp.named_buffers():
        out["fp_buf_after/" + n] = b.numpy().copy()
    fp.eval()
    with torch.no_grad():
        out["fp_eval_out"] = fp(xyz, coarse_xyz, p1, p2).numpy()
        out["fp_eval_out_nop1"] = I.randomize_module_(
            R.PointNetFeaturePropagation(32, [24, 16]), 43).eval()(xyz, coarse_xyz, None, p2).numpy()
    save("modules.npz", out)


def _tie_free(x, seeds):
    """True if, on every 3-NN level the network evaluates for input x [B,C,N], the reference's
    sort puts equal distances in ascending index order.  torch.sort(stable=False) on CPU leaves
    the order of exactly tied distances unspecified (it differs from the CUDA path and between
    CPU generations); fixtures avoid inputs whose result depends on it."""
    xyz0 = x[:, :3, :].permute(0, 2, 1).contiguous()
    for seed in seeds:
        torch.manual_seed(seed)
        lv = [xyz0]
        for S in (1024, 256, 64, 16):
            lv.append(R.index_points(lv[-1], R.farthest_point_sample(lv[-1], S)))
        for fine, coarse in zip(lv[:-1], lv[1:]):
            d = R.square_distance(fine, coarse)
            if not torch.equal(d.sort(dim=-1)[1][:, :, :3], d.sort(dim=-1, stable=True)[1][:, :, :3]):
                return False
    return True


def model():
    out = {}
    B, N, NC = 2, 2048, 18
    fseed = next(s for s in range(2, 500) if _tie_free(I.facade_batch(B, N, 9, s).transpose(2, 1), (71, 72)))
    cseed = next(s for s in range(0, 500) if _tie_free(I.cube_batch(B, N, 9, s), (71,)))
    print("  tie-free seeds: facade %d cube %d" % (fseed, cseed))
    out["facade_seed"], out["cube_seed"] = np.int64(fseed), np.int64(cseed)
    net = I.randomize_module_(RM.get_model(NC, 3), 61)
    net.drop1.p = 0.0                                           # dropout off: its mask is generator/device specific
    keys = sorted(net.state_dict().keys())
    out["state_keys"] = np.array(keys)
    out["state_shapes"] = np.array([str(tuple(net.state_dict()[k].shape)) for k in keys])
    out["param_checksum"] = np.float64(sum(I.checksum(v) for v in net.state_dict().values() if v.is_floating_point()))
    for tag, x in (("facade", I.facade_batch(B, N, 9, fseed).transpose(2, 1)), ("cube", I.cube_batch(B, N, 9, cseed))):
        net.eval()
        torch.manual_seed(71)
        with torch.no_grad():
            pred, l4 = net(x)
        out[tag + "_eval_pred"] = pred.numpy()
        out[tag + "_eval_l4"] = l4.numpy()
    # one train step's forward/backward on the facade batch
    x = I.facade_batch(B, N, 9, fseed).transpose(2, 1)
    target = I.labels(B, N, NC, 7)
    weights = torch.linspace(0.5, 1.5, NC)
    net.train()
    net.zero_grad()
    torch.manual_seed(72)
    pred, _ = net(x)
    loss = RM.get_loss()(pred.contiguous().view(-1, NC), target, None, weights)
    loss.backward()
    out["train_pred"] = pred.detach().numpy()
    out["train_loss"] = np.float64(loss.item())
    _grads(list(net.named_parameters()), out)
    for n, b in net.named_buffers():
        if b.numel() <= 512:
            out["buf_after/" + n] = b.numpy().copy()
    save("model.npz", out)


def _reference_localfunctions():
    """Import the UNMODIFIED /root/reference/localfunctions.py.  Its module-level imports pull in IO / plotting packages
    that this image does not have and add_vote does not use (laspy, open3d, h5py, matplotlib, pytz): empty stand-in
    modules satisfy the import statements, nothing of them is ever called."""
    import types
    for name in ("laspy", "open3d", "h5py", "matplotlib", "matplotlib.pyplot", "pytz"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["pytz"], "timezone"):          # localfunctions.py:102 builds a module-level timestamp with it
        import datetime
        sys.modules["pytz"].timezone = lambda name: datetime.timezone.utc
    import localfunctions as LF
    assert LF.__file__.startswith(REF), LF.__file__
    return LF


def votes():
    """localfunctions.py:336-343 add_vote (the reference's own Python loop) + :405 arg-max on seeded block batches:
    overlapping point indices, repeated (point, label) pairs, zero / inf / nan / negative sample weights."""
    LF = _reference_localfunctions()
    out = {}
    for tag, (P, NC, B, N, seed) in {"small": (300, 18, 3, 128, 0), "wide": (1000, 5, 4, 256, 1), "dense": (64, 18, 6, 512, 2)}.items():
        g = np.random.RandomState(seed)
        pool = np.zeros((P, NC))
        batches = []
        for it in range(3):
            idx = g.randint(0, P, size=(B, N)).astype(np.float64)          # the reference's batch arrays are np.zeros -> float64
            lab = g.randint(0, NC, size=(B, N)).astype(np.int64)
            w = g.rand(B, N)
            w[g.rand(B, N) < 0.2] = 0.0
            w[g.rand(B, N) < 0.05] = np.inf
            w[g.rand(B, N) < 0.05] = np.nan
            w[g.rand(B, N) < 0.05] *= -1.0
            pool = LF.add_vote(pool, idx, lab, w)
            batches.append((idx, lab, w))
        out[tag + "_idx"] = np.stack([b[0] for b in batches]).astype(np.int32)
        out[tag + "_lab"] = np.stack([b[1] for b in batches]).astype(np.uint8)
        out[tag + "_w"] = np.stack([b[2] for b in batches])
        out[tag + "_pool"] = pool.astype(np.int32)
        assert np.array_equal(pool, pool.astype(np.int32))
        out[tag + "_labels"] = np.argmax(pool, 1).astype(np.uint8)             # :405
    save("votes.npz", out)


def rotation():
    """provider.rotate_point_cloud_z (provider.py:66-84), unmodified, under a seeded numpy generator; the fixture keeps the
    angles it drew (re-drawn from the same seed) so that the device kernel can be checked without numpy's generator."""
    import provider as P
    assert P.__file__.startswith(REF), P.__file__
    out = {}
    for tag, (B, N, seed) in {"small": (3, 257, 0), "batch": (8, 1024, 1)}.items():
        xyz = I.facade_batch(B, N, 9, 40 + seed)[:, :, :3].numpy().copy()
        np.random.seed(100 + seed)
        rotated = P.rotate_point_cloud_z(xyz)
        np.random.seed(100 + seed)
        angles = np.array([np.random.uniform() * 2 * np.pi for _ in range(B)])
        out[tag + "_xyz"], out[tag + "_angles"], out[tag + "_rotated"] = xyz, angles, rotated
        assert rotated.dtype == np.float32
    save("rotation.npz", out)


def synthetic_scene(P, seed, extent=(6.3, 2.2, 3.0)):
    """A small facade-like scene: float64 coordinates with an offset origin (as LAS files have), duplicated points, points
    exactly on cell borders, labels, and red / green / blue + one non-colour extra feature."""
    g = np.random.RandomState(seed)
    pts = np.stack([g.uniform(0, extent[0], P), np.clip(g.normal(extent[1] / 2, extent[1] / 5, P), 0, extent[1]),
                    g.uniform(0, extent[2], P)], axis=1) + np.array([690000.25, 5335000.5, 512.0])
    pts[: P // 20] = pts[P // 20: 2 * (P // 20)]                       # duplicates
    pts[-8:, 0] = pts[:, 0].min() + 0.5 * np.arange(8)                  # on the cell borders (stride 0.5)
    labels = g.randint(0, 18, P).astype(np.int32)
    extra = [g.randint(0, 256, P).astype(np.float64) for _ in range(3)] + [g.normal(size=P)]
    return pts, labels, extra, ["red", "blue", "green", "planarity"]


def slicer():
    """TestCustomDataset.__getitem__ (sem_seg_testing.py:182-254), unmodified, on synthetic scenes: a dataset object built
    through the reference's own `las_file_list=None` constructor path and filled by hand (no LAS IO)."""
    _reference_localfunctions()
    import types
    if "geofunction" not in sys.modules:
        try:
            import geofunction                                          # noqa: F401
        except Exception:
            sys.modules["geofunction"] = types.ModuleType("geofunction")
            sys.modules["geofunction"].cal_geofeature = None
    import sem_seg_testing as T
    assert T.__file__.startswith(REF), T.__file__
    out = {}
    for tag, (P, seed, bp, extent) in {"wall": (6000, 0, 256, (6.3, 2.2, 3.0)), "sparse": (900, 1, 128, (3.1, 1.6, 2.0))}.items():
        pts, labels, extra, names = synthetic_scene(P, seed, extent)
        ds = T.TestCustomDataset(None, las_file_list=None, num_classes=18, block_points=bp)
        ds.scene_points_list, ds.semantic_labels_list = [pts.copy()], [labels.copy()]
        ds.num_extra_features, ds.feature_name, ds.extra_features_data = len(names), list(names), [extra]
        lw = np.histogram(labels, range(19))[0].astype(np.float32)
        lw = lw / np.sum(lw)
        ds.labelweights = np.power(np.amax(lw) / lw, 1 / 3.0)            # :178-180
        np.random.seed(50 + seed)
        data_room, label_room, sample_weight, index_room = ds[0]
        out[tag + "_points"], out[tag + "_labels"] = pts, labels
        out[tag + "_extra"] = np.stack(extra)
        out[tag + "_labelweights"] = ds.labelweights
        out[tag + "_bp"] = np.array([bp, 50 + seed])
        out[tag + "_data32"] = torch.Tensor(data_room).numpy()           # what the test loop feeds the network (localfunctions.py:394)
        out[tag + "_label"] = label_room.astype(np.int32)
        out[tag + "_weight"] = sample_weight.astype(np.float32)
        out[tag + "_index"] = index_room.astype(np.int32)
    save("slicer.npz", out)


def crops():
    """TrainCustomDataset.__getitem__ (sem_seg_training.py:200-259), unmodified, on a synthetic room: a dataset object built
    through the reference's own `las_file_list=None` constructor path and filled by hand (no LAS IO)."""
    _reference_localfunctions()
    import types
    if "geofunction" not in sys.modules:
        try:
            import geofunction                                          # noqa: F401
        except Exception:
            sys.modules["geofunction"] = types.ModuleType("geofunction")
            sys.modules["geofunction"].cal_geofeature = None
    argv, sys.argv = sys.argv, sys.argv[:1]
    import sem_seg_training as T
    sys.argv = argv
    assert T.__file__.startswith(REF), T.__file__
    out = {}
    for tag, (P, seed, npnt, extent) in {"dense": (16000, 3, 1024, (2.5, 1.5, 3.0)), "thin": (6000, 4, 2048, (2.2, 1.4, 2.0))}.items():
        pts, labels, extra, names = synthetic_scene(P, seed, extent)
        ds = T.TrainCustomDataset(None, num_classes=18, num_point=npnt)
        ds.room_points, ds.room_labels = [pts.copy()], [labels.astype(np.float64)]
        ds.room_coord_min, ds.room_coord_max = [np.amin(pts, axis=0)], [np.amax(pts, axis=0)]
        ds.num_extra_features, ds.feature_name, ds.extra_features_data = len(names), list(names), [extra]
        ds.room_idxs = np.zeros(4, dtype=np.int64)
        np.random.seed(70 + seed)
        feats, labs = zip(*[ds[i] for i in range(4)])
        out[tag + "_points"], out[tag + "_labels"], out[tag + "_extra"] = pts, labels, np.stack(extra)
        out[tag + "_meta"] = np.array([npnt, 70 + seed])
        out[tag + "_features"] = np.stack(feats)                              # float64, as the DataLoader would collate them
        out[tag + "_item_labels"] = np.stack(labs).astype(np.int32)
    save("crops.npz", out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops_small", "ops_levels", "ops_large", "modules", "model", "votes", "rotation", "slicer", "crops"]
    for w in which:
        t = time.time()
        globals()[w]()
        print("  %s done in %.1f s" % (w, time.time() - t))
