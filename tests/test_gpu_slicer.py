"""SURVEY.md 8(f) n3 -- the sliding-block slicer on the device (csrc/slicer.cu, ops.slice_scene) against the reference's
TestCustomDataset.__getitem__ (/root/reference/sem_seg_testing.py:182-254) through its fixture and the numpy restatement
in oracle/pn2_oracle.py.  Deterministic parts are exact: the cells and their order, each cell's member set and block count,
and every row bit for bit given (cell, point).  Which members pad a cell and the order inside a cell are random in the
reference too; they are checked as properties (every member present, padding drawn from the members, without repetition
when the reference samples without replacement)."""
import numpy as np
import pytest
import torch

from oracle import pn2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAMES = ["red", "blue", "green", "planarity"]


def _expected_rows(points, extra, names, idx, cell_ids, grid_x, block_size=1.0, stride=0.5):
    """rows of sem_seg_testing.py:216-241 for the given (slot -> point, slot -> cell), float64 then float32"""
    cmin, cmax = points.min(0), points.max(0)
    ix, iy = cell_ids % grid_x, cell_ids // grid_x
    s_x = np.minimum(cmin[0] + ix * stride + block_size, cmax[0]) - block_size
    s_y = np.minimum(cmin[1] + iy * stride + block_size, cmax[1]) - block_size
    p = points[idx]
    cols = [p[:, 0] - (s_x + block_size / 2.0), p[:, 1] - (s_y + block_size / 2.0), p[:, 2],
            p[:, 0] / cmax[0], p[:, 1] / cmax[1], p[:, 2] / cmax[2]]
    for e, nm in zip(extra, names):
        cols.append(e[idx] / 255 if nm in ("red", "blue", "green") else e[idx])
    return torch.Tensor(np.stack(cols, axis=1)).numpy()


def _check(pn2, points, labels, extra, lw, bp, seed):
    np.random.seed(seed)
    _, _, _, _, cells = O.slice_scene(points.copy(), labels, extra, NAMES, lw, block_points=bp)
    gen = torch.Generator(device=DEV).manual_seed(seed)
    data, lab, w, idx = pn2.slice_scene(torch.from_numpy(points).to(DEV), torch.from_numpy(labels.astype(np.int64)).to(DEV),
                                        torch.from_numpy(np.stack(extra)).to(DEV), NAMES, torch.from_numpy(lw.astype(np.float32)),
                                        block_points=bp, generator=gen)
    data, lab, w, idx = data.cpu().numpy(), lab.cpu().numpy(), w.cpu().numpy(), idx.cpu().numpy()
    assert data.shape == (sum(c[2] for c in cells), bp, 6 + len(NAMES)) and idx.shape == data.shape[:2]
    grid_x = int(np.ceil(float(points[:, 0].max() - points[:, 0].min() - 1.0) / 0.5) + 1)
    b0, cell_of_block = 0, []
    for cell, members, nblk in cells:                                   # same cells, same order, same block counts
        got = idx[b0:b0 + nblk].ravel()
        uniq, cnt = np.unique(got, return_counts=True)
        assert np.array_equal(uniq, np.sort(members)), cell              # every member, nothing but members
        if got.size - members.size <= members.size:                      # the reference pads without replacement here
            assert cnt.max() <= 2, cell
        cell_of_block += [cell] * nblk
        b0 += nblk
    cell_ids = np.repeat(np.array(cell_of_block), bp)
    want = _expected_rows(points, extra, NAMES, idx.ravel(), cell_ids, grid_x)
    assert np.array_equal(data.reshape(-1, data.shape[2]), want)          # bit-exact rows
    assert np.array_equal(lab.ravel(), labels[idx.ravel()])
    assert np.array_equal(w.ravel(), lw.astype(np.float32)[labels[idx.ravel()]])
    return data, idx, cells


@pytest.mark.parametrize("tag", ["wall", "sparse"])
def test_slice_scene_matches_reference_structure(pn2, golden, tag):
    v = golden("slicer")
    bp, seed = int(v[tag + "_bp"][0]), int(v[tag + "_bp"][1])
    data, idx, cells = _check(pn2, v[tag + "_points"], v[tag + "_labels"], list(v[tag + "_extra"]), v[tag + "_labelweights"], bp, seed)
    # against the reference's own output: same number of blocks, and per cell the same SET of points
    ref_idx = v[tag + "_index"]
    assert ref_idx.shape == idx.shape
    b0 = 0
    for cell, members, nblk in cells:
        assert np.array_equal(np.unique(ref_idx[b0:b0 + nblk]), np.unique(idx[b0:b0 + nblk]))
        b0 += nblk
    # the order inside a cell really is shuffled (not the fill order) and differs from seed to seed
    data2, idx2, _ = _check(pn2, v[tag + "_points"], v[tag + "_labels"], list(v[tag + "_extra"]), v[tag + "_labelweights"], bp, seed + 1)
    assert not np.array_equal(idx, idx2)


def test_slice_scene_larger_random_scene_and_degenerate_inputs(pn2):
    g = np.random.RandomState(3)
    P = 200_000
    pts = np.stack([g.uniform(0, 12.7, P), g.normal(0, 0.4, P), g.uniform(0, 9, P)], axis=1) + np.array([1000.0, -50.0, 3.0])
    labels = g.randint(0, 18, P).astype(np.int32)
    extra = [g.randint(0, 256, P).astype(np.float64) for _ in range(3)] + [g.normal(size=P)]
    lw = g.rand(18) + 0.5
    _check(pn2, pts, labels, extra, lw, 4096, 11)
    # a scene narrower than one block in y has grid_y = ceil(negative) + 1 <= 0 cells in the reference: no blocks
    thin = pts.copy()
    thin[:, 1] = 5.0 + 0.2 * g.rand(P)
    data, lab, w, idx = pn2.slice_scene(torch.from_numpy(thin).to(DEV))
    assert data.shape == (0, 4096, 6) and idx.shape == (0, 4096)
    np.random.seed(0)
    assert O.slice_scene(thin, labels, [], [], lw)[4] == []
    with pytest.raises(TypeError):
        pn2.slice_scene(torch.from_numpy(pts).float().to(DEV))
    with pytest.raises(ValueError):
        pn2.slice_scene(torch.from_numpy(pts).to(DEV), extra=torch.zeros(2, P, dtype=torch.float64, device=DEV), extra_names=["red"])


def _check_crops(pn2, pts, labels, extra, npnt, batch, seed):
    """structure and rows of ops.sample_training_crops against the reference's rules (sem_seg_training.py:200-259)"""
    gen = torch.Generator(device=DEV).manual_seed(seed)
    d_pts = torch.from_numpy(pts).to(DEV)
    feats, lab, centre_idx, sel = pn2.sample_training_crops(d_pts, torch.from_numpy(labels.astype(np.int64)).to(DEV), batch,
                                                            torch.from_numpy(np.stack(extra)).to(DEV), NAMES, num_point=npnt,
                                                            generator=gen)
    feats, lab, centre_idx, sel = feats.cpu().numpy(), lab.cpu().numpy(), centre_idx.cpu().numpy(), sel.cpu().numpy()
    assert feats.shape == (batch, npnt, 6 + len(NAMES)) and sel.shape == (batch, npnt)
    cmax = pts.max(0)
    for b in range(batch):
        c = pts[centre_idx[b]]
        lo, hi = c - [0.5, 0.5, 0], c + [0.5, 0.5, 0]
        members = np.where((pts[:, 0] >= lo[0]) & (pts[:, 0] <= hi[0]) & (pts[:, 1] >= lo[1]) & (pts[:, 1] <= hi[1]))[0]
        assert members.size > 1024                                          # :214 the accepted block is populated
        assert np.isin(sel[b], members).all()                                # only points of the block
        if members.size >= npnt:
            assert np.unique(sel[b]).size == npnt                            # :218 replace=False
        p = pts[sel[b]]
        cols = [p[:, 0] - c[0], p[:, 1] - c[1], p[:, 2], p[:, 0] / cmax[0], p[:, 1] / cmax[1], p[:, 2] / cmax[2]]
        for e, nm in zip(extra, NAMES):
            cols.append(e[sel[b]] / 255 if nm in ("red", "blue", "green") else e[sel[b]])
        assert np.array_equal(feats[b], torch.Tensor(np.stack(cols, axis=1)).numpy())   # bit-exact rows
        assert np.array_equal(lab[b], labels[sel[b]])
    return centre_idx, sel


@pytest.mark.parametrize("tag", ["dense", "thin"])
def test_training_crops_follow_reference_rules(pn2, golden, tag):
    v = golden("crops")
    npnt = int(v[tag + "_meta"][0])
    pts, labels, extra = v[tag + "_points"], v[tag + "_labels"], list(v[tag + "_extra"])
    c1, s1 = _check_crops(pn2, pts, labels, extra, npnt, 6, 1)
    c2, s2 = _check_crops(pn2, pts, labels, extra, npnt, 6, 2)
    assert not np.array_equal(c1, c2)                                        # different seeds, different crops
    # the reference's own items obey the same rules the device sampler is checked against (fixture sanity)
    f = v[tag + "_features"]
    assert f.shape[1] == npnt and np.all(np.abs(f[:, :, 0]) <= 0.5 + 1e-9) and np.all(np.abs(f[:, :, 1]) <= 0.5 + 1e-9)


def test_training_crops_large_room_many_crops(pn2):
    g = np.random.RandomState(8)
    P = 300_000
    pts = np.stack([g.uniform(0, 8, P), g.normal(0, 0.3, P), g.uniform(0, 6, P)], axis=1) + np.array([500.0, 20.0, 1.0])
    labels = g.randint(0, 18, P).astype(np.int32)
    extra = [g.randint(0, 256, P).astype(np.float64) for _ in range(3)] + [g.normal(size=P)]
    _check_crops(pn2, pts, labels, extra, 4096, 70, 5)                       # > 64 crops: two launches of the membership kernel
    sparse = pts[:3000]                                                      # no block can hold 1025 points
    with pytest.raises(RuntimeError):
        pn2.sample_training_crops(torch.from_numpy(sparse).to(DEV), torch.zeros(3000, dtype=torch.int64, device=DEV), 2,
                                  max_rounds=3)
