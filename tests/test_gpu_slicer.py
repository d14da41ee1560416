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
