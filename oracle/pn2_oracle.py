"""Torch-CPU restatement ("port") of the reference's PointNet++ SSG hot path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.

What it restates (all line numbers are /root/reference/models/pointnet2_utils.py
unless another file is named):

  pairwise_sqdist      square_distance            :19-40
  take_points          index_points               :43-60
  fps                  farthest_point_sample      :63-84
  ball_query           query_ball_point           :87-107
  group                sample_and_group           :110-138
  group_all            sample_and_group_all       :141-158
  OracleSA             PointNetSetAbstraction     :161-202
  OracleSAMsg          PointNetSetAbstractionMsg  :205-262
  OracleFP             PointNetFeaturePropagation :265-315
  OracleSemSeg         get_model   (models/pointnet2_sem_seg.py:6-40)
  nll                  get_loss    (models/pointnet2_sem_seg.py:44-50)
  add_vote             add_vote    (localfunctions.py:336-343)
  vote_argmax          np.argmax(vote_label_pool, 1)   (localfunctions.py:405)
  rotate_z             rotate_point_cloud_z        (provider.py:66-84)
  slice_scene          TestCustomDataset.__getitem__  (sem_seg_testing.py:182-254)
  train_crop           TrainCustomDataset.__getitem__ (sem_seg_training.py:200-259)

The arithmetic itself lives in PyTorch (third party, un-pinned by the reference;
effective pin torch 2.11.0+cu128 of this image), so this port issues the same
ATen operations in the same order (matmul/sum/sort/max/conv/batch_norm); the
machine-independent statement of the rounding order is oracle/pn2_oracle.c.

Pinning: tests/test_oracle_golden.py compares this module and the C file with
fixtures generated from the unmodified reference by tests/golden/make_golden.py
(run in the authoring container where /root/reference is mounted).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


# When set (tests only), index-producing geometry runs in this dtype even if the features are
# evaluated in another one, e.g. an fp64 evaluation of the network on the reference's fp32
# sampling / grouping / neighbour indices (used to tell rounding noise from real differences).
GEOMETRY_DTYPE = None


def _geo(t):
    return t if GEOMETRY_DTYPE is None else t.to(GEOMETRY_DTYPE)


def pairwise_sqdist(src, dst):
    # :37-39  -2*src@dst^T, then += |src|^2, then += |dst|^2 (this order)
    out = torch.matmul(src, dst.transpose(1, 2)) * -2
    out += (src ** 2).sum(-1).unsqueeze(-1)
    out += (dst ** 2).sum(-1).unsqueeze(-2)
    return out


def take_points(points, idx):
    # :53-59  points[b, idx[b, ...], :]
    B = points.shape[0]
    rows = torch.arange(B, dtype=torch.long, device=points.device)
    rows = rows.reshape([B] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[rows, idx]


def fps(xyz, npoint, start=None):
    # :73-84; the start index is one CPU-generator randint draw (:75)
    xyz = _geo(xyz)
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    far = start.to(xyz.device)
    picked = torch.empty(B, npoint, dtype=torch.long, device=xyz.device)
    nearest = torch.full((B, N), 1e10, dtype=xyz.dtype, device=xyz.device)
    rows = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        picked[:, i] = far
        centre = xyz[rows, far].unsqueeze(1)
        d = ((xyz - centre) ** 2).sum(-1)
        nearest = torch.where(d < nearest, d, nearest)
        far = nearest.max(-1)[1]
    return picked


def ball_query(radius, nsample, xyz, new_xyz):
    # :96-106  first nsample in-radius indices in ascending order, padded with the first
    xyz, new_xyz = _geo(xyz), _geo(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    d = pairwise_sqdist(new_xyz, xyz)
    cand = torch.arange(N, dtype=torch.long, device=xyz.device).expand(B, S, N).clone()
    cand[d > radius ** 2] = N
    kept = cand.sort(dim=-1)[0][:, :, :nsample]
    lead = kept[:, :, :1].expand(-1, -1, kept.shape[-1])
    return torch.where(kept == N, lead, kept)


def group(npoint, radius, nsample, xyz, points, start=None, return_idx=False):
    # :121-138
    picked = fps(xyz, npoint, start)
    new_xyz = take_points(xyz, picked)
    nbr = ball_query(radius, nsample, xyz, new_xyz)
    local = take_points(xyz, nbr) - new_xyz.unsqueeze(2)
    feats = local if points is None else torch.cat([local, take_points(points, nbr)], -1)
    if return_idx:
        return new_xyz, feats, picked, nbr
    return new_xyz, feats


def group_all(xyz, points):
    # :150-158
    B, N, C = xyz.shape
    new_xyz = xyz.new_zeros(B, 1, C)
    feats = xyz.unsqueeze(1)
    if points is not None:
        feats = torch.cat([feats, points.unsqueeze(1)], -1)
    return new_xyz, feats


def _mlp_stack(kind, widths, cin):
    conv = nn.Conv2d if kind == 2 else nn.Conv1d
    bn = nn.BatchNorm2d if kind == 2 else nn.BatchNorm1d
    convs, bns = nn.ModuleList(), nn.ModuleList()
    for w in widths:
        convs.append(conv(cin, w, 1))
        bns.append(bn(w))
        cin = w
    return convs, bns


class OracleSA(nn.Module):
    """:161-202"""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs, self.mlp_bns = _mlp_stack(2, mlp, in_channel)

    def forward(self, xyz, points):
        xyz_t = xyz.transpose(1, 2)
        pts_t = None if points is None else points.transpose(1, 2)
        if self.group_all:
            new_xyz, feats = group_all(xyz_t, pts_t)
        else:
            new_xyz, feats = group(self.npoint, self.radius, self.nsample, xyz_t, pts_t)
        h = feats.permute(0, 3, 2, 1)          # [B, C+D, nsample, npoint]  (:195)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            h = F.relu(bn(conv(h)))
        return new_xyz.transpose(1, 2), h.max(2)[0]


class OracleSAMsg(nn.Module):
    """:205-262 (feature-first concat order, :248)"""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for widths in mlp_list:
            c, b = _mlp_stack(2, widths, in_channel + 3)
            self.conv_blocks.append(c)
            self.bn_blocks.append(b)

    def forward(self, xyz, points):
        xyz_t = xyz.transpose(1, 2)
        pts_t = None if points is None else points.transpose(1, 2)
        new_xyz = take_points(xyz_t, fps(xyz_t, self.npoint))
        outs = []
        for r, k, convs, bns in zip(self.radius_list, self.nsample_list, self.conv_blocks, self.bn_blocks):
            nbr = ball_query(r, k, xyz_t, new_xyz)
            local = take_points(xyz_t, nbr) - new_xyz.unsqueeze(2)
            h = local if pts_t is None else torch.cat([take_points(pts_t, nbr), local], -1)
            h = h.permute(0, 3, 2, 1)
            for conv, bn in zip(convs, bns):
                h = F.relu(bn(conv(h)))
            outs.append(h.max(2)[0])
        return new_xyz.transpose(1, 2), torch.cat(outs, 1)


def three_nn_weights(xyz1, xyz2):
    # :296-302
    out_dtype = xyz1.dtype
    d, order = pairwise_sqdist(_geo(xyz1), _geo(xyz2)).sort(dim=-1)
    d, order = d[:, :, :3], order[:, :, :3]
    recip = 1.0 / (d + 1e-8)
    return order, (recip / recip.sum(dim=2, keepdim=True)).to(out_dtype)


class OracleFP(nn.Module):
    """:265-315"""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = _mlp_stack(1, mlp, in_channel)

    def forward(self, xyz1, xyz2, points1, points2):
        f, c = xyz1.transpose(1, 2), xyz2.transpose(1, 2)
        coarse = points2.transpose(1, 2)
        B, N, _ = f.shape
        if c.shape[1] == 1:
            up = coarse.repeat(1, N, 1)
        else:
            order, w = three_nn_weights(f, c)
            up = (take_points(coarse, order) * w.unsqueeze(-1)).sum(dim=2)
        if points1 is not None:
            up = torch.cat([points1.transpose(1, 2), up], -1)
        h = up.transpose(1, 2)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            h = F.relu(bn(conv(h)))
        return h


# (npoint, radius, nsample, mlp) per level and FP (in, mlp): models/pointnet2_sem_seg.py:9-16
SA_SPEC = ((1024, 0.1, 32, (32, 32, 64)), (256, 0.2, 32, (64, 64, 128)),
           (64, 0.4, 32, (128, 128, 256)), (16, 0.8, 32, (256, 256, 512)))
FP_SPEC = ((768, (256, 256)), (384, (256, 256)), (320, (256, 128)), (128, (128, 128, 128)))


class OracleSemSeg(nn.Module):
    """models/pointnet2_sem_seg.py:6-40 with the oracle modules."""

    def __init__(self, num_classes, num_extra_features, sa_cls=OracleSA, fp_cls=OracleFP):
        super().__init__()
        cin = 6 + 3 + num_extra_features
        for i, (npoint, radius, nsample, mlp) in enumerate(SA_SPEC, 1):
            setattr(self, "sa%d" % i, sa_cls(npoint, radius, nsample, cin, list(mlp), False))
            cin = mlp[-1] + 3
        for i, (c, mlp) in zip((4, 3, 2, 1), FP_SPEC):
            setattr(self, "fp%d" % i, fp_cls(c, list(mlp)))
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)

    def forward(self, x):
        xyz0, f0 = x[:, :3, :], x
        xyz1, f1 = self.sa1(xyz0, f0)
        xyz2, f2 = self.sa2(xyz1, f1)
        xyz3, f3 = self.sa3(xyz2, f2)
        xyz4, f4 = self.sa4(xyz3, f3)
        f3 = self.fp4(xyz3, xyz4, f3, f4)
        f2 = self.fp3(xyz2, xyz3, f2, f3)
        f1 = self.fp2(xyz1, xyz2, f1, f2)
        f0 = self.fp1(xyz0, xyz1, None, f1)
        h = self.drop1(F.relu(self.bn1(self.conv1(f0))))
        h = F.log_softmax(self.conv2(h), dim=1)
        return h.permute(0, 2, 1), f4


def nll(pred, target, weight=None):
    # models/pointnet2_sem_seg.py:47-48
    return F.nll_loss(pred, target, weight=weight)


def add_vote(vote_label_pool, point_idx, pred_label, weight):
    """localfunctions.py:336-343 -- for every (b, n) with weight != 0 and not inf: pool[int(idx), int(label)] += 1.
    numpy restatement of the Python double loop (np.add.at handles repeated indices like the loop does); the pool is
    the reference's float64 [P, NC] array, updated in place and returned."""
    import numpy as np
    w = np.asarray(weight)
    keep = (w != 0) & ~np.isinf(w)
    idx = np.asarray(point_idx)[keep].astype(np.int64)
    lab = np.asarray(pred_label)[keep].astype(np.int64)
    np.add.at(vote_label_pool, (idx, lab), 1)
    return vote_label_pool


def vote_argmax(vote_label_pool):
    """localfunctions.py:405"""
    import numpy as np
    return np.argmax(vote_label_pool, 1)


def rotate_z(batch_xyz, angles):
    """provider.py:66-84 with the angles given instead of drawn: per cloud, float32 xyz [N,3] times the float64 matrix
    [[c, s, 0], [-s, c, 0], [0, 0, 1]] (np.dot -> float64), stored into a float32 array."""
    import numpy as np
    out = np.zeros(batch_xyz.shape, dtype=np.float32)
    for k in range(batch_xyz.shape[0]):
        c, s = np.cos(angles[k]), np.sin(angles[k])
        m = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
        out[k, ...] = np.dot(batch_xyz[k, ...].reshape((-1, 3)), m)
    return out


def slice_scene(points, labels, extra, extra_names, labelweights, block_size=1.0, stride=0.5, padding=0.001, block_points=4096):
    """sem_seg_testing.py:182-254 for one scene: numpy restatement of TestCustomDataset.__getitem__ that makes the SAME
    calls to numpy's global generator in the same order (np.random.choice, np.random.shuffle per non-empty cell), so under
    the same seed it reproduces the reference's four arrays exactly.  Also returns the per-cell structure
    [(cell index iy * grid_x + ix, member point indices, number of blocks)] that the random parts do not change.
    points [P,3] float64, labels [P] int, extra: list of E arrays [P], labelweights [NC]."""
    import numpy as np
    pts = points[:, :3]
    coord_min, coord_max = np.amin(pts, axis=0)[:3], np.amax(pts, axis=0)[:3]
    grid_x = int(np.ceil(float(coord_max[0] - coord_min[0] - block_size) / stride) + 1)
    grid_y = int(np.ceil(float(coord_max[1] - coord_min[1] - block_size) / stride) + 1)
    data_room, label_room, sample_weight, index_room = np.array([]), np.array([]), np.array([]), np.array([])
    cells = []
    for index_y in range(grid_y):
        for index_x in range(grid_x):
            s_x = coord_min[0] + index_x * stride
            e_x = min(s_x + block_size, coord_max[0])
            s_x = e_x - block_size
            s_y = coord_min[1] + index_y * stride
            e_y = min(s_y + block_size, coord_max[1])
            s_y = e_y - block_size
            idxs = np.where((pts[:, 0] >= s_x - padding) & (pts[:, 0] <= e_x + padding) &
                            (pts[:, 1] >= s_y - padding) & (pts[:, 1] <= e_y + padding))[0]
            if idxs.size == 0:
                continue
            num_batch = int(np.ceil(idxs.size / block_points))
            point_size = int(num_batch * block_points)
            cells.append((index_y * grid_x + index_x, idxs.copy(), num_batch))
            replace = False if (point_size - idxs.size <= idxs.size) else True
            rep = np.random.choice(idxs, point_size - idxs.size, replace=replace)
            idxs = np.concatenate((idxs, rep))
            np.random.shuffle(idxs)
            batch = pts[idxs, :]
            norm = np.zeros((point_size, 3))
            norm[:, 0] = batch[:, 0] / coord_max[0]
            norm[:, 1] = batch[:, 1] / coord_max[1]
            norm[:, 2] = batch[:, 2] / coord_max[2]
            batch[:, 0] = batch[:, 0] - (s_x + block_size / 2.0)
            batch[:, 1] = batch[:, 1] - (s_y + block_size / 2.0)
            batch = np.concatenate((batch, norm), axis=1)
            lab = labels[idxs].astype(int)
            w = labelweights[lab]
            if len(extra_names) > 0:
                feats = np.zeros((point_size, len(extra_names)))
                for ix, name in enumerate(extra_names):
                    sel = extra[ix][idxs]
                    if name in ("red", "blue", "green"):
                        sel = sel / 255
                    feats[:, ix] = np.array(sel)
                batch = np.concatenate((batch, feats), axis=1)
            data_room = np.vstack([data_room, batch]) if data_room.size else batch
            label_room = np.hstack([label_room, lab]) if label_room.size else lab
            sample_weight = np.hstack([sample_weight, w]) if label_room.size else w
            index_room = np.hstack([index_room, idxs]) if index_room.size else idxs
    if not cells:
        return None, None, None, None, cells
    data_room = data_room.reshape((-1, block_points, data_room.shape[1]))
    return (data_room, label_room.reshape((-1, block_points)), sample_weight.reshape((-1, block_points)),
            index_room.reshape((-1, block_points)), cells)


def train_crop(points, labels, extra, extra_names, coord_max, num_point=4096, block_size=1.0):
    """sem_seg_training.py:200-259 for one item of one room: numpy restatement of TrainCustomDataset.__getitem__ making the
    same calls to numpy's global generator in the same order (choice of the centre until the block holds > 1024 points, then
    choice of num_point members), so under the same seed it reproduces the reference exactly.  Returns (features
    [num_point, 6 + E] float64, labels [num_point], centre [3], selected point indices)."""
    import numpy as np
    N_points = points.shape[0]
    while True:
        center = points[np.random.choice(N_points)][:3]
        block_min = center - [block_size / 2.0, block_size / 2.0, 0]
        block_max = center + [block_size / 2.0, block_size / 2.0, 0]
        idxs = np.where((points[:, 0] >= block_min[0]) & (points[:, 0] <= block_max[0]) & (points[:, 1] >= block_min[1]) &
                        (points[:, 1] <= block_max[1]))[0]
        if idxs.size > 1024:
            break
    sel = np.random.choice(idxs, num_point, replace=not (idxs.size >= num_point))
    selected = points[sel, :]
    cur = np.zeros((num_point, 6))
    cur[:, 3] = selected[:, 0] / coord_max[0]
    cur[:, 4] = selected[:, 1] / coord_max[1]
    cur[:, 5] = selected[:, 2] / coord_max[2]
    selected[:, 0] = selected[:, 0] - center[0]
    selected[:, 1] = selected[:, 1] - center[1]
    cur[:, 0:3] = selected
    feats = np.zeros((num_point, 6 + len(extra_names)))
    feats[:, :6] = cur
    for i, name in enumerate(extra_names):
        f = extra[i][sel]
        if name in ("red", "blue", "green"):
            f = f / 255
        feats[:, 6 + i] = f
    return feats, labels[sel], center, sel
