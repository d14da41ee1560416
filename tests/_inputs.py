"""Deterministic synthetic inputs shared by the golden-fixture generator and the tests.

S1 "unit cube" mirrors the reference's own smoke input torch.rand(6,9,2048)
(/root/reference/models/pointnet2_sem_seg.py:55); S2 "facade block" mirrors
TrainCustomDataset.__getitem__ (/root/reference/sem_seg_training.py:223-231): x in
(-0.5,0.5), y a thin wall, z un-centred in (0,3), 12.5 % duplicated points (the
reference pads blocks by re-drawing points, sem_seg_training.py:221).  Everything
comes from seeded CPU generators, which are machine independent.
"""
import numpy as np
import torch


def cube_xyz(B, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, 3, generator=g)


def facade_xyz(B, N, seed=1, dup_frac=0.125):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, generator=g) - 0.5
    y = (torch.randn(B, N, generator=g) * 0.05).clamp(-0.5, 0.5)
    z = torch.rand(B, N, generator=g) * 3.0
    p = torch.stack([x, y, z], -1)
    nd = int(N * dup_frac)
    if nd:
        src = torch.randint(0, N, (B, nd), generator=g)
        dst = torch.randint(0, N, (B, nd), generator=g)
        for b in range(B):
            p[b, dst[b]] = p[b, src[b]]
    return p.contiguous()


def facade_batch(B, N, C=9, seed=1):
    """[B, N, C] point-major batch as the DataLoader yields it (float32):
    ch 0-2 block-centred xyz, 3-5 xyz/room_max in (0,1], 6-8 colour/255."""
    xyz = facade_xyz(B, N, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    room = torch.tensor([60.0, 1.0, 25.0])
    norm = ((xyz + torch.tensor([30.0, 0.5, 0.0])) / room).clamp(0, 1)
    cols = [xyz, norm]
    if C > 6:
        cols.append(torch.randint(0, 256, (B, N, C - 6), generator=g).float() / 255.0)
    return torch.cat(cols, -1).contiguous()


def cube_batch(B, N, C=9, seed=0):
    """[B, C, N] channel-major batch == torch.rand(B, C, N) with a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, C, N, generator=g)


def start_indices(B, N, seed):
    """The draw farthest_point_sample makes (pointnet2_utils.py:75) under torch.manual_seed(seed)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, N, (B,), dtype=torch.long, generator=g)


def labels(B, N, num_classes=18, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, num_classes, (B * N,), generator=g)


def checksum(t):
    a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return float(np.float64(np.abs(a.astype(np.float64)).sum()))


def randomize_module_(module, seed):
    """Fill every parameter / BN buffer from a seeded generator, keyed by sorted
    state_dict name, so any two modules with the same state_dict layout get
    identical, non-trivial values regardless of construction order."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd):
            t = sd[name]
            if not t.is_floating_point():
                continue
            r = torch.rand(t.shape, generator=g)
            if name.endswith("running_var"):
                v = 0.5 + r
            elif name.endswith("running_mean"):
                v = (r - 0.5) * 0.2
            elif name.endswith("weight") and t.dim() == 1:      # BN gamma
                v = 0.5 + r
            elif name.endswith("bias"):
                v = (r - 0.5) * 0.2
            else:                                               # conv weight
                fan_in = t.shape[1]
                v = (r - 0.5) * (2.0 / fan_in ** 0.5) * 1.7
            t.copy_(v.to(t.dtype))
    return module
