"""K3d -- the fused inference kernel of a set-abstraction level (csrc/sa_fused.cu) against
  (i) the torch-CPU port of the reference module (oracle; fp32) and
  (ii) this library's own per-layer path in fp32 rows (same gather, GEMMs on the FMA pipes),
on the four level shapes of the SSG network and on awkward shapes (ragged group counts, no
features, out-of-range indices, wide last layers, one- and four-layer MLPs).

Tolerance: the fused kernel rounds the gathered inputs, the weights and every hidden activation
to bf16 (8-bit mantissa) and accumulates in fp32, so with L chained layers the output error is
bounded by ~L * 2^-8 of the activation scale: |err| <= 2 % of the reference's max magnitude per level
(the same bound tests/test_gpu_modules.py uses for bf16 rows), and the mean error must stay
below 0.4 %."""
import importlib

import numpy as np
import pytest
import torch

import _inputs as I
from oracle import pn2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

LEVELS = [  # (N, S, radius, D, mlp)
    (1024, 256, 0.1, 9, [32, 32, 64]),
    (512, 128, 0.2, 64, [64, 64, 128]),
    (256, 64, 0.4, 128, [128, 128, 256]),
    (64, 16, 0.8, 256, [256, 256, 512]),
]


@pytest.fixture(autouse=True)
def _restore(pn2):
    yield
    pn2.set_precision("fp32")
    importlib.import_module(pn2.__name__ + ".modules").FUSED_EVAL = True


def _run(pn2, sa, xyz, pts, fused, precision):
    mods = importlib.import_module(pn2.__name__ + ".modules")
    lib_mod = importlib.import_module(pn2.__name__ + "._lib")
    mods.FUSED_EVAL = fused
    pn2.set_precision(precision)
    lib_mod.time_entry_point("pn2_sa_fused_eval")
    torch.manual_seed(7)
    with torch.no_grad():
        nx, out = sa(xyz, pts)
    n_fused_calls = len(lib_mod.timed_calls())
    lib_mod.time_entry_point(None)
    return nx, out, n_fused_calls


@pytest.mark.parametrize("N,S,radius,D,mlp", LEVELS)
@pytest.mark.parametrize("B", [1, 3])
def test_fused_level_matches_oracle_and_fp32_path(pn2, N, S, radius, D, mlp, B):
    g = torch.Generator().manual_seed(100 + D)
    xyz_h = I.facade_xyz(B, N, 3).transpose(2, 1).contiguous()                 # [B,3,N]
    pts_h = (torch.rand(B, D, N, generator=g) - 0.3)
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(S, radius, 32, D + 3, mlp, False), 40 + D).to(DEV).eval()
    ref = I.randomize_module_(O.OracleSA(S, radius, 32, D + 3, mlp, False), 40 + D).eval()
    xyz, pts = xyz_h.to(DEV), pts_h.to(DEV)
    nx_f, out_f, n_fused = _run(pn2, sa, xyz, pts, True, "bf16")
    nx_u, out_u, n_unfused = _run(pn2, sa, xyz, pts, False, "fp32")
    torch.manual_seed(7)
    with torch.no_grad():
        nx_r, out_r = ref(xyz_h, pts_h)
    assert torch.equal(nx_f, nx_u) and np.array_equal(nx_f.cpu().numpy(), nx_r.numpy())
    assert out_f.shape == out_u.shape == (B, mlp[-1], S) and out_f.dtype == torch.float32
    assert n_fused == 1 and n_unfused == 0                            # really the fused kernel / the per-layer path
    want = out_r.numpy()
    scale = float(np.abs(want).max())
    err_u = np.abs(out_u.cpu().numpy() - want)
    assert err_u.max() <= 2e-4 * max(scale, 1.0)                      # fp32 path == oracle
    err = np.abs(out_f.cpu().numpy() - want)
    assert err.max() <= 0.02 * scale, (err.max(), scale)
    assert err.mean() <= 0.004 * scale, (err.mean(), scale)


@pytest.mark.parametrize("B,N,S,D,mlp", [(2, 200, 7, 0, [16, 32]), (1, 300, 33, 5, [48]), (2, 128, 9, 12, [32, 16, 16, 40]),
                                          (1, 100, 5, 64, [512, 24])])
def test_fused_awkward_shapes(pn2, B, N, S, D, mlp):
    g = torch.Generator().manual_seed(9)
    xyz_h = I.cube_xyz(B, N, 5).transpose(2, 1).contiguous()
    pts_h = None if D == 0 else torch.rand(B, D, N, generator=g)
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(S, 0.25, 32, D + 3, mlp, False), 77).to(DEV).eval()
    xyz = xyz_h.to(DEV)
    pts = None if pts_h is None else pts_h.to(DEV)
    _, out_f, n_fused = _run(pn2, sa, xyz, pts, True, "bf16")
    _, out_u, n_unfused = _run(pn2, sa, xyz, pts, False, "fp32")
    assert n_fused == 1 and n_unfused == 0
    scale = float(out_u.abs().max())
    assert float((out_f - out_u).abs().max()) <= 0.02 * scale


def test_fused_out_of_range_indices_gather_zero_rows(pn2):
    """An empty ball leaves index N in every slot (the reference would raise): both paths gather zero rows."""
    mods = importlib.import_module(pn2.__name__ + ".modules")
    pn2.set_precision("bf16")
    B, N, S, D = 1, 64, 8, 16
    xyz = I.cube_xyz(B, N, 1).to(DEV)
    new_xyz = xyz[:, :S].contiguous()
    feats = torch.rand(B, N, D, device=DEV)
    idx = torch.randint(0, N, (B, S, 32), device=DEV)
    idx[0, 3] = N                                  # empty ball
    idx[0, 5, 7:] = N
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(S, 0.2, 32, D + 3, [32, 64], False), 3).to(DEV).eval()
    out = mods.sa_fused_eval(idx, sa.mlp_convs, sa.mlp_bns, new_xyz, xyz, feats)
    idx2 = idx.clone()
    safe = idx2.clamp(max=N - 1)
    rows = torch.cat([xyz[0][safe[0]] - new_xyz[0][:, None, :], feats[0][safe[0]]], -1)        # [S,32,3+D]
    rows[idx[0] >= N] = 0.0
    x = rows.reshape(S * 32, 3 + D)
    for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
        w = conv.weight.reshape(conv.out_channels, -1)
        z = x @ w.t() + conv.bias
        x = torch.relu((z - bn.running_mean) / torch.sqrt(bn.running_var + bn.eps) * bn.weight + bn.bias)
    want = x.reshape(S, 32, -1).max(1).values
    assert float((out[0] - want).abs().max()) <= 0.02 * float(want.abs().max())


def test_fused_rejects_what_it_cannot_do(pn2):
    mods = importlib.import_module(pn2.__name__ + ".modules")
    pn2.set_precision("bf16")
    sa = pn2.PointNetSetAbstraction(8, 0.2, 16, 6, [16, 16], False).to(DEV).eval()      # nsample 16
    assert not mods._fused_eval_applies(sa.mlp_convs, sa.mlp_bns, sa.nsample, [])
    sa = pn2.PointNetSetAbstraction(8, 0.2, 32, 6, [24, 16], False).to(DEV).eval()      # hidden width 24
    assert not mods._fused_eval_applies(sa.mlp_convs, sa.mlp_bns, sa.nsample, [], 3)
    sa = pn2.PointNetSetAbstraction(8, 0.2, 32, 6, [1024], False).to(DEV).eval()        # wider than tensor memory
    assert not mods._fused_eval_applies(sa.mlp_convs, sa.mlp_bns, sa.nsample, [], 3)
    sa = pn2.PointNetSetAbstraction(8, 0.2, 32, 6, [32, 16], False).to(DEV).train()     # batch statistics
    assert not mods._fused_eval_applies(sa.mlp_convs, sa.mlp_bns, sa.nsample, [])
    sa.eval()
    assert mods._fused_eval_applies(sa.mlp_convs, sa.mlp_bns, sa.nsample, [])
    x = I.cube_xyz(1, 64, 0).to(DEV).transpose(2, 1)
    _, out = sa(x, x)                              # grad mode with trainable parameters -> per-layer path, differentiable
    assert out.requires_grad
