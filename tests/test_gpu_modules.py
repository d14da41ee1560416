"""Module- and model-level parity (forward, backward, running statistics) against fixtures of
the unmodified reference, in fp32 mode (tight tolerance) and bf16 mode (stated loose tolerance).

Tolerances (stated per SURVEY.md 7.3-6, set from measured error):
  fp32 rows: module activations atol 2e-4 (different fp32 summation order than MKL), eval
             log-probs atol 1e-3, train-mode log-probs atol 1e-2: measured 2.3e-3 on the B200,
             where the reference's own fp32 CPU result is 1.9e-3 away from an fp64 evaluation of
             the same network and its own CUDA path (TF32 convolutions) is 1.1 away;
             module-level gradients rtol 2e-3 of the tensor's max magnitude; whole-network
             gradients by relative L2 error <= 3e-2 and cosine >= 0.999: ReLU masks and
             max-pool arg-maxes flip where a pre-activation is within rounding noise of the
             threshold, so the gradient error goes like the square root of the forward noise --
             the reference's own fp32 CPU gradients are 1.0e-2 away from an fp64 evaluation
             of the same network on this input (ours: 1.6e-2).
  bf16 rows: activations within 2 % of the reference's max magnitude (bf16 has an 8-bit
             mantissa; three chained layers); eval log-probs atol 0.15 with >= 97 % arg-max
             agreement; train-mode log-probs by relative RMS error <= 0.2 (measured 0.09 with
             these synthetic weights, 0.05 with default initialisation; train-mode batch norm
             re-normalises every layer, so rounding noise compounds through the 27 layers -- the
             reference's own TF32 CUDA path measures 0.03 on the same input); gradients by
             cosine >= 0.98 at module level.
"""
import numpy as np
import pytest
import torch

import _inputs as I
from oracle import pn2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32_default(pn2):
    pn2.set_precision("fp32")
    yield
    pn2.set_precision("fp32")


def _close(got, want, atol, what=""):
    got = got.detach().float().cpu().numpy() if hasattr(got, "detach") else got
    err = np.abs(got - want).max()
    assert err <= atol, "%s: max abs err %.3e > %.1e" % (what, err, atol)


def _grad_close(got, want, rel, what=""):
    got = got.detach().float().cpu().numpy()
    scale = max(np.abs(want).max(), 1e-6)
    err = np.abs(got - want).max() / scale
    assert err <= rel, "%s: max err / max|ref| = %.3e > %.1e" % (what, err, rel)


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def _sa_case(pn2, g, precision):
    pn2.set_precision(precision)
    x = I.facade_batch(2, 256, 9, 5).to(DEV).transpose(2, 1)      # strided [B,9,N] like localfunctions.py:209
    xyz = x[:, :3, :]
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(64, 0.3, 16, 12, [16, 16, 32], False), 41).to(DEV).train()
    pts = x.clone().requires_grad_(True)
    torch.manual_seed(51)
    nx, out = sa(xyz, pts)
    assert nx.shape == (2, 3, 64) and out.shape == (2, 32, 64) and out.dtype == torch.float32
    wsel = torch.rand(out.shape, generator=torch.Generator().manual_seed(6)).to(DEV)
    (out * wsel).sum().backward()
    return sa, x, xyz, pts, nx, out


def test_set_abstraction_fp32(pn2, golden):
    g = golden("modules")
    sa, x, xyz, pts, nx, out = _sa_case(pn2, g, "fp32")
    assert np.array_equal(nx.cpu().numpy(), g["sa_train_new_xyz"])           # sampled centroids: exact
    _close(out, g["sa_train_out"], 2e-4, "sa train out")
    _grad_close(pts.grad, g["sa_train_dpoints"], 2e-3, "sa dpoints")
    for n, p in sa.named_parameters():
        want = g["grad/sa." + n]
        if n.startswith("mlp_convs") and n.endswith("bias"):
            _close(p.grad, np.zeros_like(want), 1e-4, n)                      # cancelled by batch norm
            assert np.abs(want).max() < 1e-2      # the reference's own value is rounding noise
        else:
            _grad_close(p.grad, want, 2e-3, n)
    for n, b in sa.named_buffers():
        if b.is_floating_point():
            _close(b, g["sa_buf_after/" + n], 1e-5, n)
        else:
            assert int(b) == int(g["sa_buf_after/" + n])
    sa.eval()
    torch.manual_seed(52)
    with torch.no_grad():
        _, out = sa(xyz, x)
    _close(out, g["sa_eval_out"], 2e-4, "sa eval out")


def test_set_abstraction_bf16(pn2, golden):
    g = golden("modules")
    sa, x, xyz, pts, nx, out = _sa_case(pn2, g, "bf16")
    assert np.array_equal(nx.cpu().numpy(), g["sa_train_new_xyz"])
    _close(out, g["sa_train_out"], 0.02 * float(np.abs(g["sa_train_out"]).max()), "sa train out (bf16)")
    assert _cos(pts.grad.cpu().numpy(), g["sa_train_dpoints"]) > 0.98
    for n, p in sa.named_parameters():
        if not (n.startswith("mlp_convs") and n.endswith("bias")):
            assert _cos(p.grad.cpu().numpy(), g["grad/sa." + n]) > 0.98, n


@pytest.mark.parametrize("precision,atol,rel", [("fp32", 2e-4, 2e-3), ("bf16", 0.15, None)])
def test_feature_propagation(pn2, golden, precision, atol, rel):
    g = golden("modules")
    pn2.set_precision(precision)
    x = I.facade_batch(2, 256, 9, 5).to(DEV).transpose(2, 1)
    xyz = x[:, :3, :]
    coarse = torch.from_numpy(g["fp_coarse_xyz"]).to(DEV)
    p1 = torch.rand(2, 7, 256, generator=torch.Generator().manual_seed(8)).to(DEV).requires_grad_(True)
    p2 = torch.rand(2, 32, 64, generator=torch.Generator().manual_seed(9)).to(DEV).requires_grad_(True)
    fp = I.randomize_module_(pn2.PointNetFeaturePropagation(39, [24, 16]), 42).to(DEV).train()
    y = fp(xyz, coarse, p1, p2)
    assert y.shape == (2, 16, 256)
    wsel = torch.rand(y.shape, generator=torch.Generator().manual_seed(10)).to(DEV)
    (y * wsel).sum().backward()
    _close(y, g["fp_train_out"], atol, "fp train out")
    checks = [(p1.grad, g["fp_train_dp1"], "dp1"), (p2.grad, g["fp_train_dp2"], "dp2")]
    checks += [(p.grad, g["grad/fp." + n], n) for n, p in fp.named_parameters()
               if not (n.startswith("mlp_convs") and n.endswith("bias"))]
    for got, want, n in checks:
        if rel is not None:
            _grad_close(got, want, rel, n)
        else:
            assert _cos(got.cpu().numpy(), want) > 0.98, n
    if precision == "fp32":
        for n, b in fp.named_buffers():
            if b.is_floating_point():
                _close(b, g["fp_buf_after/" + n], 1e-5, n)
    fp.eval()
    with torch.no_grad():
        _close(fp(xyz, coarse, p1, p2), g["fp_eval_out"], atol, "fp eval out")
        fp2 = I.randomize_module_(pn2.PointNetFeaturePropagation(32, [24, 16]), 43).to(DEV).eval()
        _close(fp2(xyz, coarse, None, p2), g["fp_eval_out_nop1"], atol, "fp eval out (points1=None)")


def test_feature_propagation_single_coarse_point(pn2):
    """S == 1 branch (pointnet2_utils.py:293-294): every fine point copies the only coarse feature."""
    fp = I.randomize_module_(pn2.PointNetFeaturePropagation(8, [8]), 3).to(DEV).eval()
    ref = I.randomize_module_(O.OracleFP(8, [8]), 3).eval()
    xyz1, xyz2 = torch.rand(2, 3, 40), torch.rand(2, 3, 1)
    p2 = torch.rand(2, 8, 1)
    with torch.no_grad():
        want = ref(xyz1, xyz2, None, p2)
        got = fp(xyz1.to(DEV), xyz2.to(DEV), None, p2.to(DEV))
    _close(got, want.numpy(), 1e-5, "S==1")


def test_group_all_and_msg_variants_match_oracle(pn2):
    x = I.facade_batch(2, 128, 9, 6).transpose(2, 1)
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(None, None, None, 12, [16, 24], True), 5).to(DEV).train()
    ref = I.randomize_module_(O.OracleSA(None, None, None, 12, [16, 24], True), 5).train()
    a = x.clone().requires_grad_(True)
    b = x.clone().to(DEV).requires_grad_(True)
    nx_w, out_w = ref(x[:, :3, :], a)
    nx_g, out_g = sa(b[:, :3, :].detach(), b)
    assert nx_g.shape == (2, 3, 1) and float(nx_g.abs().max()) == 0.0
    _close(out_g, out_w.detach().numpy(), 2e-4, "group_all out")
    out_w.sum().backward()
    out_g.sum().backward()
    _grad_close(b.grad, a.grad.numpy(), 2e-3, "group_all dpoints")

    msg = I.randomize_module_(pn2.PointNetSetAbstractionMsg(32, [0.2, 0.4], [8, 16], 9, [[16, 16], [16, 32]]), 7).to(DEV).train()
    rmsg = I.randomize_module_(O.OracleSAMsg(32, [0.2, 0.4], [8, 16], 9, [[16, 16], [16, 32]]), 7).train()
    a = x.clone().requires_grad_(True)
    b = x.clone().to(DEV).requires_grad_(True)
    torch.manual_seed(3)
    nx_w, out_w = rmsg(x[:, :3, :], a)
    torch.manual_seed(3)
    nx_g, out_g = msg(b[:, :3, :].detach(), b)
    assert np.array_equal(nx_g.cpu().numpy(), nx_w.numpy())
    _close(out_g, out_w.detach().numpy(), 2e-4, "msg out")
    out_w.sum().backward()
    out_g.sum().backward()
    _grad_close(b.grad, a.grad.numpy(), 2e-3, "msg dpoints")
    for (n, p), (_, q) in zip(msg.named_parameters(), rmsg.named_parameters()):
        if not n.endswith("bias") or "bn_blocks" in n:
            _grad_close(p.grad, q.grad.numpy(), 2e-3, n)


def _model(pn2):
    net = I.randomize_module_(pn2.get_model(18, 3), 61)
    net.drop1.p = 0.0
    return net.to(DEV)


@pytest.mark.parametrize("precision,atol", [("fp32", 1e-3), ("bf16", 0.15)])
def test_model_eval_logits(pn2, golden, precision, atol):
    g = golden("model")
    pn2.set_precision(precision)
    net = _model(pn2).eval()
    for tag, x in (("facade", I.facade_batch(2, 2048, 9, int(g["facade_seed"])).to(DEV).transpose(2, 1)), ("cube", I.cube_batch(2, 2048, 9, int(g["cube_seed"])).to(DEV))):
        torch.manual_seed(71)
        with torch.no_grad():
            pred, l4 = net(x)
        assert pred.shape == (2, 2048, 18) and l4.shape == (2, 512, 16)
        _close(pred, g[tag + "_eval_pred"], atol, tag + " log-probs")
        agree = (pred.argmax(-1).cpu().numpy() == g[tag + "_eval_pred"].argmax(-1)).mean()
        assert agree >= (0.999 if precision == "fp32" else 0.97), agree
        _close(l4, g[tag + "_eval_l4"], atol, tag + " l4_points")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model_train_step_gradients(pn2, golden, precision):
    g = golden("model")
    pn2.set_precision(precision)
    net = _model(pn2).train()
    x = I.facade_batch(2, 2048, 9, int(g["facade_seed"])).to(DEV).transpose(2, 1)
    target = I.labels(2, 2048, 18, 7).to(DEV)
    weights = torch.linspace(0.5, 1.5, 18).to(DEV)
    torch.manual_seed(72)
    pred, _ = net(x)
    loss = pn2.get_loss()(pred.contiguous().view(-1, 18), target, None, weights)
    loss.backward()
    if precision == "fp32":
        _close(pred, g["train_pred"], 1e-2, "train log-probs")
        assert abs(loss.item() - float(g["train_loss"])) < 1e-4
    else:
        want = torch.from_numpy(g["train_pred"])
        rel_rms = float((pred.detach().cpu() - want).pow(2).sum().sqrt() / want.pow(2).sum().sqrt())
        assert rel_rms <= 0.2, rel_rms
        assert abs(loss.item() - float(g["train_loss"])) < 5e-2
    worst = 1.0
    for n, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if ("mlp_convs" in n and n.endswith("bias")) or n == "conv1.bias":
            continue       # conv bias in front of a train-mode BatchNorm: its gradient is pure rounding noise
        s = g["grad_stat/" + n]
        l2 = float(p.grad.double().pow(2).sum().sqrt())
        if precision == "bf16":
            continue       # whole-network bf16 gradients: covered by finiteness here, by cosine at module level
        assert abs(l2 - s[2]) <= 3e-2 * max(s[2], 1e-6) + 1e-7, (n, l2, s[2])
        if "grad/" + n in g.files:
            want = g["grad/" + n]
            c = _cos(p.grad.cpu().numpy(), want)
            worst = min(worst, c)
            assert c > 0.999, (n, c)
            rel = np.linalg.norm(p.grad.cpu().numpy().ravel() - want.ravel()) / max(np.linalg.norm(want.ravel()), 1e-12)
            assert rel <= 3e-2, (n, rel)
    if precision == "fp32":
        for n, b in net.named_buffers():
            if "buf_after/" + n in g.files and b.is_floating_point():
                _close(b, g["buf_after/" + n], 1e-4, n)


def test_config2_full_size_train_step_properties(pn2):
    """BASELINE.json config 2 (32 x 4096 x 9 ch, bf16): size-independent invariants."""
    pn2.set_precision("bf16")
    net = pn2.get_model(18, 3).to(DEV).train()
    x = I.facade_batch(32, 4096, 9, 11).to(DEV).transpose(2, 1)
    xyz = x[:, :3, :].permute(0, 2, 1)
    torch.manual_seed(1)
    fps, new_xyz = pn2.farthest_point_sample(xyz, 1024, return_xyz=True)
    s = fps.sort(dim=1)[0]
    dup = int((s[:, 1:] == s[:, :-1]).sum())
    assert int(fps.min()) >= 0 and int(fps.max()) < 4096
    assert dup <= 32 * 8          # FPS only repeats an index once every point is already covered (duplicated points)
    ball, cnt = pn2.query_ball_point(0.1, 32, xyz, new_xyz, return_count=True)
    k = torch.arange(32, device=DEV).view(1, 1, 32)
    valid = k < cnt.unsqueeze(-1)
    inc = (ball[:, :, 1:] > ball[:, :, :-1]) | ~valid[:, :, 1:]
    assert bool(inc.all())                                              # ascending index order
    assert bool((ball[~valid] == ball[:, :, :1].expand_as(ball)[~valid]).all())   # padded with the first hit
    d = (pn2.index_points(xyz, ball) - new_xyz.unsqueeze(2)).pow(2).sum(-1)
    assert float(d.max()) <= 0.1 ** 2 * 1.001
    target = I.labels(32, 4096, 18, 3).to(DEV)
    pred, l4 = net(x)
    loss = pn2.get_loss()(pred.contiguous().view(-1, 18), target, None, torch.ones(18, device=DEV))
    loss.backward()
    assert pred.shape == (32, 4096, 18) and torch.isfinite(pred).all()
    assert abs(float(pred.exp().sum(-1).mean()) - 1.0) < 1e-3
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_predict_blocks_graph_matches_eager_and_shards(pn2):
    """Whole-scene style inference: the CUDA-graph predictor labels blocks exactly like the eager forward
    (same kernels, same CPU-generator draws for the FPS start indices), a short tail batch rides in the
    fixed-shape graph, and rank shards concatenate to the full result."""
    pn2.set_precision("bf16")
    net = I.randomize_module_(pn2.get_model(18, 3), 61).to(DEV).eval()
    blocks = I.facade_batch(7, 1024, 9, 21)                       # 7 blocks, batch 3 -> tail of 1
    torch.manual_seed(5)
    lo, hi, eager = pn2.predict_blocks(net, blocks, batch_size=3, use_graph=False)
    torch.manual_seed(5)
    _, _, graphed = pn2.predict_blocks(net, blocks, batch_size=3, use_graph=True)
    assert (lo, hi) == (0, 7) and eager.shape == (7, 1024)
    assert torch.equal(eager[:6], graphed[:6])
    # the tail batch draws 3 start indices per level in the fixed-shape graph and 1 in the eager call: another
    # (equally valid) sampling, so only the label range is checked
    assert int(graphed[6].min()) >= 0 and int(graphed[6].max()) < 18
    parts = []
    for r in range(2):
        torch.manual_seed(5)
        lo, hi, lab = pn2.predict_blocks(net, blocks, batch_size=3, rank=r, world=2, use_graph=False)
        parts.append((lo, hi, lab))
    assert [(p[0], p[1]) for p in parts] == [(0, 4), (4, 7)]
    # blocks are independent in eval mode, but the FPS start draws follow batch order: compare label agreement
    both = torch.cat([p[2] for p in parts])
    assert both.shape == eager.shape
    assert float((both[:3] == eager[:3]).float().mean()) == 1.0   # first batch of rank 0 saw the same draws
