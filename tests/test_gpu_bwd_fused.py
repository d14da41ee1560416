"""csrc/bwd_fused.cu (pn2_mlp_bwd_layer): one fused backward launch per MLP layer on bf16 rows.

Checked two ways:
* against the per-step kernels it replaces (reduce / dz / data gradient / weight gradient: modules.FUSED_BWD off) on the
  same module, input and upstream gradient -- the two paths round the SAME quantities to bf16 (dZ, the masked
  dA_{l-1}), so they agree to fp32 summation order -- except that dgamma is formed here as invstd . (sum dA'.z - mean . sum dA')
  from tensor-core products on the stored bf16 z (sum dA'.zhat with zhat in fp32 there): relative L2 error <= 6e-3 on
  every gradient (measured <= 2.3e-3);
* against the fp32 torch-CPU port of the reference modules (pointnet2_utils.py:161-202, :265-315): module-level
  gradient cosine >= 0.98, the repo's bf16 bar.
Shapes cover: nsample-pooled set abstraction with a 3 + D wide first layer whose width is not a multiple of 4 (fixed-order
weight-gradient partials), widths of 32 / 64 / 128, ragged M (not a multiple of the 128-row tile), a feature-propagation
chain with and without the input gradient, frozen (eval-mode) BatchNorm statistics, and the fused head.
"""
import importlib

import pytest
import torch

import _inputs as I
from oracle import pn2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _bf16(pn2):
    pn2.set_precision("bf16")
    yield
    pn2.set_precision("fp32")
    importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules").FUSED_BWD = True


def _modules():
    return importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules")


def _grads(mod, params, run):
    for p in params:
        p.grad = None
    extra = run()
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in params], extra


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-20))


def _ab(pn2, module, inputs, run):
    """gradients of `run()` with the fused kernel and with the per-step kernels"""
    M = _modules()
    lib = pn2.load()
    params = [p for p in module.parameters()]
    calls = []
    lib_mod = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200._lib")
    M.FUSED_BWD = True
    lib_mod.time_entry_point("pn2_mlp_bwd_layer")
    g_fused, x_fused = _grads(module, params + inputs, run)
    calls = lib_mod.timed_calls()
    lib_mod.time_entry_point(None)
    M.FUSED_BWD = False
    g_steps, x_steps = _grads(module, params + inputs, run)
    M.FUSED_BWD = True
    return g_fused, g_steps, len(calls), [n for n, _ in module.named_parameters()] + ["input%d" % i for i in range(len(inputs))]


@pytest.mark.parametrize("B,N,S,K,D,widths", [
    (2, 512, 128, 32, 9, [32, 32, 64]),        # sa1-like: K0 = 12
    (2, 256, 64, 32, 64, [64, 64, 128]),       # sa2-like: K0 = 67 (fixed-order partials for the first layer)
    (3, 200, 37, 8, 5, [16, 24, 32]),          # ragged M = 3 * 37 * 8 = 888, odd widths
])
def test_set_abstraction_fused_backward_matches_per_step_kernels(pn2, B, N, S, K, D, widths):
    torch.manual_seed(0)
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(S, 0.4, K, 3 + D, widths, False), 41).to(DEV).train()
    xyz = I.facade_xyz(B, N, 3).to(DEV).permute(0, 2, 1).contiguous()
    pts = torch.rand(B, D, N, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(True)
    wsel = torch.rand(B, widths[-1], S, generator=torch.Generator().manual_seed(2)).to(DEV)

    def run():
        torch.manual_seed(5)
        _, out = sa(xyz, pts)
        (out * wsel).sum().backward()

    g_f, g_s, n_calls, names = _ab(pn2, sa, [pts], run)
    assert n_calls == len(widths), "the fused kernel ran for %d of %d layers" % (n_calls, len(widths))
    for n, a, b in zip(names, g_f, g_s):
        if n.endswith("bias") and "convs" in n:
            continue
        assert _rel(a, b) <= 6e-3, (n, _rel(a, b))


@pytest.mark.parametrize("with_skip,need_dx", [(True, True), (False, True), (False, False)])
def test_feature_propagation_fused_backward_matches_per_step_kernels(pn2, with_skip, need_dx):
    B, N, S, D1, D2 = 2, 333, 64, (64 if with_skip else 0), 64
    fp = I.randomize_module_(pn2.PointNetFeaturePropagation(D1 + D2, [128, 128, 64]), 42).to(DEV).train()
    xyz1 = I.facade_xyz(B, N, 3).to(DEV).permute(0, 2, 1).contiguous()
    xyz2 = xyz1[:, :, :S].contiguous()
    p1 = torch.rand(B, D1, N, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(need_dx) if with_skip else None
    p2 = torch.rand(B, D2, S, generator=torch.Generator().manual_seed(2)).to(DEV).requires_grad_(need_dx)
    wsel = torch.rand(B, 64, N, generator=torch.Generator().manual_seed(3)).to(DEV)

    def run():
        out = fp(xyz1, xyz2, p1, p2)
        (out * wsel).sum().backward()

    inputs = [t for t in (p1, p2) if t is not None and t.requires_grad]
    g_f, g_s, n_calls, names = _ab(pn2, fp, inputs, run)
    assert n_calls == 3
    for n, a, b in zip(names, g_f, g_s):
        if n.endswith("bias") and "convs" in n:
            continue
        assert _rel(a, b) <= 6e-3, (n, _rel(a, b))


def test_pooled_top_layer_inside_the_fused_kernel(pn2):
    """da_mode 2 (modules.FUSED_BWD_POOLED, off by default: it measured slower): dZ of the max-pooled layer formed inside the
    fused kernel from (dOut, arg-max map, Z) must agree with the dense pn2_pool_bn_relu_bwd_dz pass"""
    M = _modules()
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(96, 0.4, 32, 3 + 9, [32, 32, 64], False), 44).to(DEV).train()
    xyz = I.facade_xyz(2, 512, 3).to(DEV).permute(0, 2, 1).contiguous()
    pts = torch.rand(2, 9, 512, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(True)
    wsel = torch.rand(2, 64, 96, generator=torch.Generator().manual_seed(2)).to(DEV)

    def run():
        torch.manual_seed(5)
        _, out = sa(xyz, pts)
        (out * wsel).sum().backward()

    params = list(sa.parameters()) + [pts]
    try:
        M.FUSED_BWD_POOLED = True
        g_pool, _ = _grads(sa, params, run)
    finally:
        M.FUSED_BWD_POOLED = False
    g_dense, _ = _grads(sa, params, run)
    for (n, _), a, b in zip(list(sa.named_parameters()) + [("dpoints", None)], g_pool, g_dense):
        if n.endswith("bias") and "convs" in n:
            continue
        assert _rel(a, b) <= 6e-3, (n, _rel(a, b))


def test_fused_backward_with_frozen_batchnorm_statistics(pn2):
    """eval-mode BatchNorm inside a differentiated forward: dZ = scale . mask . dA (no mean / zhat terms)"""
    fp = I.randomize_module_(pn2.PointNetFeaturePropagation(32 + 32, [64, 64]), 43).to(DEV).eval()
    B, N, S = 2, 256, 32
    xyz1 = I.facade_xyz(B, N, 3).to(DEV).permute(0, 2, 1).contiguous()
    xyz2 = xyz1[:, :, :S].contiguous()
    p1 = torch.rand(B, 32, N, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(True)
    p2 = torch.rand(B, 32, S, generator=torch.Generator().manual_seed(2)).to(DEV).requires_grad_(True)
    wsel = torch.rand(B, 64, N, generator=torch.Generator().manual_seed(3)).to(DEV)

    def run():
        (fp(xyz1, xyz2, p1, p2) * wsel).sum().backward()

    g_f, g_s, n_calls, names = _ab(pn2, fp, [p1, p2], run)
    assert n_calls == 2
    for n, a, b in zip(names, g_f, g_s):
        assert _rel(a, b) <= 6e-3, (n, _rel(a, b))


def test_fused_backward_vs_fp32_oracle_module_gradients(pn2):
    """sa + fp modules in bf16 with the fused backward vs the fp32 port of the reference: cosine >= 0.98 (the bf16 bar)"""
    B, N, S = 2, 512, 128
    x = I.facade_batch(B, N, 9, 5)
    sa = I.randomize_module_(pn2.PointNetSetAbstraction(S, 0.3, 32, 12, [32, 32, 64], False), 41).to(DEV).train()
    ref = I.randomize_module_(O.OracleSA(S, 0.3, 32, 12, [32, 32, 64], False), 41).train()
    wsel = torch.rand(B, 64, S, generator=torch.Generator().manual_seed(6))
    xd = x.to(DEV).transpose(2, 1)
    pts = xd.clone().requires_grad_(True)
    torch.manual_seed(51)
    _, out = sa(xd[:, :3, :], pts)
    (out * wsel.to(DEV)).sum().backward()
    xr = x.transpose(2, 1).clone().requires_grad_(True)
    torch.manual_seed(51)
    _, rout = ref(x.transpose(2, 1)[:, :3, :], xr)             # (coordinates carry no gradient: localfunctions.py never asks for one)
    (rout * wsel).sum().backward()
    cos = torch.nn.functional.cosine_similarity
    assert float(cos(pts.grad.cpu().flatten().double(), xr.grad.flatten().double(), dim=0)) >= 0.98
    for (n, p), (_, q) in zip(sa.named_parameters(), ref.named_parameters()):
        if n.endswith("bias") and "convs" in n:
            continue
        c = float(cos(p.grad.cpu().flatten().double(), q.grad.flatten().double(), dim=0))
        assert c >= 0.98, (n, c)


def test_fused_head_chain_uses_the_fused_backward(pn2):
    """fp1 + conv1/bn1 as a 4-layer chain behind the fused head: the head's bf16 gradient enters the fused kernel
    unmasked (da_mode 0)"""
    net = I.randomize_module_(pn2.get_model(18, 3), 61).to(DEV).train()
    net.drop1.p = 0.0
    x = I.facade_batch(2, 1024, 9, 3).to(DEV).transpose(2, 1)
    target = I.labels(2, 1024, 18, 7).to(DEV)
    w = torch.ones(18, device=DEV)
    params = list(net.parameters())

    def run():
        torch.manual_seed(9)
        loss, _, _ = net.forward_loss(x, target, w)
        loss.backward()

    g_f, g_s, n_calls, names = _ab(pn2, net, [], run)
    # sa1 (3) + sa2 (3) + sa3.2 (1) + fp1 + conv1 (4) take the fused kernel; the wide / few-row layers keep the per-step path
    assert n_calls >= 11, n_calls
    worst = max(_rel(a, b) for n, a, b in zip(names, g_f, g_s) if not (n.endswith("bias") and ("convs" in n or n == "conv1.bias")))
    # whole-network: two runs of the SAME bf16 code differ by the fp32 atomic order amplified through ReLU / max decisions
    # (bench.py data_parallel_check: 4e-3 .. 7e-3 relative L2 between ranks on one batch)
    assert worst <= 5e-2, worst
