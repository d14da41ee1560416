"""csrc/head.cu -- the segmentation head's tail on rows (relu(bn) -> dropout -> conv2 -> log_softmax) and its backward,
against a torch fp32/fp64 statement of /root/reference/models/pointnet2_sem_seg.py:36-39 on the same bf16 inputs, and
the whole fp1+head chain (modules.PointNetFeaturePropagation.forward_with_head) against the PyTorch head.

Tolerances: the kernels read bf16 Z, multiply on mma.sync with bf16 hi+lo split operands (three products, fp32
accumulation: 2^-16 relative per term) -> log-probabilities to 1e-4 (absolute) of the fp64 torch evaluation of the same
bf16 inputs; dA is stored as bf16 (2^-8 relative).  The fused-loss entry points (pn2_head_tail_loss_fwd/_bwd) are checked
against F.nll_loss(weight=..., ignore_index=-100) in fp64 and against the dense-gradient path."""
import importlib

import numpy as np
import pytest
import torch

import _inputs as I

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def lib(pn2):
    return importlib.import_module(pn2.__name__ + "._lib")


def _case(M, C, NC, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(M, C, generator=g).to(DEV).bfloat16()
    scale, shift = (0.5 + torch.rand(C, generator=g)).to(DEV), (torch.randn(C, generator=g) * 0.3).to(DEV)
    W2, b2 = (torch.randn(NC, C, generator=g) / C ** 0.5).to(DEV), torch.randn(NC, generator=g).to(DEV)
    return z, scale, shift, W2, b2


@pytest.mark.parametrize("M,C,NC", [(1000, 128, 18), (4096, 128, 13), (513, 64, 32), (77, 256, 2), (131072, 128, 18)])
def test_head_tail_forward_backward_no_dropout(lib, M, C, NC):
    z, scale, shift, W2, b2 = _case(M, C, NC, M + NC)
    logp = torch.empty(M, NC, device=DEV)
    act = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    labels = torch.full((M,), -1, device=DEV, dtype=torch.int64)
    lib.call("pn2_head_tail_fwd", lib.ptr(z), C, lib.ptr(scale), lib.ptr(shift), lib.ptr(W2), lib.ptr(b2), M, C, NC, 0.0,
             None, lib.ptr(logp), lib.ptr(act), C, lib.ptr(labels), lib.stream())
    assert torch.equal(labels, logp.argmax(dim=1))                      # fused arg-max == torch.argmax of the stored values
    a = torch.relu(z.double() * scale.double() + shift.double()).requires_grad_(True)
    want = torch.log_softmax(a @ W2.double().t() + b2.double(), dim=1)
    assert float((logp.double() - want).abs().max()) < 1e-4
    assert float((act.double() - a.detach()).abs().max()) <= 2.0 ** -8 * float(a.abs().max())
    g = torch.Generator().manual_seed(3)
    dlogp = torch.randn(M, NC, generator=g).to(DEV)
    want.backward(dlogp.double())
    dA = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    lddl = (NC + 7) // 8 * 8
    dl = torch.full((M, lddl), float("nan"), device=DEV, dtype=torch.bfloat16)
    accum = torch.zeros(64, device=DEV, dtype=torch.float64)
    db2 = torch.empty(NC, device=DEV)
    lib.call("pn2_head_tail_bwd", lib.ptr(dlogp), lib.ptr(logp), lib.ptr(W2), M, C, NC, 0.0, None, lib.ptr(dA), C, lib.ptr(dl),
             lddl, lib.ptr(accum), lib.ptr(db2), lib.stream())
    dlogits = dlogp.double() - want.detach().exp() * dlogp.double().sum(1, keepdim=True)
    err = (dA.double() - a.grad).abs().max() / a.grad.abs().max()
    assert float(err) <= 2.0 ** -7
    assert float((dl[:, :NC].double() - dlogits).abs().max()) <= 2.0 ** -7 * float(dlogits.abs().max())
    assert float(dl[:, NC:].float().abs().max() if lddl > NC else 0.0) == 0.0
    assert torch.allclose(db2.double(), dlogits.sum(0), rtol=1e-4, atol=1e-3)
    assert float(accum.abs().sum()) == 0.0


def test_head_tail_dropout_is_consistent_and_unbiased(lib):
    M, C, NC = 20000, 128, 18
    z, scale, shift, W2, b2 = _case(M, C, NC, 5)
    shift = shift + 3.0                                     # keep most pre-activations positive
    outs = []
    for seed_val in (1234, 1234, 99):
        seed = torch.tensor([seed_val], device=DEV, dtype=torch.int64)
        logp = torch.empty(M, NC, device=DEV)
        act = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
        lib.call("pn2_head_tail_fwd", lib.ptr(z), C, lib.ptr(scale), lib.ptr(shift), lib.ptr(W2), lib.ptr(b2), M, C, NC, 0.5,
                 lib.ptr(seed), lib.ptr(logp), lib.ptr(act), C, None, lib.stream())
        outs.append((logp, act, seed))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])      # same seed, same mask
    assert not torch.equal(outs[0][1], outs[2][1])                                           # another seed, another mask
    logp, act, seed = outs[0]
    a = torch.relu(z.float() * scale + shift)
    positive = a > 0.05
    kept = (act.float() != 0) & positive
    frac = float(kept.sum()) / float(positive.sum())
    assert abs(frac - 0.5) < 0.005, frac                                                      # keep probability 1 - p
    ratio = act.float()[kept] / a[kept]
    assert float((ratio - 2.0).abs().max()) < 0.02                                            # kept values scaled by 1/(1-p)
    per_col = kept.float().sum(0) / positive.float().sum(0).clamp(min=1)
    assert float((per_col - 0.5).abs().max()) < 0.03                                          # no column bias
    want = torch.log_softmax(act.float() @ W2.t() + b2, dim=1)
    assert float((logp - want).abs().max()) < 0.1      # the kernel uses the un-rounded fp32 activation; `act` is its bf16 copy
    # backward applies the SAME mask
    dlogp = torch.randn(M, NC, generator=torch.Generator().manual_seed(1)).to(DEV)
    dA = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    accum = torch.zeros(64, device=DEV, dtype=torch.float64)
    db2 = torch.empty(NC, device=DEV)
    lib.call("pn2_head_tail_bwd", lib.ptr(dlogp), lib.ptr(logp), lib.ptr(W2), M, C, NC, 0.5, lib.ptr(seed), lib.ptr(dA), C, None,
             0, lib.ptr(accum), lib.ptr(db2), lib.stream())
    dropped = (act.float() == 0) & positive
    assert float(dA.float()[dropped].abs().max()) == 0.0
    assert float((dA.float()[kept] != 0).float().mean()) > 0.99


def test_fp1_with_fused_head_matches_pytorch_head(pn2):
    """Whole chain: PointNetFeaturePropagation.forward_with_head vs the module followed by the PyTorch head (dropout off),
    forward and every gradient (bf16 rows in both; conv1/conv2 in the PyTorch head run in fp32/TF32)."""
    pn2.set_precision("bf16")
    try:
        net_a = I.randomize_module_(pn2.get_model(18, 3), 61).to(DEV).train()
        net_b = I.randomize_module_(pn2.get_model(18, 3), 61).to(DEV).train()
        net_a.drop1.p = net_b.drop1.p = 0.0
        net_a.fused_head, net_b.fused_head = True, False
        x = I.facade_batch(2, 2048, 9, 3).to(DEV).transpose(2, 1)
        target = I.labels(2, 2048, 18, 7).to(DEV)
        w = torch.linspace(0.5, 1.5, 18).to(DEV)
        res = []
        for net in (net_a, net_b):
            torch.manual_seed(72)
            pred, _ = net(x)
            loss = pn2.get_loss()(pred.contiguous().view(-1, 18), target, None, w)
            loss.backward()
            res.append((pred.detach(), float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters()}))
        (pa, la, ga), (pb, lb, gb) = res
        assert pa.shape == pb.shape == (2, 2048, 18)
        assert float((pa - pb).abs().max()) < 0.05 and abs(la - lb) < 5e-3
        for n in ga:
            if n.endswith("bias") and ("mlp_convs" in n or n == "conv1.bias"):
                continue                                   # cancelled by train-mode batch norm (rounding noise only)
            a, b = ga[n].double().flatten(), gb[n].double().flatten()
            cos = float(a @ b / (a.norm() * b.norm() + 1e-30))
            assert cos > 0.98, (n, cos)
        for key in ("conv2.weight", "conv2.bias", "conv1.weight", "bn1.weight", "bn1.bias"):
            a, b = ga[key].double(), gb[key].double()
            assert float((a - b).norm() / (b.norm() + 1e-30)) < 0.05, key
        assert int(net_a.bn1.num_batches_tracked) == int(net_b.bn1.num_batches_tracked) == 1
        assert torch.allclose(net_a.bn1.running_mean, net_b.bn1.running_mean, atol=2e-3)
        # eval mode: same chain with running statistics, no dropout
        net_a.eval(), net_b.eval()
        with torch.no_grad():
            torch.manual_seed(5)
            ea, _ = net_a(x)
            torch.manual_seed(5)
            eb, _ = net_b(x)
        assert float((ea - eb).abs().max()) < 0.05
    finally:
        pn2.set_precision("fp32")


@pytest.mark.parametrize("M,C,NC,weighted", [(1000, 128, 18, True), (4099, 128, 13, False), (513, 64, 32, True), (77, 256, 2, True),
                                              (131072, 128, 18, True)])
def test_head_tail_fused_loss_matches_nll_loss(lib, M, C, NC, weighted):
    """pn2_head_tail_loss_fwd/_bwd == pn2_head_tail_fwd -> F.nll_loss(weight, ignore_index=-100) -> pn2_head_tail_bwd."""
    z, scale, shift, W2, b2 = _case(M, C, NC, 7 * M + NC)
    g = torch.Generator().manual_seed(M)
    target = torch.randint(0, NC, (M,), generator=g)
    target[torch.rand(M, generator=g) < 0.1] = -100                     # F.nll_loss's default ignore_index
    target = target.to(DEV)
    cw = (0.25 + torch.rand(NC, generator=g)).to(DEV) if weighted else None
    logp_a, logp_b = torch.empty(M, NC, device=DEV), torch.empty(M, NC, device=DEV)
    act_a, act_b = (torch.empty(M, C, device=DEV, dtype=torch.bfloat16) for _ in range(2))
    lib.call("pn2_head_tail_fwd", lib.ptr(z), C, lib.ptr(scale), lib.ptr(shift), lib.ptr(W2), lib.ptr(b2), M, C, NC, 0.0,
             None, lib.ptr(logp_a), lib.ptr(act_a), C, None, lib.stream())
    accum = torch.zeros(64, device=DEV, dtype=torch.float64)
    loss = torch.full((2,), float("nan"), device=DEV)
    lib.call("pn2_head_tail_loss_fwd", lib.ptr(z), C, lib.ptr(scale), lib.ptr(shift), lib.ptr(W2), lib.ptr(b2), M, C, NC, 0.0,
             None, lib.ptr(target), lib.ptr(cw), lib.ptr(logp_b), lib.ptr(act_b), C, lib.ptr(accum), lib.ptr(loss), lib.stream())
    assert torch.equal(logp_a, logp_b) and torch.equal(act_a, act_b)
    assert float(accum.abs().sum()) == 0.0
    lp = logp_a.double().requires_grad_(True)
    want = torch.nn.functional.nll_loss(lp, target, weight=None if cw is None else cw.double())
    assert abs(float(loss[0]) - float(want)) <= 1e-5 * abs(float(want)) + 1e-6
    wsum = float((cw[target.clamp(min=0)] * (target >= 0)).sum()) if weighted else float((target >= 0).sum())
    assert abs(float(loss[1]) - wsum) <= 1e-5 * wsum
    # backward with dL/dloss = 0.7: dense path fed with autograd's gradient of nll_loss vs the fused path
    (0.7 * want).backward()
    dlogp = lp.grad.float().contiguous()
    lddl = (NC + 7) // 8 * 8
    outs = []
    for fused in (False, True):
        dA = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
        dl = torch.full((M, lddl), float("nan"), device=DEV, dtype=torch.bfloat16)
        db2 = torch.empty(NC, device=DEV)
        if fused:
            dloss = torch.tensor(0.7, device=DEV)
            lib.call("pn2_head_tail_loss_bwd", lib.ptr(logp_a), lib.ptr(target), lib.ptr(cw), lib.ptr(loss), lib.ptr(dloss),
                     lib.ptr(W2), M, C, NC, 0.0, None, lib.ptr(dA), C, lib.ptr(dl), lddl, lib.ptr(accum), lib.ptr(db2),
                     lib.stream())
        else:
            lib.call("pn2_head_tail_bwd", lib.ptr(dlogp), lib.ptr(logp_a), lib.ptr(W2), M, C, NC, 0.0, None, lib.ptr(dA), C,
                     lib.ptr(dl), lddl, lib.ptr(accum), lib.ptr(db2), lib.stream())
        outs.append((dA.double(), dl.double(), db2.double()))
    assert float(accum.abs().sum()) == 0.0
    (dA0, dl0, db0), (dA1, dl1, db1) = outs
    dlogits = dlogp.double() - logp_a.double().exp() * dlogp.double().sum(1, keepdim=True)      # exact statement
    assert float((dl1[:, :NC] - dlogits).abs().max()) <= 2.0 ** -7 * float(dlogits.abs().max())
    assert float(dl1[:, NC:].abs().max() if lddl > NC else 0.0) == 0.0
    want_dA = dlogits @ W2.double()
    assert float((dA1 - want_dA).abs().max()) <= 2.0 ** -7 * float(want_dA.abs().max())
    assert float((dA1 - dA0).abs().max()) <= 2.0 ** -7 * float(want_dA.abs().max())
    assert torch.allclose(db1, dlogits.sum(0), rtol=1e-4, atol=1e-6)
    ignored = (target < 0)
    assert float(dA1[ignored].abs().max()) == 0.0 and float(dl1[ignored].abs().max()) == 0.0     # ignored points: no gradient


def test_forward_loss_matches_forward_plus_get_loss(pn2):
    """get_model.forward_loss (loss inside the head kernels) == forward() + get_loss on the same network: loss value and
    every parameter gradient (both run the fused head; the only difference is where the loss gradient is formed)."""
    pn2.set_precision("bf16")
    try:
        nets = [I.randomize_module_(pn2.get_model(18, 3), 61).to(DEV).train() for _ in range(2)]
        for net in nets:
            net.drop1.p = 0.0
        x = I.facade_batch(2, 2048, 9, 3).to(DEV).transpose(2, 1)
        target = I.labels(2, 2048, 18, 7).to(DEV)
        target[5] = -100
        w = torch.linspace(0.5, 1.5, 18).to(DEV)
        torch.manual_seed(72)
        loss_a, pred_a, _ = nets[0].forward_loss(x, target, w)
        assert not pred_a.requires_grad and loss_a.requires_grad
        loss_a.backward()
        torch.manual_seed(72)
        pred_b, _ = nets[1](x)
        loss_b = pn2.get_loss()(pred_b.contiguous().view(-1, 18), target, None, w)
        loss_b.backward()
        assert float((pred_a - pred_b).abs().max()) < 1e-5
        assert abs(float(loss_a) - float(loss_b)) < 1e-5
        for (n, pa), (_, pb) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
            if n.endswith("bias") and ("mlp_convs" in n or n == "conv1.bias"):
                continue                                   # cancelled by train-mode batch norm (rounding noise only)
            # the two paths round dlogits (bf16 rows) at different points: ~2^-9 relative noise per element, which
            # survives cancellation differently per tensor -- weights to 5 %, everything to cosine 0.99
            a, b = pa.grad.double().flatten(), pb.grad.double().flatten()
            assert float(a @ b / (a.norm() * b.norm() + 1e-30)) > 0.99, n
            if pa.dim() > 1:
                assert float((a - b).norm() / (b.norm() + 1e-30)) < 5e-2, n
        # eval / no_grad: same value, no graph
        nets[0].eval()
        with torch.no_grad():
            torch.manual_seed(5)
            loss_e, pred_e, _ = nets[0].forward_loss(x, target, w)
            want = pn2.get_loss()(pred_e.contiguous().view(-1, 18), target, None, w)
        assert abs(float(loss_e) - float(want)) < 1e-5
    finally:
        pn2.set_precision("fp32")


def test_head_tail_argmax_takes_first_maximum_on_ties(lib):
    """Rows whose logits tie exactly (identical W2 rows / zero activations): labels == torch.argmax (lowest index)."""
    M, C, NC = 300, 64, 18
    z, scale, shift, W2, b2 = _case(M, C, NC, 11)
    W2[7] = W2[3]
    W2[12] = W2[3]
    b2[7] = b2[12] = b2[3] = 2.5                      # classes 3, 7, 12 always tie, and usually win
    z[100:200] = -50.0                               # relu -> all-zero activations: logits == b2 exactly
    logp = torch.empty(M, NC, device=DEV)
    labels = torch.full((M,), -1, device=DEV, dtype=torch.int64)
    lib.call("pn2_head_tail_fwd", lib.ptr(z), C, lib.ptr(scale), lib.ptr(shift), lib.ptr(W2), lib.ptr(b2), M, C, NC, 0.0,
             None, lib.ptr(logp), None, 0, lib.ptr(labels), lib.stream())
    assert torch.equal(labels, logp.argmax(dim=1))
    assert int((labels == 3).sum()) > 100 and int((labels == 7).sum()) == 0 and int((labels == 12).sum()) == 0
