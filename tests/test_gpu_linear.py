"""The MLP layer kernels through the C ABI against a torch fp32 reference of the same contraction
(this is the one floating-point kernel family, so it keeps a torch reference next to the oracle):
tcgen05 path (bf16 rows + weight scratch), FMA-pipe path (fp32 rows, or bf16 rows without scratch).

Tolerance: operands are rounded to bf16 exactly as the kernel does, products accumulate in fp32,
so the only differences are summation order (1e-5 relative) and the final bf16 rounding of the
stored output (2^-8 relative)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = [(1000, 12, 32), (647, 32, 64), (4096, 67, 64), (3000, 131, 128), (2048, 259, 256), (1024, 256, 512),
          (777, 768, 256), (513, 320, 256), (640, 64, 16), (130, 128, 128), (1, 32, 32)]


def _ld(k):
    return (k + 7) // 8 * 8


def _rows(M, K, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(M, _ld(K) if dtype == torch.bfloat16 else K)
    x[:, :K] = torch.randn(M, K, generator=g)
    return x.to(DEV).to(dtype)


@pytest.fixture(scope="module")
def lib(pn2):
    return importlib.import_module(pn2.__name__ + "._lib")


def _forward(lib, x, K, W, bias, scale, shift, stats, use_tc):
    M, N = x.shape[0], W.shape[0]
    ldz = _ld(N) if x.dtype == torch.bfloat16 else N
    z = torch.full((M, ldz), float("nan"), device=DEV, dtype=x.dtype)
    partials = torch.zeros(4, 2, N, device=DEV, dtype=torch.float64) if stats else None   # PN2_STAT_REPLICAS fp64 column-sum accumulators
    wpack = torch.empty(lib.load().pn2_linear_wpack_bytes(K, N), device=DEV, dtype=torch.uint8) if use_tc else None
    lib.call("pn2_linear_fwd", lib.ptr(x), x.shape[1], lib.dt(x), lib.ptr(scale), lib.ptr(shift), lib.ptr(W), lib.ptr(bias),
             M, K, N, lib.ptr(z), ldz, lib.dt(z), lib.ptr(partials), lib.ptr(wpack), lib.stream())
    return z, partials


def _reference(x, K, W, bias, scale, shift, bf16):
    a = x[:, :K].float()
    if scale is not None:
        a = torch.relu(a * scale + shift)
    w = W
    if bf16:
        a, w = a.bfloat16().float(), W.bfloat16().float()
    z = a.double() @ w.double().t()
    return (z + bias.double() if bias is not None else z)


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("mode", ["tc", "simt_bf16", "simt_fp32"])
@pytest.mark.parametrize("prologue", [False, True])
def test_linear_forward(lib, M, K, N, mode, prologue):
    dtype = torch.float32 if mode == "simt_fp32" else torch.bfloat16
    g = torch.Generator().manual_seed(M + K + N)
    x = _rows(M, K, dtype, 1)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV) if not prologue else None
    scale = (0.5 + torch.rand(K, generator=g)).to(DEV) if prologue else None
    shift = (torch.randn(K, generator=g) * 0.3).to(DEV) if prologue else None
    stats = prologue                                    # train mode: no bias, statistics on
    z, partials = _forward(lib, x, K, W, bias, scale, shift, stats, mode == "tc")
    want = _reference(x, K, W, bias, scale, shift, mode == "tc")     # only the tensor-core path rounds act(X), W to bf16
    got = z[:, :N].double()
    tol = 2.0 ** -8 if dtype == torch.bfloat16 else 1e-5
    err = ((got - want).abs() / (want.abs() + 1.0)).max().item()
    assert err <= tol, (mode, err)
    if z.shape[1] > N and mode == "tc":
        assert float(z[:, N:].float().abs().max()) == 0.0        # row padding is written as zeros
    if stats:
        s = partials.sum(0)
        ref_vals = got if mode == "tc" else want                # the tensor-core path sums the stored values
        np.testing.assert_allclose(s[0].cpu().numpy(), ref_vals.sum(0).cpu().numpy(), rtol=2e-3, atol=2e-3 * M ** 0.5)
        np.testing.assert_allclose(s[1].cpu().numpy(), (ref_vals ** 2).sum(0).cpu().numpy(), rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("mode", ["tc", "simt_fp32"])
def test_linear_backward_data(lib, M, K, N, mode):
    dtype = torch.float32 if mode == "simt_fp32" else torch.bfloat16
    g = torch.Generator().manual_seed(M * 3 + K + N)
    dz = _rows(M, N, dtype, 2)
    W = (torch.randn(N, K, generator=g) / N ** 0.5).to(DEV)
    ldd = _ld(K) if dtype == torch.bfloat16 else K
    dx = torch.full((M, ldd), float("nan"), device=DEV, dtype=dtype)
    wpack = torch.empty(lib.load().pn2_linear_wpack_bytes(N, K), device=DEV, dtype=torch.uint8) if mode == "tc" else None
    lib.call("pn2_linear_bwd_data", lib.ptr(dz), dz.shape[1], lib.dt(dz), lib.ptr(W), M, K, N, lib.ptr(dx), ldd, lib.dt(dx),
             lib.ptr(wpack), lib.stream())
    a, w = dz[:, :N].float(), W
    if dtype == torch.bfloat16:
        a, w = a.bfloat16().float(), W.bfloat16().float()
    want = a.double() @ w.double()
    err = ((dx[:, :K].double() - want).abs() / (want.abs() + 1.0)).max().item()
    assert err <= (2.0 ** -8 if dtype == torch.bfloat16 else 1e-5), (mode, err)


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_backward_weight(lib, M, K, N, dtype):
    g = torch.Generator().manual_seed(M + 7 * K + N)
    x, dz = _rows(M, K, dtype, 3), _rows(M, N, dtype, 4)
    scale = (0.5 + torch.rand(K, generator=g)).to(DEV)
    shift = (torch.randn(K, generator=g) * 0.3).to(DEV)
    dW = torch.full((N, K), float("nan"), device=DEV)
    scratch = torch.empty(lib.load().pn2_linear_wgrad_scratch_bytes(M, K, N), device=DEV, dtype=torch.uint8)
    lib.call("pn2_linear_bwd_weight", lib.ptr(dz), dz.shape[1], lib.dt(dz), lib.ptr(x), x.shape[1], lib.dt(x), lib.ptr(scale),
             lib.ptr(shift), M, K, N, lib.ptr(dW), lib.ptr(scratch), lib.stream())
    a = torch.relu(x[:, :K].float() * scale + shift)
    want = dz[:, :N].double().t() @ a.double()
    tol = 1e-4 if dtype == torch.float32 else 2e-2        # bf16: the tensor-core path rounds act(X) to bf16
    err = ((dW.double() - want).abs().max() / (want.abs().max() + 1e-9)).item()
    assert err <= tol, err


@pytest.mark.parametrize("G,ns,C", [(512, 32, 64), (100, 32, 128), (33, 16, 256), (7, 5, 24), (64, 32, 12), (3, 1, 512)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_bn_relu_max_tail(lib, G, ns, C, dtype):
    """max over nsample of relu(scale*z + shift) (pointnet2_utils.py:198-200) with the arg-max sample (first on ties) for the
    backward pass: vectorised bf16 kernel (C % 8 == 0) and the scalar kernel against torch on the same stored values."""
    g = torch.Generator().manual_seed(G + C)
    ld = _ld(C) if dtype == torch.bfloat16 else C
    Z = torch.zeros(G * ns, ld)
    Z[:, :C] = torch.randn(G * ns, C, generator=g)
    if ns > 3:
        Z.view(G, ns, ld)[:, 3] = Z.view(G, ns, ld)[:, 1]                   # exact ties: sample 3 repeats sample 1
    Z = Z.to(DEV).to(dtype)
    scale = (torch.rand(C, generator=g) - 0.3).to(DEV)                     # some negative gammas
    shift = (torch.randn(C, generator=g) * 0.2).to(DEV)
    out = torch.full((G, C), float("nan"), device=DEV)
    arg = torch.full((G, C), -1, device=DEV, dtype=torch.int32)
    lib.call("pn2_bn_relu_max", lib.ptr(Z), ld, lib.dt(Z), lib.ptr(scale), lib.ptr(shift), G, ns, C, lib.ptr(out), lib.ptr(arg),
             lib.stream())
    act = torch.relu(Z[:, :C].double() * scale.double() + shift.double()).view(G, ns, C)
    want, _ = act.max(dim=1)
    assert float((out.double() - want).abs().max()) <= 1e-5
    assert int(arg.min()) >= 0 and int(arg.max()) < ns
    picked = act.gather(1, arg.long().unsqueeze(1)).squeeze(1)
    assert float((picked - want).abs().max()) <= 1e-5                       # the recorded sample attains the maximum
    if ns > 3:
        assert not bool((arg == 3).any())                                   # the first of two equal samples wins


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("misalign", [0, 1])
def test_linear_backward_weight_accumulate(lib, M, K, N, misalign):
    """pn2_linear_bwd_weight_accum: dW += dZ^T act(X) with L2 reductions (vector form when dW is 16-byte aligned and
    K % 4 == 0, scalar otherwise); same tolerance as the fixed-order entry, on top of a non-zero starting value."""
    g = torch.Generator().manual_seed(M + 7 * K + N)
    x, dz = _rows(M, K, torch.bfloat16, 3), _rows(M, N, torch.bfloat16, 4)
    scale = (0.5 + torch.rand(K, generator=g)).to(DEV)
    shift = (torch.randn(K, generator=g) * 0.3).to(DEV)
    buf = torch.zeros(N * K + 4, device=DEV)
    dW = buf[misalign:misalign + N * K].view(N, K)
    start = torch.randn(N, K, generator=g).to(DEV)
    dW.copy_(start)
    for use_act in (True, False):
        lib.call("pn2_linear_bwd_weight_accum", lib.ptr(dz), dz.shape[1], lib.dt(dz), lib.ptr(x), x.shape[1], lib.dt(x),
                 lib.ptr(scale) if use_act else None, lib.ptr(shift) if use_act else None, M, K, N, lib.ptr(dW), lib.stream())
    a = torch.relu(x[:, :K].float() * scale + shift)
    want = dz[:, :N].double().t() @ a.double() + dz[:, :N].double().t() @ x[:, :K].double()
    err = ((dW.double() - start.double() - want).abs().max() / (want.abs().max() + 1e-9)).item()
    assert err <= 2e-2, err
    assert float(buf[:misalign].abs().sum()) == 0.0 and float(buf[misalign + N * K:].abs().sum()) == 0.0


@pytest.mark.parametrize("M,C", [(1000, 32), (4096, 64), (333, 128), (5000, 256), (640, 24), (8, 512)])
@pytest.mark.parametrize("mode", ["bf16", "f32dA", "pool", "fp32rows"])
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("ns", [8, 32, 5])
def test_bn_relu_backward_kernels(lib, M, C, mode, train, ns):
    """dgamma/dbeta reduction (fp64 accumulators) and dz of BN(train or frozen)+ReLU against the formulas of
    include/pn2b200.h, dense and pooled (arg-max routed) variants, vectorised bf16 and scalar paths."""
    if mode != "pool" and ns != 8:
        pytest.skip("nsample only matters for the pooled variants")
    if M < ns:
        M = 2 * ns
    g = torch.Generator().manual_seed(M + C)
    zt = torch.float32 if mode == "fp32rows" else torch.bfloat16
    ldz = C if zt == torch.float32 else _ld(C)
    Z = torch.zeros(M, ldz, dtype=zt, device=DEV)
    Z[:, :C] = torch.randn(M, C, generator=g).to(DEV).to(zt)
    scale = (torch.rand(C, generator=g) - 0.3).to(DEV)           # some negative gammas
    shift = (torch.randn(C, generator=g) * 0.2).to(DEV)
    mean = (torch.randn(C, generator=g) * 0.1).to(DEV)
    invstd = (0.5 + torch.rand(C, generator=g)).to(DEV)
    accum = torch.zeros(4, 2, C, dtype=torch.float64, device=DEV)
    zf = Z[:, :C].float()
    mask = (zf * scale + shift) > 0
    if mode == "pool":
        G = M // ns
        M = G * ns
        Z, zf, mask = Z[:M], zf[:M], mask[:M]
        dOut = torch.randn(G, C, generator=g).to(DEV)
        arg = torch.randint(0, ns, (G, C), generator=g, dtype=torch.int32).to(DEV)
        dense = torch.zeros(G, ns, C, device=DEV)
        dense.scatter_(1, arg.long().unsqueeze(1), dOut.unsqueeze(1))
        gfull = dense.view(M, C)
        lib.call("pn2_pool_bn_relu_bwd_reduce", lib.ptr(dOut), lib.ptr(arg), lib.ptr(Z), ldz, lib.dt(Z), lib.ptr(scale),
                 lib.ptr(shift), lib.ptr(mean), lib.ptr(invstd), G, ns, C, lib.ptr(accum), lib.stream())
    else:
        at = torch.float32 if mode in ("f32dA", "fp32rows") else torch.bfloat16
        ldda = C if at == torch.float32 else _ld(C)
        dA = torch.zeros(M, ldda, dtype=at, device=DEV)
        dA[:, :C] = torch.randn(M, C, generator=g).to(DEV).to(at)
        gfull = dA[:, :C].float()
        lib.call("pn2_bn_relu_bwd_reduce", lib.ptr(dA), ldda, lib.dt(dA), lib.ptr(Z), ldz, lib.dt(Z), lib.ptr(scale),
                 lib.ptr(shift), lib.ptr(mean), lib.ptr(invstd), M, C, lib.ptr(accum), lib.stream())
    gm = torch.where(mask, gfull, torch.zeros_like(gfull)).double()
    zhat = ((zf - mean) * invstd).double()
    want_dbeta, want_dgamma = gm.sum(0), (gm * zhat).sum(0)
    dgb = torch.empty(2, C, device=DEV)
    lib.call("pn2_bn_bwd_finalize", lib.ptr(accum), C, lib.ptr(dgb[0]), lib.ptr(dgb[1]), lib.stream())
    assert float(accum.abs().max()) == 0.0                       # finalize leaves the accumulator clean
    tol = 1e-3 * max(1.0, M ** 0.5)
    np.testing.assert_allclose(dgb[1].cpu().numpy(), want_dbeta.cpu().numpy(), rtol=1e-4, atol=tol)
    np.testing.assert_allclose(dgb[0].cpu().numpy(), want_dgamma.cpu().numpy(), rtol=1e-4, atol=tol)
    dzt = zt
    dZ = torch.full((M, ldz), float("nan"), dtype=dzt, device=DEV)
    mu_p, is_p = (lib.ptr(mean), lib.ptr(invstd)) if train else (None, None)
    if mode == "pool":
        lib.call("pn2_pool_bn_relu_bwd_dz", lib.ptr(dOut), lib.ptr(arg), lib.ptr(Z), ldz, lib.dt(Z), lib.ptr(scale),
                 lib.ptr(shift), mu_p, is_p, lib.ptr(dgb[0]), lib.ptr(dgb[1]), G, ns, C, lib.ptr(dZ), ldz, lib.dt(dZ),
                 lib.stream())
    else:
        lib.call("pn2_bn_relu_bwd_dz", lib.ptr(dA), ldda, lib.dt(dA), lib.ptr(Z), ldz, lib.dt(Z), lib.ptr(scale),
                 lib.ptr(shift), mu_p, is_p, lib.ptr(dgb[0]), lib.ptr(dgb[1]), M, C, lib.ptr(dZ), ldz, lib.dt(dZ),
                 lib.stream())
    if train:
        want = scale.double() * (gm - dgb[1].double() / M - zhat * dgb[0].double() / M)
    else:
        want = scale.double() * gm
    err = ((dZ[:, :C].double() - want).abs() / (want.abs() + 1.0)).max().item()
    assert err <= (2.0 ** -7 if dzt == torch.bfloat16 else 1e-5), err
    if mode == "bf16":                       # in place (dZ aliases dA), as modules.mlp_backward calls it: same bits
        lib.call("pn2_bn_relu_bwd_dz", lib.ptr(dA), ldda, lib.dt(dA), lib.ptr(Z), ldz, lib.dt(Z), lib.ptr(scale),
                 lib.ptr(shift), mu_p, is_p, lib.ptr(dgb[0]), lib.ptr(dgb[1]), M, C, lib.ptr(dA), ldda, lib.dt(dA),
                 lib.stream())
        assert torch.equal(dA[:, :C], dZ[:, :C])


# ---- the launch-saving variants: one pack launch for many weights, prepacked layer calls with the train-mode
#      BatchNorm finalize fused in ("last CTA finalizes"), reductions with their finalize fused in -------------------
import ctypes


@pytest.mark.parametrize("M,K,N", [(1000, 12, 32), (4096, 67, 64), (2048, 259, 256), (1024, 256, 512), (777, 768, 256),
                                   (130000, 32, 64)])
def test_prepacked_forward_with_fused_finalize_matches_two_step_path(lib, M, K, N):
    g = torch.Generator().manual_seed(M + K)
    x = _rows(M, K, torch.bfloat16, 5)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    gamma, beta = (0.5 + torch.rand(N, generator=g)).to(DEV), torch.randn(N, generator=g).to(DEV)
    scale_in = (0.5 + torch.rand(K, generator=g)).to(DEV)
    shift_in = (torch.randn(K, generator=g) * 0.3).to(DEV)
    L = lib.load()
    ldz = _ld(N)

    def buffers():
        return dict(z=torch.full((M, ldz), float("nan"), device=DEV, dtype=torch.bfloat16),
                    accum=torch.zeros(4, 2, N, device=DEV, dtype=torch.float64), out=torch.zeros(4, N, device=DEV),
                    rm=torch.full((N,), 0.25, device=DEV), rv=torch.full((N,), 2.0, device=DEV),
                    nbt=torch.tensor(3, device=DEV, dtype=torch.int64))

    a = buffers()           # reference: pn2_linear_fwd (packs inside) + pn2_bn_train_finalize
    wp = torch.empty(L.pn2_linear_wpack_bytes(K, N), device=DEV, dtype=torch.uint8)
    lib.call("pn2_linear_fwd", lib.ptr(x), x.shape[1], lib.dt(x), lib.ptr(scale_in), lib.ptr(shift_in), lib.ptr(W), None, M,
             K, N, lib.ptr(a["z"]), ldz, 1, lib.ptr(a["accum"]), lib.ptr(wp), lib.stream())
    lib.call("pn2_bn_train_finalize", lib.ptr(a["accum"]), M, N, lib.ptr(gamma), lib.ptr(beta), lib.ptr(bias), 1e-5, 0.1,
             lib.ptr(a["rm"]), lib.ptr(a["rv"]), lib.ptr(a["out"][0]), lib.ptr(a["out"][1]), lib.ptr(a["out"][2]),
             lib.ptr(a["out"][3]), lib.ptr(a["nbt"]), lib.stream())
    b = buffers()           # one pack launch (both orientations), then the prepacked call with the fused finalize
    img_f = torch.empty(L.pn2_linear_wpack_bytes(K, N), device=DEV, dtype=torch.uint8)
    img_b = torch.empty(L.pn2_linear_wpack_bytes(N, K), device=DEV, dtype=torch.uint8)
    vp, ci = ctypes.c_void_p * 2, ctypes.c_int * 2
    lib.call("pn2_pack_weights", 2, vp(W.data_ptr(), W.data_ptr()), ci(K, K), ci(N, N), ci(0, 1),
             vp(img_f.data_ptr(), img_b.data_ptr()), lib.stream())
    assert torch.equal(img_f, wp)                                    # same image as the in-call packing
    ticket = torch.zeros(4, device=DEV, dtype=torch.int32)
    for rep in range(2):                                             # the ticket and the accumulator are self-cleaning
        fin = lib.BnFinalize(lib.ptr(ticket), lib.ptr(gamma), lib.ptr(beta), lib.ptr(bias), 1e-5, 0.1, lib.ptr(b["rm"]),
                             lib.ptr(b["rv"]), lib.ptr(b["out"][0]), lib.ptr(b["out"][1]), lib.ptr(b["out"][2]),
                             lib.ptr(b["out"][3]), lib.ptr(b["nbt"]))
        lib.call("pn2_linear_fwd_prepacked", lib.ptr(x), x.shape[1], lib.dt(x), lib.ptr(scale_in), lib.ptr(shift_in),
                 lib.ptr(W), None, M, K, N, lib.ptr(b["z"]), ldz, 1, lib.ptr(b["accum"]), lib.ptr(img_f),
                 ctypes.addressof(fin), lib.stream())
        torch.cuda.synchronize()
        assert int(ticket[0]) == 0 and float(b["accum"].abs().sum()) == 0.0
        if rep == 0:
            assert torch.equal(a["z"], b["z"])
            for k in ("out", "rm", "rv"):
                assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-7), k
            assert int(a["nbt"]) == int(b["nbt"]) == 4
    assert int(b["nbt"]) == 5
    # data gradient through the transposed image of the same pack launch
    dz = _rows(M, N, torch.bfloat16, 6)
    ldd = _ld(K)
    d1 = torch.full((M, ldd), float("nan"), device=DEV, dtype=torch.bfloat16)
    d2 = torch.full((M, ldd), float("nan"), device=DEV, dtype=torch.bfloat16)
    wp2 = torch.empty(L.pn2_linear_wpack_bytes(N, K), device=DEV, dtype=torch.uint8)
    lib.call("pn2_linear_bwd_data", lib.ptr(dz), dz.shape[1], 1, lib.ptr(W), M, K, N, lib.ptr(d1), ldd, 1, lib.ptr(wp2), lib.stream())
    lib.call("pn2_linear_bwd_data_prepacked", lib.ptr(dz), dz.shape[1], 1, lib.ptr(W), M, K, N, lib.ptr(d2), ldd, 1,
             lib.ptr(img_b), lib.stream())
    assert torch.equal(d1, d2)


@pytest.mark.parametrize("M,C,da", [(5000, 32, torch.bfloat16), (4096, 128, torch.float32), (777, 24, torch.float32),
                                    (60000, 256, torch.bfloat16)])
def test_bn_backward_reduce_with_fused_finalize(lib, M, C, da):
    g = torch.Generator().manual_seed(C)
    ld = _ld(C)
    z = _rows(M, C, torch.bfloat16, 8)
    dA = _rows(M, C, da, 9) if da == torch.bfloat16 else torch.randn(M, C, generator=g).to(DEV)
    scale, shift = (0.5 + torch.rand(C, generator=g)).to(DEV), (torch.randn(C, generator=g) * 0.2).to(DEV)
    mean, invstd = (torch.randn(C, generator=g) * 0.1).to(DEV), (0.5 + torch.rand(C, generator=g)).to(DEV)
    acc1, acc2 = (torch.zeros(4, 2, C, device=DEV, dtype=torch.float64) for _ in range(2))
    want, got = torch.zeros(2, C, device=DEV), torch.zeros(2, C, device=DEV)
    lib.call("pn2_bn_relu_bwd_reduce", lib.ptr(dA), dA.shape[1], lib.dt(dA), lib.ptr(z), ld, 1, lib.ptr(scale), lib.ptr(shift),
             lib.ptr(mean), lib.ptr(invstd), M, C, lib.ptr(acc1), lib.stream())
    lib.call("pn2_bn_bwd_finalize", lib.ptr(acc1), C, lib.ptr(want[0]), lib.ptr(want[1]), lib.stream())
    ticket = torch.zeros(4, device=DEV, dtype=torch.int32)
    for _ in range(2):
        lib.call("pn2_bn_relu_bwd_reduce_finalize", lib.ptr(dA), dA.shape[1], lib.dt(dA), lib.ptr(z), ld, 1, lib.ptr(scale),
                 lib.ptr(shift), lib.ptr(mean), lib.ptr(invstd), M, C, lib.ptr(acc2), lib.ptr(ticket), lib.ptr(got[0]),
                 lib.ptr(got[1]), lib.stream())
        torch.cuda.synchronize()
        assert int(ticket[0]) == 0 and float(acc2.abs().sum()) == 0.0
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
