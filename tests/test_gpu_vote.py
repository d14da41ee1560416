"""SURVEY.md 8(f) n1 -- vote accumulation and final arg-max on the device (csrc/vote.cu) against fixtures written by the
reference's own add_vote (/root/reference/localfunctions.py:336-343, tests/golden/make_golden.py votes) and against the
numpy restatement in oracle/pn2_oracle.py; integer counts, so everything is bit-exact."""
import numpy as np
import pytest
import torch

import _inputs as I
from oracle import pn2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("tag,NC", [("small", 18), ("wide", 5), ("dense", 18)])
@pytest.mark.parametrize("wdtype", [torch.float64, torch.float32])
def test_add_vote_matches_reference_fixture(pn2, golden, tag, NC, wdtype):
    v = golden("votes")
    want = v[tag + "_pool"]
    pool = pn2.new_vote_pool(want.shape[0], NC, DEV)
    for it in range(v[tag + "_idx"].shape[0]):
        idx = torch.from_numpy(v[tag + "_idx"][it].astype(np.float64))        # float64 indices, as the reference holds them
        lab = torch.from_numpy(v[tag + "_lab"][it].astype(np.int64)).to(DEV)
        w = torch.from_numpy(v[tag + "_w"][it]).to(wdtype)
        out = pn2.add_vote(pool, idx, lab, w)
        assert out is pool
    assert np.array_equal(pool.cpu().numpy(), want)
    assert np.array_equal(pn2.vote_argmax(pool).cpu().numpy(), v[tag + "_labels"].astype(np.int64))
    assert np.array_equal(pn2.vote_argmax(pool, torch.uint8).cpu().numpy(), v[tag + "_labels"])


def test_add_vote_large_random_matches_oracle(pn2):
    g = np.random.RandomState(5)
    P, NC, B, N = 200_000, 18, 64, 4096
    idx = g.randint(0, P, size=(B, N))
    lab = g.randint(0, NC, size=(B, N))
    w = g.rand(B, N).astype(np.float32)
    w[g.rand(B, N) < 0.3] = 0.0
    pool = pn2.new_vote_pool(P, NC, DEV)
    pn2.add_vote(pool, torch.from_numpy(idx), torch.from_numpy(lab), torch.from_numpy(w))
    pn2.add_vote(pool, torch.from_numpy(idx[:7]).to(DEV), torch.from_numpy(lab[:7]).to(DEV), None)   # no weights: all pairs vote
    want = O.add_vote(np.zeros((P, NC)), idx, lab, w)
    want = O.add_vote(want, idx[:7], lab[:7], np.ones((7, N)))
    assert np.array_equal(pool.cpu().numpy(), want.astype(np.int32))
    assert np.array_equal(pn2.vote_argmax(pool).cpu().numpy(), O.vote_argmax(want))
    assert int(pool.sum()) == int((w != 0).sum()) + 7 * N                     # every voting pair counted exactly once


def test_add_vote_argument_checks(pn2):
    pool = pn2.new_vote_pool(10, 4, DEV)
    idx = torch.zeros(2, 8, dtype=torch.int64)
    with pytest.raises(ValueError):
        pn2.add_vote(pool, idx, torch.zeros(2, 7, dtype=torch.int64))
    with pytest.raises(TypeError):
        pn2.add_vote(pool.float(), idx, idx)
    with pytest.raises(ValueError):
        pn2.add_vote(pool.cpu(), idx, idx)
    # out-of-range pairs (an IndexError in the reference) are skipped, in-range ones still count
    pn2.add_vote(pool, torch.tensor([[0, 10, -1, 3]]), torch.tensor([[1, 1, 1, 4]]))
    assert int(pool.sum()) == 1 and int(pool[0, 1]) == 1
    empty = pn2.new_vote_pool(0, 4, DEV)
    assert pn2.vote_argmax(empty).numel() == 0


def test_predict_scene_equals_blocks_then_host_votes(pn2):
    """predict_scene (device votes, pipelined graph, rank shards + merged pools) == predict_blocks labels voted on the
    host by the oracle's add_vote, on a synthetic scene whose blocks overlap."""
    pn2.set_precision("bf16")
    B, N, C, NC, P = 4, 1024, 9, 18, 3000
    torch.manual_seed(3)
    net = I.randomize_module_(pn2.get_model(NC, C - 6), 23).to(DEV).eval()
    nb = 11
    blocks = torch.cat([I.facade_batch(B, N, C, 900 + i) for i in range(3)])[:nb]
    g = np.random.RandomState(1)
    pidx = torch.from_numpy(g.randint(0, P, size=(nb, N)))
    w = torch.from_numpy((g.rand(nb, N) > 0.1).astype(np.float64))
    torch.manual_seed(8)
    _, _, lab = pn2.predict_blocks(net, blocks, batch_size=B, device=DEV, pipeline=False)
    want_pool = O.add_vote(np.zeros((P, NC)), pidx.numpy(), lab.numpy(), w.numpy())
    for pipeline in (False, True):
        torch.manual_seed(8)
        labels, pool = pn2.predict_scene(net, blocks, pidx, w, P, NC, batch_size=B, device=DEV, pipeline=pipeline)
        assert np.array_equal(pool.cpu().numpy(), want_pool.astype(np.int32))
        assert np.array_equal(labels.numpy(), O.vote_argmax(want_pool))
    # two ranks, pools merged by hand (what the all-reduce does): same scene labels.  Each rank's FPS start draws are its
    # own, so compare through rank-wise predict_blocks labels instead of the single-rank run.
    merged, want = None, np.zeros((P, NC))
    for r in range(2):
        torch.manual_seed(40 + r)
        lo, hi, lab_r = pn2.predict_blocks(net, blocks, batch_size=B, rank=r, world=2, device=DEV, pipeline=False)
        want = O.add_vote(want, pidx[lo:hi].numpy(), lab_r.numpy(), w[lo:hi].numpy())
        torch.manual_seed(40 + r)
        _, merged = pn2.predict_scene(net, blocks, pidx, w, P, NC, batch_size=B, rank=r, world=2, device=DEV,
                                      vote_pool=merged, merge=False)
    assert np.array_equal(merged.cpu().numpy(), want.astype(np.int32))
    pn2.set_precision("fp32")


@pytest.mark.parametrize("tag", ["small", "batch"])
def test_rotate_z_matches_reference_fixture(pn2, golden, tag):
    """provider.rotate_point_cloud_z (provider.py:66-84) on the device, in place on the xyz channels of a
    [B, N, 9] batch (point-major and the strided channel-first view), with the angles the reference drew: float64 products
    and sums in np.dot's order, rounded once to float32 -> bit-exact."""
    v = golden("rotation")
    xyz, ang, want = v[tag + "_xyz"], v[tag + "_angles"], v[tag + "_rotated"]
    B, N, _ = xyz.shape
    cs = torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)], axis=1))
    batch = torch.rand(B, N, 9)
    batch[:, :, :3] = torch.from_numpy(xyz)
    dev = batch.to(DEV)
    out = pn2.rotate_point_cloud_z_(dev, cs)
    assert out is dev
    assert np.array_equal(dev[:, :, :3].cpu().numpy(), want)
    assert torch.equal(dev[:, :, 3:].cpu(), batch[:, :, 3:])                  # the other channels are untouched
    assert np.array_equal(O.rotate_z(xyz, ang), want)
    # device-resident angles and a non-contiguous batch view
    dev2 = batch.to(DEV).transpose(1, 2).contiguous().transpose(1, 2)          # [B, N, 9] view of a [B, 9, N] buffer
    pn2.rotate_point_cloud_z_(dev2, cs.to(DEV))
    assert np.array_equal(dev2[:, :, :3].cpu().numpy(), want)
    with pytest.raises(ValueError):
        pn2.rotate_point_cloud_z_(dev, cs[:1])


def test_trainer_rotation_augmentation_uses_reference_draws(pn2):
    """SemSegTrainer(augment_rotate_z=True): the batch the step trains on is the reference's rotated batch (numpy draws in
    cloud order), in every launch mode; labels and the other channels unchanged."""
    B, N, C, NC = 4, 512, 9, 18
    pts, lab = I.facade_batch(B, N, C, 5).to(DEV), I.labels(B, N, NC, 6).to(DEV)
    for mode in ("eager", "graph", "pipeline"):
        pn2.set_precision("bf16")
        torch.manual_seed(1)
        tr = pn2.SemSegTrainer(NC, C - 6, device=DEV, augment_rotate_z=True)
        if mode != "eager":
            tr.enable_cuda_graph(B, N, C, pipeline=mode == "pipeline")
        np.random.seed(9)
        tr.step_device(pts, lab)
        np.random.seed(9)
        ang = np.array([np.random.uniform() * 2 * np.pi for _ in range(B)])
        want = O.rotate_z(pts[:, :, :3].cpu().numpy(), ang)
        if mode == "eager":
            continue                                                        # (the rotated clone is not kept)
        seen = tr._g_points                                                 # the slot of the batch submitted last (both modes)
        if mode == "pipeline":
            torch.cuda.synchronize()
        assert np.array_equal(seen[:, :, :3].cpu().numpy(), want), mode
        assert torch.equal(seen[:, :, 3:], pts[:, :, 3:])
        assert torch.equal(pts[:, :, :3].cpu(), I.facade_batch(B, N, C, 5)[:, :, :3])   # the caller's tensor is not modified
    pn2.set_precision("fp32")
