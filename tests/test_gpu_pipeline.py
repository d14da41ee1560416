"""Software-pipelined drivers (trainer.SemSegTrainer / SemSegPredictor with pipeline=True): the index pipeline of
batch i+1 runs beside the feature path of batch i inside one CUDA graph.  Per batch the arithmetic is the same as in
the un-pipelined graph, so with the same seeds the two must agree: labels exactly (the inference kernels have no
atomics), training losses to 1e-3 relative over several optimizer steps (fp32 scatter-add order differs run to run).
The body mirrors /root/reference/localfunctions.py:202-218 (train batch) and :396-400 (test batch)."""
import pytest
import torch

import _inputs as I

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, N, C, NC = 4, 1024, 9, 18


def _batches(n, seed):
    return [(I.facade_batch(B, N, C, seed + i).to(DEV), I.labels(B, N, NC, seed + 50 + i).to(DEV)) for i in range(n)]


def _trainer(pn2, pipeline, lr=0.0):
    pn2.set_precision("bf16")
    torch.manual_seed(7)
    # lr 0 (and no weight decay): the parameters stay put, so every batch's loss depends on that batch, its indices and the
    # kernels only -- two launch modes must agree to rounding.  (With lr > 0 Adam's first updates are lr * sign(g): the
    # run-to-run order of the fp32 atomic additions flips near-zero gradients and two runs of the SAME code drift apart by
    # 1e-3 within five steps, which says nothing about the pipeline.)
    t = pn2.SemSegTrainer(NC, C - 6, lr=lr, weight_decay=0.0 if lr == 0.0 else 1e-4, device=DEV)
    t.model.drop1.p = 0.0                       # the dropout mask comes from the CUDA generator: keep the two runs comparable
    t.enable_cuda_graph(B, N, C, pipeline=pipeline)
    return t


@pytest.mark.parametrize("sm_budget,tol", [("0", 1e-5), (None, 1e-3)])
def test_pipelined_train_steps_match_unpipelined(pn2, monkeypatch, sm_budget, tol):
    """sm_budget "0": the pipelined graph launches exactly the un-pipelined graph's kernels (same grids), so the losses agree
    to rounding.  Default: the pipelined graph sizes the grids of sa1 / sa2's forward layers for the SMs FPS leaves free
    (trainer._fps_sm_budget); fewer CTAs -> other fp32 partial sums of the BatchNorm statistics -> ReLU decisions of
    near-zero activations flip: the same 1e-4-level noise two runs with different atomic order show."""
    if sm_budget is not None:
        monkeypatch.setenv("PN2_SA_SM_BUDGET", sm_budget)
    data = _batches(5, 300)
    plain = _trainer(pn2, False)
    torch.manual_seed(99)                       # FPS start draws: 4 per batch, in batch order, in both modes
    want = [float(plain.step_device(p, t)) for p, t in data]
    piped = _trainer(pn2, True)
    torch.manual_seed(99)
    got = []
    for p, t in data:
        loss = piped.step_device(p, t)
        if loss is not None:
            got.append(float(loss))
    assert len(got) == len(data) - 1            # one batch is still in flight
    rest = piped.flush()
    assert len(rest) == 1 and piped.flush() == []
    got += rest
    for a, b in zip(got, want):
        assert abs(a - b) <= tol * abs(b), (got, want)      # same batch, same indices, same parameters
    # running statistics after the same five batches (a skipped or repeated batch would show here)
    for (k, a), b in zip(piped.model.state_dict().items(), plain.model.state_dict().values()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-4 if tol < 1e-4 else 1e-2, atol=1e-5 if tol < 1e-4 else 2e-3), k
        else:
            assert torch.equal(a, b), k          # num_batches_tracked
    pn2.set_precision("fp32")


def test_pipelined_steps_from_host_match_device_steps(pn2):
    """step() stages the host batch over a copy stream and reads losses back one call late (two more pipeline stages: the
    loss of the batch handed in three calls earlier comes back); the losses are those of step_device() on the same batches."""
    data = _batches(5, 400)
    dev_tr = _trainer(pn2, True)
    torch.manual_seed(77)
    want = []
    for p, t in data:
        loss = dev_tr.step_device(p, t)                  # a static tensor the next replay rewrites: read it now
        if loss is not None:
            want.append(float(loss))
    want += dev_tr.flush()
    host_tr = _trainer(pn2, True)
    torch.manual_seed(77)
    got = []
    for i, (p, t) in enumerate(data):
        loss = host_tr.step(p.cpu().pin_memory(), t.cpu().pin_memory())
        assert (loss is None) == (i < 3)
        if loss is not None:
            assert isinstance(loss, float)
            got.append(loss)
    got += host_tr.flush()
    assert len(got) == len(want) == len(data) and host_tr.flush() == []
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-5 * abs(b), (got, want)
    pn2.set_precision("fp32")


def test_pipelined_training_updates_parameters(pn2):
    """With a real learning rate the pipelined graph does train: parameters move every step and the loss of a repeated
    batch goes down."""
    piped = _trainer(pn2, True, lr=1e-3)
    p, t = _batches(1, 900)[0]
    before = [q.detach().clone() for q in piped.model.parameters()]
    losses = []
    for _ in range(12):
        loss = piped.step_device(p, t)
        if loss is not None:
            losses.append(float(loss))
    losses += piped.flush()
    assert len(losses) == 12 and all(l == l and l < 10.0 for l in losses)
    moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, piped.model.parameters()))
    assert moved >= len(before) - 8              # (conv biases under train-mode BatchNorm receive exactly zero gradient... and decay)
    assert min(losses[-3:]) < losses[0]
    pn2.set_precision("fp32")


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_pipelined_predictor_matches_unpipelined(pn2, precision):
    pn2.set_precision(precision)
    torch.manual_seed(3)
    net = I.randomize_module_(pn2.get_model(NC, C - 6), 17).to(DEV).eval()
    data = [I.facade_batch(B, N, C, 500 + i) for i in range(4)]
    data.append(I.facade_batch(B - 1, N, C, 600))                  # a short tail batch
    plain = pn2.SemSegPredictor(net, B, N, C, DEV)
    torch.manual_seed(5)
    want = [plain.predict_host(x).clone() for x in data]
    piped = pn2.SemSegPredictor(net, B, N, C, DEV, pipeline=True)
    torch.manual_seed(5)
    for on_device in (False, True):                                # host input: one more stage (the copy stream)
        torch.manual_seed(5)
        got = []
        for i, x in enumerate(data):
            out = piped.submit(x.to(DEV) if on_device else x.pin_memory())
            assert (out is None) == (i < (2 if on_device else 3))
            if out is not None:
                got.append(out.clone())
        got += piped.flush()
        assert piped.flush() == []
        assert [g.shape for g in got] == [w.shape for w in want]
        for g, w in zip(got, want):
            assert torch.equal(g, w)
    with pytest.raises(RuntimeError):
        piped.predict_host(data[0])
    pn2.set_precision("fp32")


def test_predict_blocks_pipelined_equals_plain(pn2):
    pn2.set_precision("bf16")
    torch.manual_seed(3)
    net = I.randomize_module_(pn2.get_model(NC, C - 6), 19).to(DEV).eval()
    blocks = torch.cat([I.facade_batch(B, N, C, 700 + i) for i in range(3)])[:11]
    torch.manual_seed(8)
    lo, hi, want = pn2.predict_blocks(net, blocks, batch_size=B, device=DEV, pipeline=False)
    torch.manual_seed(8)
    lo2, hi2, got = pn2.predict_blocks(net, blocks, batch_size=B, device=DEV, pipeline=True)
    assert (lo, hi) == (lo2, hi2) == (0, 11)
    assert torch.equal(got, want)
    pn2.set_precision("fp32")


def test_predictor_cache_does_not_leak_into_later_forwards(pn2):
    """SemSegPredictor computes folded BatchNorm / packed weights once (modules.frozen_parameters): its labels equal the eager
    forward's at construction time; after a parameter change the eager forward (no cache outside the context) follows
    immediately, and so does a NEW predictor (the old one must be rebuilt: part of what it reads is frozen, part is live)."""
    pn2.set_precision("bf16")
    torch.manual_seed(3)
    net = I.randomize_module_(pn2.get_model(NC, C - 6), 31).to(DEV).eval()
    x = I.facade_batch(B, N, C, 800)
    start_seed = 21

    def eager():
        torch.manual_seed(start_seed)
        with torch.no_grad():
            return net(x.to(DEV).transpose(2, 1))[0].argmax(2).cpu()

    pred = pn2.SemSegPredictor(net, B, N, C, DEV)
    torch.manual_seed(start_seed)
    first = pred.predict_host(x).clone()
    assert torch.equal(first, eager())
    with torch.no_grad():                                   # change the parameters the way training would
        net.conv2.bias.add_(torch.linspace(-3, 3, NC, device=DEV))
        net.fp1.mlp_bns[0].running_mean.mul_(-1.0)
        net.sa1.mlp_convs[0].weight.mul_(1.5)
    changed = eager()
    assert not torch.equal(changed, first)                  # eager: recomputed from the live parameters
    pred2 = pn2.SemSegPredictor(net, B, N, C, DEV)
    torch.manual_seed(start_seed)
    assert torch.equal(pred2.predict_host(x), changed)      # a new one sees the new parameters
    pn2.set_precision("fp32")


def test_recapture_keeps_optimizer_state_and_refuses_pending_batches(pn2):
    """enable_cuda_graph's warm-up steps must neither count as training nor reset Adam (a loaded checkpoint's
    moments / step counter, or those of the epochs trained so far when the caller re-captures)."""
    data = _batches(4, 500)
    t = _trainer(pn2, True, lr=1e-3)
    torch.manual_seed(5)
    for p, y in data:
        t.step_device(p, y)
    with pytest.raises(RuntimeError, match="flush"):
        t.enable_cuda_graph(B, N, C, pipeline=True)          # a batch is still in flight
    t.flush()
    torch.cuda.synchronize()
    opt = t.optimizer
    before = (opt.exp_avg.clone(), opt.exp_avg_sq.clone(), float(opt.step_count))
    params = [p.detach().clone() for p in t.model.parameters()]
    assert before[2] == 4.0 and float(before[0].abs().sum()) > 0
    t.enable_cuda_graph(B, N, C, pipeline=True)
    torch.cuda.synchronize()
    assert float(opt.step_count) == 4.0
    assert torch.equal(opt.exp_avg, before[0]) and torch.equal(opt.exp_avg_sq, before[1])
    for p, q in zip(t.model.parameters(), params):
        assert torch.equal(p, q)
    # a torch.optim.Adam checkpoint survives the capture too
    ref_opt = torch.optim.Adam(t.model.parameters(), lr=1e-3)
    for p in t.model.parameters():
        p.grad = torch.ones_like(p)
    ref_opt.step()
    sd = ref_opt.state_dict()
    t.optimizer.load_state_dict(sd)
    t.enable_cuda_graph(B, N, C)
    torch.cuda.synchronize()
    assert float(opt.step_count) == 1.0 and float(opt.exp_avg.abs().sum()) > 0
    pn2.set_precision("fp32")


def test_captured_step_follows_the_momentum_schedule_without_recapture(pn2):
    """localfunctions.py:191-195 sets m.momentum on every BatchNorm each epoch; a captured bf16 step reads it from device
    memory: after the change the replayed step must update the running statistics like an eager step with that momentum."""
    (p, y), = _batches(1, 600)

    def run(captured):
        t = _trainer(pn2, False)
        if not captured:
            t._graph = None
        for m in t.model.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.momentum = 0.025                           # after the capture (which baked nothing in)
        torch.manual_seed(3)
        t.step_device(p, y)
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in t.model.state_dict().items() if "running" in k}

    eager, replay = run(False), run(True)
    for k in eager:
        assert torch.allclose(eager[k], replay[k], rtol=1e-5, atol=1e-6), k
    # and the value is the scheduled one: running_mean = 0.975 * 0 + 0.025 * batch mean, so it stays small
    t0 = _trainer(pn2, False)
    torch.manual_seed(3)
    t0.step_device(p, y)
    torch.cuda.synchronize()
    k = "sa1.mlp_bns.0.running_mean"
    ratio = float(replay[k].abs().sum() / t0.model.state_dict()[k].abs().sum())
    assert abs(ratio - 0.25) < 1e-3, ratio                   # 0.025 / 0.1
    pn2.set_precision("fp32")
