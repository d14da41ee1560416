"""Host-side multi-rank logic on CPU: block sharding and the flat-gradient all-reduce of
data-parallel training, run as TWO real processes over the gloo backend,
plus bench.py's reference arm under a 2-rank launch (rank 0 prints, the other exits 0).
No CUDA kernels are involved: the tensors here only exercise trainer.shard_range /
trainer.FlatGradients and the launcher contract."""
import json
import os
import socket
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import importlib, os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)

# --- block sharding: contiguous, disjoint, covering (trainer.shard_range) ---------------------
n_blocks = 37
lo, hi = pn2.shard_range(n_blocks, rank, world)
owned = torch.zeros(n_blocks, dtype=torch.int64)
owned[lo:hi] = 1
dist.all_reduce(owned)
assert bool((owned == 1).all()), owned.tolist()

# --- flat-gradient all-reduce: every rank ends with the MEAN of the rank gradients --------------
torch.manual_seed(0)                      # identical parameters on both ranks
net = torch.nn.Sequential(torch.nn.Conv1d(6, 8, 1), torch.nn.BatchNorm1d(8), torch.nn.Conv1d(8, 3, 1))
flat = pn2.FlatGradients(net.parameters())
assert flat.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in net.parameters())      # slices start on 16-byte boundaries
torch.manual_seed(100 + rank)             # rank-local data
x = torch.randn(4, 6, 16)
flat.zero()
net(x).square().mean().backward()
flat.adopt()
for p, v in zip(flat.params, flat.views):
    assert p.grad.data_ptr() == v.data_ptr()
local = flat.flat.clone()
gathered = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(gathered, local)
flat.all_reduce_mean()
want = torch.stack(gathered).mean(0)
assert torch.allclose(flat.flat, want, rtol=1e-6, atol=1e-7)
assert not torch.equal(gathered[0], gathered[1])          # the ranks really saw different data
# a second step re-zeroes the buffer and gradients stay views of it
flat.zero()
assert float(flat.flat.abs().sum()) == 0.0
net(x).square().mean().backward()
flat.adopt()
assert torch.allclose(flat.flat, local, rtol=1e-6, atol=1e-7)
dist.barrier()
dist.destroy_process_group()
print(json.dumps({"rank": rank, "lo": lo, "hi": hi, "ok": True}))
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _launch(n, argv, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE=str(n),
               OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable] + argv, env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), cwd=ROOT,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(n)]
    outs = [p.communicate(timeout=timeout) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    return [so for so, _ in outs]


def test_shard_range_partitions():
    import importlib
    pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
    for n in (0, 1, 7, 8, 9, 8501):
        for world in (1, 2, 3, 8):
            spans = [pn2.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert a <= b == c <= d
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= max(1, (n + world - 1) // world)


def test_two_rank_gloo_sharding_and_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    outs = _launch(2, [str(script)])
    got = sorted((json.loads(o.strip().splitlines()[-1]) for o in outs), key=lambda d: d["rank"])
    assert [g["ok"] for g in got] == [True, True]
    assert (got[0]["lo"], got[0]["hi"], got[1]["lo"], got[1]["hi"]) == (0, 19, 19, 37)


def test_reference_arm_two_rank_launch_prints_once():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 alone runs and prints ONE JSON line,
    the other rank exits 0 without work (a tiny sample so the test stays short)."""
    outs = _launch(2, ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                       "--ref-sample-clouds", "2"], timeout=900)
    assert outs[1].strip() == ""
    line = json.loads(outs[0].strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "points/s"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
