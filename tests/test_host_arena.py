"""Host-side logic that needs no GPU: the operator-output arena behind the two-slot pipelined graphs
(ops.record_outputs / ops.reuse_outputs) and the loud failures of the CUDA-only drivers on CPU tensors."""
import importlib

import pytest
import torch

pn2 = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
ops = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.ops")


def test_record_then_reuse_hands_out_the_same_tensors_in_order():
    cpu = torch.device("cpu")
    with ops.record_outputs() as rec:
        a = ops._out((2, 3), torch.int64, cpu)
        b = ops._out((2, 3, 3), torch.float32, cpu)
    assert [t.data_ptr() for t in rec.tensors] == [a.data_ptr(), b.data_ptr()]
    fresh = ops._out((2, 3), torch.int64, cpu)                      # outside a block: ordinary allocation
    assert fresh.data_ptr() not in (a.data_ptr(), b.data_ptr())
    with ops.reuse_outputs(rec.tensors):
        a2 = ops._out((2, 3), torch.int64, cpu)
        b2 = ops._out((2, 3, 3), torch.float32, cpu)
    assert a2 is a and b2 is b


def test_reuse_outputs_rejects_a_different_operator_sequence():
    cpu = torch.device("cpu")
    with ops.record_outputs() as rec:
        ops._out((4,), torch.int64, cpu)
        ops._out((4, 3), torch.float32, cpu)
    with pytest.raises(RuntimeError, match="recorded"):             # wrong shape
        with ops.reuse_outputs(rec.tensors):
            ops._out((5,), torch.int64, cpu)
    with pytest.raises(RuntimeError, match="recorded"):             # wrong dtype
        with ops.reuse_outputs(rec.tensors):
            ops._out((4,), torch.int32, cpu)
    with pytest.raises(RuntimeError, match="1 of 2"):               # fewer outputs than recorded
        with ops.reuse_outputs(rec.tensors):
            ops._out((4,), torch.int64, cpu)
    with pytest.raises(RuntimeError, match="more operator outputs"):
        with ops.reuse_outputs(rec.tensors):
            ops._out((4,), torch.int64, cpu)
            ops._out((4, 3), torch.float32, cpu)
            ops._out((4, 3), torch.float32, cpu)
    assert ops._ARENA is None                                        # every block restored the outer state


def test_flat_adam_has_no_cpu_path():
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    grads = pn2.FlatGradients(params)
    assert grads.offsets == [0, 12] and grads.flat.numel() == 20     # 16-byte aligned slices
    with pytest.raises(ValueError, match="CUDA"):
        pn2.FlatAdam(grads)


def test_training_chains_lists_every_mlp_once():
    net = pn2.get_model(18, 3)
    chains = net.training_chains()
    assert len(chains) == 8
    convs = [c for chain in chains for c in chain]
    assert len(convs) == len(set(map(id, convs))) == 4 * 3 + 2 + 2 + 2 + 3      # conv1 joins fp1's chain only with the fused head (CUDA, bf16)
