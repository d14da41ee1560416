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


def test_gradient_sink_registry_checks_object_identity():
    """modules._GRAD_SINK is keyed by id(parameter) but an entry only counts for the parameter object it was made for
    (CPython hands a dead object's id to the next allocation): registry logic, no GPU needed"""
    import gc
    import weakref
    M = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.modules")
    M.clear_grad_sink()
    p, q = torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(3, 4))
    view = torch.zeros(3, 4)
    M.set_grad_sink([p], [view])
    assert M._GRAD_SINK[id(p)][0]() is p and M._GRAD_SINK[id(p)][1] is view
    assert M._sink(p) is None                  # a CPU buffer is never used as a sink (no CPU path)
    dead = torch.nn.Parameter(torch.zeros(3, 4))
    M._GRAD_SINK[id(q)] = (weakref.ref(dead), view)      # what id() re-use leaves behind for q
    del dead
    gc.collect()
    assert M._sink(q) is None
    M.set_grad_sink([], [])                    # any registration prunes entries whose parameter is gone
    assert id(q) not in M._GRAD_SINK and id(p) in M._GRAD_SINK
    M.clear_grad_sink([q])                     # not the owner: no effect
    assert id(p) in M._GRAD_SINK
    M.clear_grad_sink([p])
    assert id(p) not in M._GRAD_SINK


def test_fps_sm_budget_of_the_pipelined_drivers(monkeypatch):
    """trainer._fps_sm_budget: {set-abstraction index: SMs} for the levels that run beside the next batch's FPS kernels"""
    T = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200.trainer")
    monkeypatch.delenv("PN2_SA_SM_BUDGET", raising=False)
    monkeypatch.delenv("PN2_SA_SM_BUDGET_SMS", raising=False)
    assert T._fps_sm_budget(32) == {0: 116, 1: 116}          # one FPS CTA per cloud: 148 - 32
    assert T._fps_sm_budget(8) == {0: 140, 1: 140}
    assert T._fps_sm_budget(128) is None                      # more clouds than half the SMs: the grids are left alone
    assert T._fps_sm_budget(0) is None
    monkeypatch.setenv("PN2_SA_SM_BUDGET", "0")
    assert T._fps_sm_budget(32) is None
    monkeypatch.setenv("PN2_SA_SM_BUDGET", "0,1,2")
    monkeypatch.setenv("PN2_SA_SM_BUDGET_SMS", "120")
    assert T._fps_sm_budget(32) == {0: 120, 1: 120, 2: 120}


def test_sm_budget_context_is_a_no_op_without_a_value():
    """_lib.sm_budget(None / 0) must not touch the library (it is entered on every set-abstraction call of the model)"""
    L = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200._lib")
    with L.sm_budget(None):
        pass
    with L.sm_budget(0):
        pass
