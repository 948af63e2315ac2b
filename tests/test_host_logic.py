"""CPU: host-side logic of the network drop-in -- parameter count, plan recording (layer
list, FLOP count, arena packing) without launching anything."""
import ctypes as C

import pytest
import torch

import rtpe_b200
from rtpe_b200 import _lib as L


def test_parameter_count_matches_reference_comment():
    net = rtpe_b200.PoseHigherResolutionNet()
    assert sum(p.numel() for p in net.parameters()) == 63827139   # rtpe/students.py:208
    assert len(net.state_dict()) == 1810


def test_plan_recording_layer_list_and_flops():
    """Recording on the CPU 'device' exercises BN folding, tap tables, the deconv phase split,
    concat-buffer wiring and the arena without any kernel launch."""
    lib = L.load(require_cuda=False)
    net = rtpe_b200.PoseHigherResolutionNet().eval()
    net.conv_engine = L.ENGINE_FFMA        # the tcgen05 engines need a CUDA driver (tensor maps)
    dev = torch.device("cpu")
    R, outs = net._record(1, 128, 128, "fp32", dev, False, False)
    in_buf = torch.zeros(1, 3, 128, 128)
    plan = R.build(in_buf)
    try:
        kinds = [op[0] for op in R.ops]
        # 302 Conv2d - conv1 (stem kernel) + 4 deconv phases = 305 conv launches
        assert kinds.count("conv") == 301 + 4
        assert kinds.count("stem") == 1
        assert kinds.count("nchw") == 2
        assert kinds.count("fuse") == 2 + 4 * 3 + 2 * 4 + 1
        assert lib.brtpe_plan_num_ops(plan) == len(R.ops)
        flops = lib.brtpe_plan_conv_flops(plan)
        # 149.469 GMAC at 640x640 (conv1's 0.177 GMAC runs in the stem
        # kernel; the 82->96 channel padding of the deconv input adds 0.275 GMAC of zeros)
        expect = (149.469 - 0.177) * 2e9 * (128 / 640) ** 2
        assert abs(flops - expect) / expect < 0.01
        assert [tuple(o.shape) for o in outs] == [(1, 34, 32, 32), (1, 17, 64, 64)]
        # liveness packing: far fewer bytes than the sum of all activations
        total = sum(t.numel() * 4 for t in R.tensors)
        assert R.arena_bytes < 0.2 * total
    finally:
        lib.brtpe_plan_destroy(plan)


def test_invalid_configurations_raise():
    import pytest
    with pytest.raises(ValueError):
        rtpe_b200.HighResolutionModule(2, rtpe_b200.BasicBlock, [4], [48, 96], [48, 96], "SUM")
    with pytest.raises(KeyError):
        rtpe_b200.PoseHigherResolutionNet(s2_block_type="NOPE")
    with pytest.raises(ValueError):
        rtpe_b200.PoseHigherResolutionNet(final_conv_ksize=5)
    with pytest.raises(ValueError):
        rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, munkres_start_rule="nope")


def test_eval_student_prediction_selection():
    """rtpe/engine.py:41 expects one tensor from ``model(img, out_hw)``; the drop-in students return a
    tensor (RefinerStudent), a list of stage outputs (CamStudent / MultistageStudent: last stage) or
    ``(att, det)`` (attention students: det)."""
    import torch
    from rtpe_b200.engine import _student_prediction
    a, b = torch.zeros(1, 18, 4, 4), torch.ones(1, 18, 4, 4)
    img = torch.zeros(1, 3, 16, 16)
    assert _student_prediction(lambda x, hw: b, img, (16, 16)) is b
    assert _student_prediction(lambda x, hw: [a, b], img, (16, 16)) is b
    assert _student_prediction(lambda x, hw: (a, b), img, (16, 16)) is b
    with pytest.raises(NotImplementedError):
        from rtpe_b200 import eval_student
        eval_student(torch.nn.Identity(), None, [], "cpu", plot_every=1)
