"""GPU parity of the AttentionStudent drop-in (BASELINE config 4) against the CPU oracle
(oracle/student_ref.py, pinned to rtpe.students.AttentionStudent) and the fixture the reference
produced.  Metric per output tensor: max|delta| / max|ref|; fp32 mode <= 1e-4, bf16 mode <= 2e-2.
The decode of the student's outputs (tag_per_joint=False, rtpe/engine.py:45-49) is bit-exact."""
import os

import numpy as np
import pytest
import torch

import rtpe_b200
from rtpe_b200.students import AttentionStudent
from oracle import group_ref as G
from oracle.student_ref import attention_student_forward_ref
from oracle.weights import fill_params_deterministic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(got, ref):
    return ((got.double().cpu() - ref.double()).abs().max() / ref.double().abs().max()).item()


def _student(half, seed):
    net = AttentionStudent(None, "cpu", inplanes=48, num_heatmaps=17, ae_dims=1, half_precision=half)
    fill_params_deterministic(net, seed)
    return net.eval()


def test_student_fp32_vs_oracle(cuda_device):
    net = _student(False, 13)
    x = torch.randn(3, 3, 96, 128, generator=torch.Generator().manual_seed(14))
    att_ref, det_ref = attention_student_forward_ref(net.state_dict(), x)
    net = net.cuda()
    with torch.no_grad():
        att, det = net(x.cuda())
    assert att.shape == att_ref.shape == (3, 1, 24, 32) and det.shape == det_ref.shape == (3, 18, 24, 32)
    assert _rel(att, att_ref) <= 1e-4 and _rel(det, det_ref) <= 1e-4


def test_student_fp32_vs_reference_fixture(cuda_device):
    z = np.load(os.path.join(GOLD, "student_64x96.npz"))
    net = _student(False, int(z["seed"])).cuda()
    with torch.no_grad():
        att, det = net(torch.from_numpy(z["x"]).cuda())
    assert _rel(att, torch.from_numpy(z["att"])) <= 1e-4
    assert _rel(det, torch.from_numpy(z["det"])) <= 1e-4


def test_student_bf16_vs_oracle(cuda_device):
    net = _student(True, 15)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(16))
    att_ref, det_ref = attention_student_forward_ref(sd, x)
    net = net.cuda()
    with torch.no_grad():
        att, det = net(x.cuda())
    assert _rel(att, att_ref) <= 2e-2 and _rel(det, det_ref) <= 2e-2


def test_student_chunks_and_eager_agree(cuda_device):
    """several plan replays (chunk < batch) and the un-graphed plan give the same result"""
    net = _student(False, 17).cuda()
    x = torch.randn(5, 3, 64, 64, generator=torch.Generator().manual_seed(18)).cuda()
    with torch.no_grad():
        net.chunk_size = 8
        a1, d1 = net(x)
        net.chunk_size = 2
        a2, d2 = net(x)
        net.use_cuda_graph = False
        a3, d3 = net(x)
    assert torch.equal(a1, a2) and torch.equal(d1, d2) and torch.equal(a1, a3) and torch.equal(d1, d3)


def test_student_decode_shared_tag_plane(cuda_device):
    """rtpe/engine.py:45-49: heat-maps det[:, :17], one shared tag plane det[:, 17:] (tag_per_joint
    False); the decode of the device maps equals the oracle decode bit for bit."""
    det, tag = rtpe_b200.synth_decode_batch(3, height=64, width=64, max_people=6, seed=41,
                                            tag_per_joint=False)
    kw = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
              use_detection_val=True, ignore_too_much=False, tag_per_joint=False)
    hp = rtpe_b200.HeatmapParser(**kw)
    got = hp.parse_batch(det.cuda(), tag.cuda(), True, True)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), G.DecodeParams(**kw), True, True)
    for (gp, gs), (wp, ws) in zip(got, want):
        assert gp.shape == np.asarray(wp).shape and np.array_equal(gp, wp)
        assert np.array_equal(np.asarray(gs, np.float32), np.asarray(ws, np.float32))


# ---- CamStudent (rtpe/students.py:502-592)
def _cam_student(half, seed):
    from rtpe_b200.students import CamStudent
    net = CamStudent(None, "cpu", inplanes=48, num_stages=3, num_heatmaps=17, ae_dims=1,
                     half_precision=half)
    fill_params_deterministic(net, seed)
    return net.eval()


def test_cam_student_fp32_vs_oracle_and_fixture(cuda_device):
    from oracle.student_ref import cam_student_forward_ref
    net = _cam_student(False, 21)
    x = torch.randn(3, 3, 96, 128, generator=torch.Generator().manual_seed(22))
    (ref,) = cam_student_forward_ref(net.state_dict(), x)
    (ref_up,) = cam_student_forward_ref(net.state_dict(), x, out_hw=(50, 70))
    net = net.cuda()
    with torch.no_grad():
        (pred,) = net(x.cuda())
        (pred_up,) = net(x.cuda(), out_hw=(50, 70))
    assert pred.shape == ref.shape == (3, 18, 24, 32) and pred_up.shape == ref_up.shape
    assert _rel(pred, ref) <= 1e-4 and _rel(pred_up, ref_up) <= 1e-4
    z = np.load(os.path.join(GOLD, "cam_student_64x96.npz"))
    net = _cam_student(False, int(z["seed"])).cuda()
    with torch.no_grad():
        (pred,) = net(torch.from_numpy(z["x"]).cuda())
        (pred_up,) = net(torch.from_numpy(z["x"]).cuda(), out_hw=tuple(z["pred_up"].shape[2:]))
    assert _rel(pred, torch.from_numpy(z["pred"])) <= 1e-4
    assert _rel(pred_up, torch.from_numpy(z["pred_up"])) <= 1e-4
    with pytest.raises(NotImplementedError):
        net(torch.from_numpy(z["x"]).cuda(), return_intermediate=True)


def test_cam_student_bf16_vs_oracle(cuda_device):
    from oracle.student_ref import cam_student_forward_ref
    net = _cam_student(True, 23)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(24))
    (ref,) = cam_student_forward_ref(sd, x)
    net = net.cuda()
    with torch.no_grad():
        (pred,) = net(x.cuda())
    assert _rel(pred, ref) <= 2e-2


# ---- RefinerStudent (rtpe/students.py:302-386)
def _refiner_student(half, seed):
    from rtpe_b200.students import RefinerStudent
    net = RefinerStudent(None, "cpu", half_precision=half)
    fill_params_deterministic(net, seed)
    return net.eval()


def test_refiner_student_fp32_vs_oracle_and_fixture(cuda_device):
    from oracle.student_ref import refiner_student_forward_ref
    net = _refiner_student(False, 31)
    x = torch.randn(3, 3, 64, 96, generator=torch.Generator().manual_seed(32))
    ref = refiner_student_forward_ref(net.state_dict(), x)
    ref_up = refiner_student_forward_ref(net.state_dict(), x, out_hw=(50, 70))
    net = net.cuda()
    with torch.no_grad():
        pred = net(x.cuda())
        pred_up = net(x.cuda(), out_hw=(50, 70))
    assert pred.shape == ref.shape == (3, 18, 16, 24) and pred_up.shape == ref_up.shape
    assert _rel(pred, ref) <= 1e-4 and _rel(pred_up, ref_up) <= 1e-4
    z = np.load(os.path.join(GOLD, "refiner_student_64x96.npz"))
    net = _refiner_student(False, int(z["seed"])).cuda()
    with torch.no_grad():
        pred = net(torch.from_numpy(z["x"]).cuda())
        pred_up = net(torch.from_numpy(z["x"]).cuda(), out_hw=tuple(z["pred_up"].shape[2:]))
    assert _rel(pred, torch.from_numpy(z["pred"])) <= 1e-4
    assert _rel(pred_up, torch.from_numpy(z["pred_up"])) <= 1e-4


def test_refiner_student_bf16_vs_oracle(cuda_device):
    from oracle.student_ref import refiner_student_forward_ref
    net = _refiner_student(True, 33)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(34))
    ref = refiner_student_forward_ref(sd, x)
    net = net.cuda()
    with torch.no_grad():
        pred = net(x.cuda())
    assert _rel(pred, ref) <= 2e-2
