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


# ---- MultistageStudent (rtpe/students.py:389-499)
def _multistage_student(half, seed, **kw):
    from rtpe_b200.students import MultistageStudent
    net = MultistageStudent(None, "cpu", half_precision=half, **kw)
    fill_params_deterministic(net, seed)
    return net.eval()


def test_multistage_student_fp32_vs_oracle_and_fixture(cuda_device):
    from oracle.student_ref import multistage_student_forward_ref
    net = _multistage_student(False, 41, layers_per_stage=[2, 3, 1, 2])
    x = torch.randn(3, 3, 64, 96, generator=torch.Generator().manual_seed(42))
    ref = multistage_student_forward_ref(net.state_dict(), x)
    ref_up = multistage_student_forward_ref(net.state_dict(), x, out_hw=(50, 70))
    net = net.cuda()
    with torch.no_grad():
        outs = net(x.cuda())
        outs_up = net(x.cuda(), out_hw=(50, 70))
        again = net(x.cuda())                                   # the out_hw plan did not replace it
    assert len(outs) == len(ref) == 4 and outs[0].shape == ref[0].shape == (3, 18, 16, 24)
    assert outs_up[0].shape == ref_up[0].shape == (3, 18, 50, 70)
    for g, r in zip(outs + outs_up, ref + ref_up):
        assert _rel(g, r) <= 1e-4
    assert all(torch.equal(a, b) for a, b in zip(outs, again))
    z = np.load(os.path.join(GOLD, "multistage_student_64x96.npz"))
    net = _multistage_student(False, int(z["seed"])).cuda()
    with torch.no_grad():
        outs = net(torch.from_numpy(z["x"]).cuda())
        outs_up = net(torch.from_numpy(z["x"]).cuda(), out_hw=tuple(z["outs_up"].shape[3:]))
    for g, r in zip(outs, z["outs"]):
        assert _rel(g, torch.from_numpy(r)) <= 1e-4
    for g, r in zip(outs_up, z["outs_up"]):
        assert _rel(g, torch.from_numpy(r)) <= 1e-4


def test_multistage_student_bf16_vs_oracle(cuda_device):
    from oracle.student_ref import multistage_student_forward_ref
    net = _multistage_student(True, 43)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(44))
    ref = multistage_student_forward_ref(sd, x)
    ref_up = multistage_student_forward_ref(sd, x, out_hw=(40, 56))
    net = net.cuda()
    with torch.no_grad():
        outs = net(x.cuda())
        outs_up = net(x.cuda(), out_hw=(40, 56))
    for g, r in zip(outs + outs_up, ref + ref_up):
        assert torch.isfinite(g).all()
        assert _rel(g, r) <= 2e-2


# ---- AttentionStudentSteps (rtpe/students.py:786-1073, eval_attention.py:95-104)
def _steps_student(half, seed, inplanes, ae_dims=1):
    from rtpe_b200.students import AttentionStudentSteps
    net = AttentionStudentSteps(None, "cpu", inplanes, 17, ae_dims, half)
    fill_params_deterministic(net, seed)
    return net.eval()


def test_attention_student_steps_fp32_vs_oracle_and_fixture(cuda_device):
    from oracle.student_ref import attention_student_steps_forward_ref
    net = _steps_student(False, 51, 80, ae_dims=0)              # eval_attention.py's configuration
    g = torch.Generator().manual_seed(52)
    x = torch.randn(3, 3, 64, 96, generator=g)
    alt = torch.randn(3, 3, 64, 96, generator=g)
    ref = attention_student_steps_forward_ref(net.state_dict(), x, alt)
    ref20 = attention_student_steps_forward_ref(net.state_dict(), x, alt, att_divisor=20)
    net = net.cuda()
    net.chunk_size = 2                                           # 3 images = chunks of 2 + 1
    with torch.no_grad():
        got = net(x.cuda(), alt=alt.cuda())
        got20 = net(x.cuda(), alt=alt.cuda(), att_divisor=20)
    assert got[0].shape == ref[0].shape == (3, 1, 16, 24) and got[1].shape == ref[1].shape == (3, 17, 16, 24)
    for a, b in zip(got + got20, ref + ref20):
        assert _rel(a, b) <= 1e-4
    z = np.load(os.path.join(GOLD, "attention_steps_64x96.npz"))
    net = _steps_student(False, int(z["seed"]), int(z["inplanes"])).cuda()
    x, alt = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["alt"]).cuda()
    with torch.no_grad():
        att, det = net(x, alt=alt)
        att20, det20 = net(x, alt=alt, att_divisor=20)
    for a, key in ((att, "att"), (det, "det"), (att20, "att20"), (det20, "det20")):
        assert _rel(a, torch.from_numpy(z[key])) <= 1e-4
    with pytest.raises(NotImplementedError):
        net(x)


@pytest.mark.parametrize("inplanes", [48, 80])
def test_attention_student_steps_bf16_vs_oracle(cuda_device, inplanes):
    from oracle.student_ref import attention_student_steps_forward_ref
    net = _steps_student(True, 53, inplanes)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(54)
    x = torch.randn(2, 3, 128, 128, generator=g)
    alt = torch.randn(2, 3, 128, 128, generator=g)
    ref = attention_student_steps_forward_ref(sd, x, alt, att_divisor=20)
    net = net.cuda()
    with torch.no_grad():
        got = net(x.cuda(), alt=alt.cuda(), att_divisor=20)
    for a, b in zip(got, ref):
        assert torch.isfinite(a).all()
        assert _rel(a, b) <= 2e-2


# ---- eval_student (rtpe/engine.py:21-75)
class _FakeValSet:
    def evaluate(self, preds, scores, out_dir, a, b):
        self.got = (preds, scores, out_dir, a, b)
        return {"AP": 0.5, "people": sum(len(p) for p in preds)}, 0.5


class _FakeLoader:
    def __init__(self, batches):
        self.dataset = _FakeValSet()
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)


class _BlobModel(torch.nn.Module):
    """Stand-in student: returns the decode generator's maps, resized like a student's out_hw."""
    def __init__(self, det, tag):
        super().__init__()
        self.pred = torch.cat((det, tag[..., 0]), dim=1)          # (N, 17 + 1, h, w)
        self.at = 0

    def forward(self, img, out_hw=None):
        n = img.shape[0]
        out = self.pred[self.at:self.at + n].to(img.device)
        self.at += n
        return out


def test_eval_student_batched_equals_reference_loop(cuda_device):
    """Batched decode of a whole loader batch == the reference's image-by-image ``parse`` loop
    (rtpe/engine.py:38-51), and the evaluate() call gets the reference's arguments."""
    det, tag = rtpe_b200.synth_decode_batch(5, height=64, width=80, max_people=5, seed=61,
                                            tag_per_joint=False)
    kw = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
              use_detection_val=True, ignore_too_much=False, tag_per_joint=False)
    hp = rtpe_b200.HeatmapParser(**kw)
    imgs = torch.zeros(5, 3, 64, 80)
    mk = lambda sizes: _FakeLoader([(torch.arange(a, b), imgs[a:b], None, None, None, None)  # noqa: E731
                                    for a, b in sizes])
    one = mk([(i, i + 1) for i in range(5)])
    many = mk([(0, 2), (2, 5)])
    r1 = rtpe_b200.eval_student(_BlobModel(det, tag), hp, one, "cuda", batched=False, verbose=False)
    r2 = rtpe_b200.eval_student(_BlobModel(det, tag), hp, many, "cuda", verbose=False)
    assert r1 == r2 and r1["people"] > 0
    assert one.dataset.got[2:] == (".", False, False)
    want = G.parse_batch_ref(det.numpy().copy(), tag.numpy().copy(), G.DecodeParams(**kw), True, True)
    for got in (one.dataset.got, many.dataset.got):
        preds, scores = got[:2]
        assert len(preds) == len(scores) == 5
        for gp, gs, (wp, ws) in zip(preds, scores, want):
            wp = [x for x in np.asarray(wp) if x.size > 0]
            assert len(gp) == len(wp) and all(np.array_equal(a, b) for a, b in zip(gp, wp))
            assert np.array_equal(np.asarray(gs, np.float32), np.asarray(ws, np.float32))
    with pytest.raises(NotImplementedError):
        rtpe_b200.eval_student(_BlobModel(det, tag), hp, one, "cuda", plot_every=1)
