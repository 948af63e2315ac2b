"""CPU, this container only: the oracle restatements against the UNMODIFIED reference
imported from /root/reference (skipped where the tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

import rtpe_b200
from oracle import group_ref as G
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


@pytest.mark.parametrize("seed,h,w,t,people,tpj", [(1, 64, 80, 1, 8, True), (2, 72, 64, 2, 12, True),
                                                   (3, 48, 48, 1, 30, True), (4, 56, 40, 1, 5, False)])
def test_decode_oracle_vs_reference(ref, seed, h, w, t, people, tpj):
    ref_group, _ = ref
    det, tag = rtpe_b200.synth_decode_batch(2, height=h, width=w, tag_dims=t, max_people=people,
                                            seed=seed, tag_per_joint=tpj)
    hp = ref_group.HeatmapParser(17, 30, 0.1, 1.0, True, False, tpj, 5, 2)
    p = G.DecodeParams(tag_per_joint=tpj)
    for i in range(2):
        ans, scores = hp.parse(det[i:i + 1].clone(), tag[i:i + 1].clone(), True, True)
        gp, gs = G.parse_image_ref(det[i:i + 1].numpy().copy(), tag[i:i + 1].numpy().copy(), p)
        assert np.array_equal(np.asarray(ans[0]), gp)
        assert np.array_equal(np.asarray(scores, np.float32), np.asarray(gs, np.float32))


def test_decode_oracle_vs_reference_collisions(ref):
    ref_group, _ = ref
    det, tag = rtpe_b200.synth_decode_batch(2, height=64, width=64, max_people=30, seed=55)
    tag = tag.to(torch.bfloat16).float() * 0.25
    hp = ref_group.HeatmapParser(17, 30, 0.1, 1.0, True, False, True, 5, 2)
    p = G.DecodeParams()
    for i in range(2):
        ans, scores = hp.parse(det[i:i + 1].clone(), tag[i:i + 1].clone(), True, True)
        gp, gs = G.parse_image_ref(det[i:i + 1].numpy().copy(), tag[i:i + 1].numpy().copy(), p)
        assert np.array_equal(np.asarray(ans[0]), gp)


def test_model_oracle_vs_reference(ref):
    from oracle.hhrnet_ref import hhrnet_forward_ref
    _, ref_model = ref
    torch.manual_seed(0)
    net = ref_model.PoseHigherResolutionNet().eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        want = net(x)
        got = hhrnet_forward_ref(net.state_dict(), x)
    for a, b in zip(got, want):
        assert ((a - b).abs().max() / b.abs().max()).item() <= 1e-5


@pytest.mark.parametrize("name", ["two_deconvs_k3_heads", "bottleneck_stages_k3_k2_deconvs",
                                  "no_deconv_shared_tag"])
def test_model_variants_oracle_and_dropin_vs_reference(ref, name):
    """constructor options the W48 teacher does not use (BOTTLENECK stages, 3x3 heads, several deconv
    stages with kernel sizes 4 / 3 / 2): the drop-in has the reference's state-dict layout and the
    oracle reproduces the reference's outputs."""
    from oracle.hhrnet_ref import hhrnet_forward_ref
    from oracle.variants import variant_kwargs
    from oracle.weights import fill_params_deterministic
    _, ref_model = ref
    net = ref_model.PoseHigherResolutionNet(**variant_kwargs(name)).eval()
    fill_params_deterministic(net, 31)
    mine = rtpe_b200.PoseHigherResolutionNet(**variant_kwargs(name))
    a, b = net.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    mine.load_state_dict(a, strict=True)
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(32))
    with torch.no_grad():
        want = net(x)
        got = hhrnet_forward_ref(net.state_dict(), x)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert ((g - w).abs().max() / w.abs().max()).item() <= 1e-5


def test_dropin_state_dict_matches_reference(ref):
    _, ref_model = ref
    a = ref_model.PoseHigherResolutionNet().state_dict()
    b = rtpe_b200.PoseHigherResolutionNet().state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    from rtpe.third_party.fp16_utils.fp16util import network_to_half as ref_n2h
    ra = ref_n2h(ref_model.PoseHigherResolutionNet()).state_dict()
    rb = rtpe_b200.get_hrnet_w48_teacher(None).state_dict()
    assert list(ra.keys()) == list(rb.keys())
    assert all(ra[k].dtype == rb[k].dtype for k in ra)


def test_student_oracle_vs_reference_class():
    """oracle/student_ref.py against rtpe.students.AttentionStudent itself (config 4 kwargs),
    same weights, same input; also pins the drop-in's state-dict layout to the reference's."""
    from oracle.ref_loader import load_reference_students
    from oracle.student_ref import attention_student_forward_ref
    from oracle.weights import fill_params_deterministic
    from rtpe_b200.students import AttentionStudent
    S = load_reference_students()
    torch.manual_seed(0)
    ref = S.AttentionStudent(None, "cpu", inplanes=48, num_heatmaps=17, ae_dims=1,
                             half_precision=False).eval()
    fill_params_deterministic(ref, 3)
    x = torch.randn(2, 3, 48, 80, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        att, det = ref(x)
    a2, d2 = attention_student_forward_ref(ref.state_dict(), x)
    assert torch.equal(att, a2) and torch.equal(det, d2)
    mine = AttentionStudent(None, "cpu", inplanes=48, num_heatmaps=17, ae_dims=1, half_precision=False)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_cam_student_oracle_vs_reference_class():
    """oracle cam_student_forward_ref against rtpe.students.CamStudent itself,
    same weights, same input, with and without the out_hw resize; also pins the drop-in's
    state-dict layout to the reference's."""
    from oracle.ref_loader import load_reference_students
    from oracle.student_ref import cam_student_forward_ref
    from oracle.weights import fill_params_deterministic
    from rtpe_b200.students import CamStudent
    S = load_reference_students()
    torch.manual_seed(0)
    ref = S.CamStudent(None, "cpu", inplanes=48, num_stages=3, num_heatmaps=17, ae_dims=1,
                       half_precision=False).eval()
    fill_params_deterministic(ref, 5)
    x = torch.randn(2, 3, 48, 80, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        (a,) = ref(x)
        (b,) = ref(x, out_hw=(30, 50))
    (a2,) = cam_student_forward_ref(ref.state_dict(), x)
    (b2,) = cam_student_forward_ref(ref.state_dict(), x, out_hw=(30, 50))
    assert torch.equal(a, a2) and torch.equal(b, b2)
    mine = CamStudent(None, "cpu", inplanes=48, num_stages=3, num_heatmaps=17, ae_dims=1,
                      half_precision=False)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_refiner_student_oracle_vs_reference_class():
    """oracle refiner_student_forward_ref against rtpe.students.RefinerStudent itself."""
    from oracle.ref_loader import load_reference_students
    from oracle.student_ref import refiner_student_forward_ref
    from oracle.weights import fill_params_deterministic
    from rtpe_b200.students import RefinerStudent
    S = load_reference_students()
    torch.manual_seed(0)
    ref = S.RefinerStudent(None, "cpu", half_precision=False).eval()
    fill_params_deterministic(ref, 7)
    x = torch.randn(2, 3, 48, 80, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        a = ref(x)
        b = ref(x, out_hw=(30, 50))
    a2 = refiner_student_forward_ref(ref.state_dict(), x)
    b2 = refiner_student_forward_ref(ref.state_dict(), x, out_hw=(30, 50))
    assert torch.equal(a, a2) and torch.equal(b, b2)
    mine = RefinerStudent(None, "cpu", half_precision=False)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_multistage_student_oracle_vs_reference_class():
    """oracle multistage_student_forward_ref against rtpe.students.MultistageStudent itself
    (students.py:389-499; built through the harness's CPU construction workaround)."""
    from oracle.ref_loader import load_reference_students
    from oracle.make_golden import build_reference_multistage
    from oracle.student_ref import multistage_student_forward_ref
    from oracle.weights import fill_params_deterministic
    from rtpe_b200.students import MultistageStudent
    S = load_reference_students()
    torch.manual_seed(0)
    ref = build_reference_multistage(S, layers_per_stage=[2, 3])
    fill_params_deterministic(ref, 7)
    x = torch.randn(2, 3, 48, 80, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        a = ref(x)
        b = ref(x, out_hw=(30, 50))
    a2 = multistage_student_forward_ref(ref.state_dict(), x)
    b2 = multistage_student_forward_ref(ref.state_dict(), x, out_hw=(30, 50))
    assert len(a) == len(a2) == 2 and all(torch.equal(p, q) for p, q in zip(a, a2))
    assert all(torch.equal(p, q) for p, q in zip(b, b2))
    mine = MultistageStudent(None, "cpu", layers_per_stage=[2, 3], half_precision=False)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_attention_student_steps_oracle_vs_reference_class():
    """oracle attention_student_steps_forward_ref against rtpe.students.AttentionStudentSteps itself
    (students.py:786-1073), eval_attention.py's inplanes = 80, with and without att_divisor."""
    import warnings
    from oracle.ref_loader import load_reference_students
    from oracle.student_ref import attention_student_steps_forward_ref
    from oracle.weights import fill_params_deterministic
    from rtpe_b200.students import AttentionStudentSteps
    S = load_reference_students()
    torch.manual_seed(0)
    ref = S.AttentionStudentSteps(None, "cpu", 80, 17, 0, False).eval()
    fill_params_deterministic(ref, 7)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 48, 80, generator=g)
    alt = torch.randn(2, 3, 48, 80, generator=g)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a, d = ref(x, alt=alt)
        a20, d20 = ref(x, alt=alt, att_divisor=20)
    a2, d2 = attention_student_steps_forward_ref(ref.state_dict(), x, alt)
    a3, d3 = attention_student_steps_forward_ref(ref.state_dict(), x, alt, att_divisor=20)
    assert torch.equal(a, a2) and torch.equal(d, d2) and torch.equal(a20, a3) and torch.equal(d20, d3)
    mine = AttentionStudentSteps(None, "cpu", 80, 17, 0, False)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    mine.load_state_dict(ref.state_dict(), strict=True)
    with pytest.raises(NotImplementedError):
        mine(x)                                      # "ATM alt is expected" (students.py:983)
