"""CPU: the oracle restatements reproduce the fixtures that the UNMODIFIED reference
produced (oracle/make_golden.py).  Decode is bit-exact; the model is float32 torch CPU on
both sides, compared to 1e-5 of the tensor max (conv algorithm choice may differ by size)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import group_ref as G
from oracle.hhrnet_ref import hhrnet_forward_ref
from oracle.weights import fill_params_deterministic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_decode_fixture(path):
    z = np.load(path)
    det = z["det"].astype(np.float32)
    tag = z["tag"]
    tpj = tag.shape[1] == det.shape[1]
    return z, det, tag, tpj


def split_people(z):
    out, off = [], 0
    for c in z["counts"]:
        out.append((z["people"][off:off + c], z["scores"][off:off + c]))
        off += c
    return out


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "decode_*.npz"))))
def test_decode_oracle_matches_reference_fixture(path):
    z, det, tag, tpj = load_decode_fixture(path)
    p = G.DecodeParams(tag_per_joint=tpj)
    tk = G.top_k_ref(det, tag, p)
    # values are order-independent among ties; the fixtures have none above the threshold
    assert np.array_equal(tk["val_k"], z["val_k"])
    live = z["val_k"] > 0
    assert np.array_equal(tk["loc_k"][live], z["loc_k"].astype(np.int64)[live])
    assert np.array_equal(tk["tag_k"][live], z["tag_k"][live])
    got = G.parse_batch_ref(det.copy(), tag.copy(), p, True, True)
    want = split_people(z)
    assert len(got) == len(want)
    for (gp, gs), (wp, ws) in zip(got, want):
        gp = np.asarray(gp, np.float32).reshape((-1,) + wp.shape[1:])
        assert np.array_equal(gp, wp)
        assert np.array_equal(np.asarray(gs, np.float32), ws)


def test_hhrnet_oracle_matches_reference_fixture():
    import rtpe_b200
    z = np.load(os.path.join(GOLD, "hhrnet_64x96.npz"))
    net = rtpe_b200.PoseHigherResolutionNet()
    assert sum(p.numel() for p in net.parameters()) == int(z["params"]) == 63827139
    fill_params_deterministic(net, int(z["seed"]))
    with torch.no_grad():
        y0, y1 = hhrnet_forward_ref(net.state_dict(), torch.from_numpy(z["x"]))
    for got, want in ((y0, z["y0"]), (y1, z["y1"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5


def test_student_oracle_matches_reference_fixture():
    """AttentionStudent (config 4): the drop-in's parameter tree filled by (order, shape) gives
    the weights the reference had, and the oracle reproduces the reference's outputs."""
    from rtpe_b200.students import AttentionStudent
    from oracle.student_ref import attention_student_forward_ref
    z = np.load(os.path.join(GOLD, "student_64x96.npz"))
    net = AttentionStudent(None, "cpu", inplanes=48, num_heatmaps=17, ae_dims=1, half_precision=False)
    assert len(net.state_dict()) == int(z["entries"]) == 364
    fill_params_deterministic(net, int(z["seed"]))
    att, det = attention_student_forward_ref(net.state_dict(), torch.from_numpy(z["x"]))
    for got, want in ((att, z["att"]), (det, z["det"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5


def test_cam_student_oracle_matches_reference_fixture():
    """CamStudent (students.py:502-592): the drop-in's parameter tree filled by (order, shape) gives
    the weights the reference had, and the oracle reproduces the reference's outputs."""
    from rtpe_b200.students import CamStudent
    from oracle.student_ref import cam_student_forward_ref
    z = np.load(os.path.join(GOLD, "cam_student_64x96.npz"))
    net = CamStudent(None, "cpu", inplanes=48, num_stages=3, num_heatmaps=17, ae_dims=1,
                     half_precision=False)
    assert len(net.state_dict()) == int(z["entries"])
    fill_params_deterministic(net, int(z["seed"]))
    x = torch.from_numpy(z["x"])
    (pred,) = cam_student_forward_ref(net.state_dict(), x)
    (pred_up,) = cam_student_forward_ref(net.state_dict(), x, out_hw=tuple(z["pred_up"].shape[2:]))
    for got, want in ((pred, z["pred"]), (pred_up, z["pred_up"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5


def test_refiner_student_oracle_matches_reference_fixture():
    """RefinerStudent (students.py:302-386)."""
    from rtpe_b200.students import RefinerStudent
    from oracle.student_ref import refiner_student_forward_ref
    z = np.load(os.path.join(GOLD, "refiner_student_64x96.npz"))
    net = RefinerStudent(None, "cpu", half_precision=False)
    assert len(net.state_dict()) == int(z["entries"])
    fill_params_deterministic(net, int(z["seed"]))
    x = torch.from_numpy(z["x"])
    pred = refiner_student_forward_ref(net.state_dict(), x)
    pred_up = refiner_student_forward_ref(net.state_dict(), x, out_hw=tuple(z["pred_up"].shape[2:]))
    for got, want in ((pred, z["pred"]), (pred_up, z["pred_up"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5


def test_multistage_student_oracle_matches_reference_fixture():
    """MultistageStudent (students.py:389-499): every stage output, with and without out_hw."""
    from rtpe_b200.students import MultistageStudent
    from oracle.student_ref import multistage_student_forward_ref
    z = np.load(os.path.join(GOLD, "multistage_student_64x96.npz"))
    net = MultistageStudent(None, "cpu", half_precision=False)
    assert len(net.state_dict()) == int(z["entries"])
    fill_params_deterministic(net, int(z["seed"]))
    x = torch.from_numpy(z["x"])
    outs = multistage_student_forward_ref(net.state_dict(), x)
    outs_up = multistage_student_forward_ref(net.state_dict(), x, out_hw=tuple(z["outs_up"].shape[3:]))
    for got, want in ((outs, z["outs"]), (outs_up, z["outs_up"])):
        assert len(got) == want.shape[0]
        for g, w_ in zip(got, want):
            w_ = torch.from_numpy(w_)
            assert g.shape == w_.shape
            assert ((g - w_).abs().max() / w_.abs().max()).item() <= 1e-5


def test_attention_student_steps_oracle_matches_reference_fixture():
    """AttentionStudentSteps (students.py:786-1073), with and without att_divisor."""
    from rtpe_b200.students import AttentionStudentSteps
    from oracle.student_ref import attention_student_steps_forward_ref
    z = np.load(os.path.join(GOLD, "attention_steps_64x96.npz"))
    net = AttentionStudentSteps(None, "cpu", int(z["inplanes"]), 17, 1, False)
    assert len(net.state_dict()) == int(z["entries"])
    fill_params_deterministic(net, int(z["seed"]))
    x, alt = torch.from_numpy(z["x"]), torch.from_numpy(z["alt"])
    att, det = attention_student_steps_forward_ref(net.state_dict(), x, alt)
    att20, det20 = attention_student_steps_forward_ref(net.state_dict(), x, alt, att_divisor=20)
    for got, key in ((att, "att"), (det, "det"), (att20, "att20"), (det20, "det20")):
        want = torch.from_numpy(z[key])
        assert got.shape == want.shape
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5


def test_config1_oracle_pipeline_matches_reference_fixture():
    """BASELINE.json configs[0] (validate_hhrnet.py:84-105 on a bundled JPEG, default init under seed
    0): JPEG bytes -> oracle pre-processing -> oracle forward -> in-tree aggregation -> oracle parse,
    against what the UNMODIFIED reference produced (oracle/make_golden_config1.py).  Float maps to
    1e-5 of the tensor max (conv algorithm choice), the decode result exactly when the maps are
    bit-identical, else to the quarter-pixel."""
    import io
    import rtpe_b200
    from PIL import Image
    from oracle import preprocess_ref
    from oracle.aggregate_ref import aggregate_intree_ref
    z = np.load(os.path.join(GOLD, "config1_bundled.npz"))
    st, stb = int(z["stride"]), int(z["stride_big"])
    torch.manual_seed(0)
    sd = rtpe_b200.PoseHigherResolutionNet().state_dict()
    p = "img1_"                                    # 555 x 640 -> 640 x 768 (the smaller of the two)
    img = np.array(Image.open(io.BytesIO(z[p + "jpeg"].tobytes())).convert("RGB"))
    h, w = int(z[p + "hw"][0]), int(z[p + "hw"][1])
    assert img.shape[:2] == (h, w)
    u8, center, scale = preprocess_ref.resize_align_multi_scale(img, 640, 1, 1)
    t = torch.from_numpy(preprocess_ref.to_tensor_normalize(u8))[None]
    assert tuple(t.shape) == tuple(int(v) for v in z[p + "input_shape"])
    assert torch.equal(t[:, :, ::stb, ::stb], torch.from_numpy(z[p + "input_s"]))
    with torch.no_grad():
        y0, y1 = hhrnet_forward_ref(sd, t)
    for got, key in ((y0, "y0"), (y1, "y1")):
        want = torch.from_numpy(z[p + key + "_s"])
        assert (got[:, :, ::st, ::st] - want).abs().max().item() <= 1e-5 * float(z[p + key + "_absmax"])
    det, tag = aggregate_intree_ref(y0, y1, (h, w), 17)
    exact = True
    for got, key in ((det, "hms"), (tag[..., 0], "aes")):
        want = torch.from_numpy(z[p + key + "_s"])
        assert (got[:, :, ::stb, ::stb] - want).abs().max().item() <= 1e-5 * float(z[p + key + "_absmax"])
        exact = exact and torch.equal(got[:, :, ::stb, ::stb], want)
    people, scores = G.parse_image_ref(det.numpy(), tag.numpy(), G.DecodeParams(), True, True)
    people = np.asarray(people, np.float32)
    assert people.shape == z[p + "people"].shape
    if exact:
        assert np.array_equal(people, z[p + "people"])
        assert np.array_equal(np.asarray(scores, np.float32), z[p + "scores"])
