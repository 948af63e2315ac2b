"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference: rtpe/third_party/group.py and pose_higher_hrnet.py) in this container.

    python -m oracle.make_golden

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path
(tests/test_golden_gpu.py).  `munkres` is absent from the image; the reference's
``from munkres import Munkres`` is satisfied by oracle/munkres_ref.py, so the Hungarian
step of these fixtures is only as pinned as that restatement (see its header).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, load_reference_students  # noqa: E402
from oracle.weights import fill_params_deterministic  # noqa: E402
import rtpe_b200  # noqa: E402  (only its torch-side synthetic generator is used)

OUT = os.path.join(ROOT, "tests", "golden")

PARSER = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
              use_detection_val=True, ignore_too_much=False, tag_per_joint=True, nms_ksize=5,
              nms_padding=2)


def decode_fixture(ref_group, name, n, h, w, t, people, seed, quantise=False, tag_scale=1.0,
                   tag_per_joint=True):
    det, tag = rtpe_b200.synth_decode_batch(n, height=h, width=w, tag_dims=t, max_people=people,
                                            seed=seed, tag_per_joint=tag_per_joint)
    if quantise:
        tag = tag.to(torch.bfloat16).to(torch.float32)
    tag = tag * tag_scale
    kw = dict(PARSER)
    kw["tag_per_joint"] = tag_per_joint
    hp = ref_group.HeatmapParser(**kw)
    # float32 on purpose: quantising det would create equal peak values, whose top-k order
    # is undefined in the reference (torch.topk), and the fixture would stop being a pin.
    out = {"det": det.numpy(), "tag": tag.numpy()}
    counts, people_all, scores_all = [], [], []
    for i in range(n):
        ans, scores = hp.parse(det[i:i + 1].clone(), tag[i:i + 1].clone(), True, True)
        arr = np.asarray(ans[0], np.float32)
        if arr.size == 0:
            arr = np.zeros((0, 17, 3 + t), np.float32)
        counts.append(arr.shape[0])
        people_all.append(arr)
        scores_all.append(np.asarray(scores, np.float32))
    out["counts"] = np.asarray(counts, np.int32)
    out["people"] = np.concatenate(people_all, 0)
    out["scores"] = np.concatenate(scores_all, 0) if scores_all else np.zeros((0,), np.float32)
    tk = hp.top_k(det, tag)
    # canonical tie order (value desc, index asc) -- reference order among equals is undefined
    out["val_k"] = tk["val_k"]
    out["loc_k"] = tk["loc_k"].astype(np.int32)
    out["tag_k"] = tk["tag_k"]
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "people per image", counts)


def model_fixture(ref_model, name, h, w, seed):
    torch.manual_seed(0)
    net = ref_model.PoseHigherResolutionNet().eval()
    fill_params_deterministic(net, seed)
    x = torch.randn(1, 3, h, w, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        y0, y1 = net(x)
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), y0=y0.numpy(), y1=y1.numpy(),
                        seed=np.int64(seed), params=np.int64(sum(p.numel() for p in net.parameters())))
    print(name, tuple(y0.shape), tuple(y1.shape), float(y0.abs().max()), float(y1.abs().max()))


def student_fixture(ref_students, name, h, w, seed):
    """AttentionStudent (rtpe/students.py:595-771), BASELINE config 4 hyper-parameters."""
    torch.manual_seed(0)
    net = ref_students.AttentionStudent(None, "cpu", inplanes=48, num_heatmaps=17, ae_dims=1,
                                        half_precision=False).eval()
    fill_params_deterministic(net, seed)
    x = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        att, det = net(x)
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), att=att.numpy(), det=det.numpy(),
                        seed=np.int64(seed), entries=np.int64(len(net.state_dict())))
    print(name, tuple(att.shape), tuple(det.shape), float(att.min()), float(att.max()),
          float(det.abs().max()))


def cam_student_fixture(ref_students, name, h, w, seed):
    """CamStudent (rtpe/students.py:502-592), default hyper-parameters, fp32."""
    torch.manual_seed(0)
    net = ref_students.CamStudent(None, "cpu", inplanes=48, num_stages=3, num_heatmaps=17, ae_dims=1,
                                  half_precision=False).eval()
    fill_params_deterministic(net, seed)
    x = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        (pred,) = net(x)
        (pred_up,) = net(x, out_hw=(21, 35))
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), pred=pred.numpy(), pred_up=pred_up.numpy(),
                        seed=np.int64(seed), entries=np.int64(len(net.state_dict())))
    print(name, tuple(pred.shape), tuple(pred_up.shape), float(pred.abs().max()))


def refiner_student_fixture(ref_students, name, h, w, seed):
    """RefinerStudent (rtpe/students.py:302-386), default hyper-parameters, fp32."""
    torch.manual_seed(0)
    net = ref_students.RefinerStudent(None, "cpu", half_precision=False).eval()
    fill_params_deterministic(net, seed)
    x = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        pred = net(x)
        pred_up = net(x, out_hw=(21, 35))
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), pred=pred.numpy(), pred_up=pred_up.numpy(),
                        seed=np.int64(seed), entries=np.int64(len(net.state_dict())))
    print(name, tuple(pred.shape), tuple(pred_up.shape), float(pred.abs().max()))


def build_reference_multistage(ref_students, **kw):
    """The reference's MultistageStudent.__init__ first runs RefinerStudent.__init__ with ITS default
    device "cuda" (students.py:405 -> :305-335), so the class cannot be built on a CPU-only host as
    it stands: the harness swaps that one default for "cpu" while constructing (no file is touched)."""
    init = ref_students.RefinerStudent.__init__
    keep = init.__defaults__
    init.__defaults__ = tuple("cpu" if d == "cuda" else d for d in keep)
    try:
        return ref_students.MultistageStudent(None, "cpu", half_precision=False, **kw).eval()
    finally:
        init.__defaults__ = keep


def multistage_student_fixture(ref_students, name, h, w, seed):
    """MultistageStudent (rtpe/students.py:389-499), default hyper-parameters, fp32."""
    torch.manual_seed(0)
    net = build_reference_multistage(ref_students)
    fill_params_deterministic(net, seed)
    x = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        outs = net(x)
        outs_up = net(x, out_hw=(21, 35))
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), outs=np.stack([o.numpy() for o in outs]),
                        outs_up=np.stack([o.numpy() for o in outs_up]),
                        seed=np.int64(seed), entries=np.int64(len(net.state_dict())))
    print(name, len(outs), tuple(outs[0].shape), tuple(outs_up[0].shape), float(outs[-1].abs().max()))


def attention_steps_fixture(ref_students, name, h, w, seed, inplanes=48):
    """AttentionStudentSteps (rtpe/students.py:786-1073), fp32, with and without att_divisor."""
    import warnings
    torch.manual_seed(0)
    net = ref_students.AttentionStudentSteps(None, "cpu", inplanes, 17, 1, False).eval()
    fill_params_deterministic(net, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(2, 3, h, w, generator=g)
    alt = torch.randn(2, 3, h, w, generator=g)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        att, det = net(x, alt=alt)
        att20, det20 = net(x, alt=alt, att_divisor=20)
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), alt=alt.numpy(), att=att.numpy(),
                        det=det.numpy(), att20=att20.numpy(), det20=det20.numpy(),
                        inplanes=np.int64(inplanes), seed=np.int64(seed),
                        entries=np.int64(len(net.state_dict())))
    print(name, tuple(att.shape), tuple(det.shape), float(det.abs().max()), float(att.min()),
          float(att.max()))


def main():
    if "--new-students" in sys.argv:
        multistage_student_fixture(load_reference_students(), "multistage_student_64x96.npz", 64, 96,
                                   seed=12)
        attention_steps_fixture(load_reference_students(), "attention_steps_64x96.npz", 64, 96, seed=13)
        return
    os.makedirs(OUT, exist_ok=True)
    ref_group, ref_model = load_reference()
    decode_fixture(ref_group, "decode_a.npz", 2, 56, 72, 1, 6, seed=21)
    decode_fixture(ref_group, "decode_flip_t2.npz", 2, 48, 64, 2, 10, seed=22)
    decode_fixture(ref_group, "decode_collisions.npz", 2, 64, 64, 1, 30, seed=23, quantise=True,
                   tag_scale=0.25)
    decode_fixture(ref_group, "decode_shared_tag.npz", 2, 48, 56, 1, 4, seed=24, tag_per_joint=False)
    model_fixture(ref_model, "hhrnet_64x96.npz", 64, 96, seed=7)
    student_fixture(load_reference_students(), "student_64x96.npz", 64, 96, seed=9)
    cam_student_fixture(load_reference_students(), "cam_student_64x96.npz", 64, 96, seed=10)
    refiner_student_fixture(load_reference_students(), "refiner_student_64x96.npz", 64, 96, seed=11)
    multistage_student_fixture(load_reference_students(), "multistage_student_64x96.npz", 64, 96, seed=12)
    attention_steps_fixture(load_reference_students(), "attention_steps_64x96.npz", 64, 96, seed=13)


if __name__ == "__main__":
    main()
