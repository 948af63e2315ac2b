"""ORACLE (test infrastructure only) -- functional CPU restatement of
``AttentionStudent.forward`` (rtpe/students.py:733-768) and the blocks it calls:
``StemHRNet.forward`` (:252-264; Bottleneck = pose_higher_hrnet.py:78-116), ``mid_stem``
(:617-629), ``ContextAwareModule.forward`` (:181-201), ``SELayer.forward`` (:137-142), the
attention / detection pyramids (:653-713).

It consumes a plain ``state_dict`` with the reference's key names and evaluates the student with
torch CPU operators in a chosen dtype, eval-mode BatchNorm.  Quirks of the reference that are
kept on purpose: ``mid`` and ``lo`` of both pyramids are the SAME upsampled ``lo`` map (so the sum
is hi + 2*up(lo)), ``det_hi`` is applied twice to the same input and ``det_mid`` is never used
(:756-761), the returned attention map is ``sigmoid(att / 20)`` (:752).

Pinned against the reference class itself (same weights, same input) in
tests/test_oracle_vs_reference.py and through tests/golden/student_64x96.npz.
Only tests/, smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"].to(x.dtype), sd[p + ".running_var"].to(x.dtype),
                        sd[p + ".weight"].to(x.dtype), sd[p + ".bias"].to(x.dtype), False, 0.0, EPS)


def _conv(x, sd, p, stride=1, padding=0, dilation=1):
    b = sd.get(p + ".bias")
    return F.conv2d(x, sd[p + ".weight"].to(x.dtype), None if b is None else b.to(x.dtype),
                    stride=stride, padding=padding, dilation=dilation)


def _bottleneck(x, sd, p):
    """pose_higher_hrnet.py:96-116."""
    out = F.relu(_bn(_conv(x, sd, p + ".conv1"), sd, p + ".bn1"))
    out = F.relu(_bn(_conv(out, sd, p + ".conv2", padding=1), sd, p + ".bn2"))
    out = _bn(_conv(out, sd, p + ".conv3"), sd, p + ".bn3")
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(_conv(x, sd, p + ".downsample.0"), sd, p + ".downsample.1")
    return F.relu(out + res)


def stem_ref(sd, x, p="stem.1"):
    """students.py:252-264."""
    x = F.relu(_bn(_conv(x, sd, p + ".conv1", stride=2, padding=1), sd, p + ".bn1"))
    x = F.relu(_bn(_conv(x, sd, p + ".conv2", stride=2, padding=1), sd, p + ".bn2"))
    for i in range(4):
        x = _bottleneck(x, sd, "%s.layer1.%d" % (p, i))
    return x


def cam_ref(sd, x, p, dilations=None):
    """students.py:181-201 (no HDC upsampling: the dilated convs keep the size).  ``dilations``:
    the module's ``hdc_dilations`` (default 1..n, the attention student's)."""
    residual = F.relu(_bn(_conv(x, sd, p + ".residual.0"), sd, p + ".residual.1"))
    y = x.mean(dim=(2, 3))                                                    # AdaptiveAvgPool2d(1)
    y = F.relu(F.linear(y, sd[p + ".se.fc.0.weight"].to(x.dtype), sd[p + ".se.fc.0.bias"].to(x.dtype)))
    y = torch.sigmoid(F.linear(y, sd[p + ".se.fc.2.weight"].to(x.dtype), sd[p + ".se.fc.2.bias"].to(x.dtype)))
    outs = []
    i = 0
    while (p + ".hdcs.%d.0.weight" % i) in sd:
        d = i + 1 if dilations is None else dilations[i]                      # 1..n (:659,:688)
        outs.append(F.relu(_bn(_conv(x, sd, p + ".hdcs.%d.0" % i, padding=d, dilation=d), sd,
                               p + ".hdcs.%d.1" % i)))
        i += 1
    out = F.relu(_bn(_conv(torch.cat(outs, dim=1), sd, p + ".hdc_top.0"), sd, p + ".hdc_top.1"))
    return F.relu(residual + out * y[:, :, None, None])


def _pool(x):
    return F.avg_pool2d(x, kernel_size=3, stride=2, padding=1, count_include_pad=False)


@torch.no_grad()
def attention_student_forward_ref(state_dict, x, dtype=torch.float32):
    """-> (att (N,1,H/4,W/4), det (N,C,H/4,W/4)) like AttentionStudent.forward."""
    sd = state_dict
    x = x.to(dtype)
    s = stem_ref(sd, x)
    s = F.relu(_bn(_conv(s, sd, "mid_stem.0", padding=1), sd, "mid_stem.1"))
    s = F.relu(_bn(_conv(s, sd, "mid_stem.3", padding=1), sd, "mid_stem.4"))
    hw = s.shape[-2:]
    hi = cam_ref(sd, s, "att_hi.0")
    mid = cam_ref(sd, _pool(s), "att_mid.1")
    lo = cam_ref(sd, _pool(mid), "att_lo.1")
    lo_up = F.interpolate(lo, hw, mode="nearest")
    att = hi + lo_up + lo_up
    att = _conv(att, sd, "att_top.0", padding=1)
    att = torch.sigmoid(att / 20)
    s = s + att.expand(s.shape)
    hi = cam_ref(sd, s, "det_hi.0")
    mid = hi                                            # det_hi applied twice to the same input
    lo = cam_ref(sd, _pool(mid), "det_lo.1")
    lo_up = F.interpolate(lo, hw, mode="nearest")
    det = hi + lo_up + lo_up
    det = _conv(det, sd, "det_top.0", padding=1)
    return att, det


CAM_STUDENT_DILATIONS = (1, 2, 3, 5, 8, 12)                                  # students.py:563


@torch.no_grad()
def cam_student_forward_ref(state_dict, x, out_hw=None, dtype=torch.float32):
    """``CamStudent.forward`` (students.py:568-592): -> [pred]; the sum of all context-aware
    modules applied to the SAME mid-stem output, then ``hm_convs[-1]`` only (:581)."""
    sd = state_dict
    x = x.to(dtype)
    s = stem_ref(sd, x)
    s = F.relu(_bn(_conv(s, sd, "mid_stem.0", padding=1), sd, "mid_stem.1"))
    n = 0
    while ("cams.%d.residual.0.weight" % n) in sd:
        n += 1
    acc = cam_ref(sd, s, "cams.0", CAM_STUDENT_DILATIONS)
    for i in range(1, n):
        acc = acc + cam_ref(sd, s, "cams.%d" % i, CAM_STUDENT_DILATIONS)
    out = _conv(acc, sd, "hm_convs.%d" % (n - 1), padding=1)
    if out_hw is not None:
        out = F.interpolate(out, out_hw, mode="bilinear", align_corners=True)
    return [out]


def skip_conv_ref(sd, x, p):
    """SkipConv.forward (students.py:73-90) with 3x3 / padding-1 convs (get_straight_skip_conv)."""
    residual = _bn(_conv(x, sd, p + ".downsample.0"), sd, p + ".downsample.1")
    i = 0
    while (p + ".convs.%d.weight" % i) in sd:
        k = sd[p + ".convs.%d.weight" % i].shape[-1]
        x = F.relu(_bn(_conv(x, sd, p + ".convs.%d" % i, padding=k // 2), sd, p + ".bns.%d" % i))
        i += 1
    return F.relu(x + residual)


@torch.no_grad()
def refiner_student_forward_ref(state_dict, x, out_hw=None, dtype=torch.float32):
    """``RefinerStudent.forward`` (students.py:374-386)."""
    sd = state_dict
    s = stem_ref(sd, x.to(dtype))
    n = 0
    while ("stages.%d.convs.0.weight" % n) in sd:
        n += 1
    y = skip_conv_ref(sd, s, "stages.0")
    for i in range(1, n):
        y = skip_conv_ref(sd, s + y, "stages.%d" % i)
    if out_hw is not None:
        y = F.interpolate(y, out_hw, mode="bilinear", align_corners=True)
    return y


@torch.no_grad()
def multistage_student_forward_ref(state_dict, x, out_hw=None, dtype=torch.float32):
    """``MultistageStudent.forward`` (students.py:473-499) -> list of stage outputs.  With
    ``out_hw`` the stem output is resized first, the stages run at that size and the per-stage
    resizes (same size, align_corners=True) are kept as the reference has them."""
    sd = state_dict
    s = stem_ref(sd, x.to(dtype))
    if out_hw is not None:
        s = F.interpolate(s, out_hw, mode="bilinear", align_corners=True)
    n = 0
    while ("stages.%d.convs.0.weight" % n) in sd:
        n += 1
    y = skip_conv_ref(sd, s, "stages.0")
    if out_hw is not None:
        y = F.interpolate(y, out_hw, mode="bilinear", align_corners=True)
    outs = [y]
    for i in range(1, n):
        y = skip_conv_ref(sd, torch.cat([s, outs[-1]], dim=1), "stages.%d" % i)
        if out_hw is not None:
            y = F.interpolate(y, out_hw, mode="bilinear", align_corners=True)
        outs.append(y)
    return outs


@torch.no_grad()
def attention_student_steps_forward_ref(state_dict, x, alt, att_divisor=None, dtype=torch.float32):
    """``AttentionStudentSteps.forward`` (students.py:966-1052) -> (att, det).  Quirks kept: ``mid``
    and ``lo`` are the same up-sampled ``lo`` map (:993-996), the attention multiplies the
    concatenated (stem, image) tensor (:1018), ``out_hw`` is unused."""
    sd = state_dict
    x, alt = x.to(dtype), alt.to(dtype)
    s = stem_ref(sd, x)
    s = F.relu(_bn(_conv(s, sd, "mid_stem.0", padding=1), sd, "mid_stem.1"))
    s = F.relu(_bn(_conv(s, sd, "mid_stem.3", padding=1), sd, "mid_stem.4"))
    a = F.relu(_bn(_conv(alt, sd, "alt_img_stem.0", stride=2, padding=2), sd, "alt_img_stem.1"))
    a = F.relu(_bn(_conv(a, sd, "alt_img_stem.3", stride=2, padding=2), sd, "alt_img_stem.4"))
    hw = s.shape[-2:]
    s = torch.cat((s, F.interpolate(alt, hw, mode="bilinear")), dim=1)
    dil4 = (1, 2, 3, 4)
    hi = cam_ref(sd, s, "att_hi.0", dil4)
    mid = cam_ref(sd, _pool(s), "att_mid.1", dil4)
    lo = cam_ref(sd, _pool(mid), "att_lo.1", dil4)
    lo_up = F.interpolate(lo, hw, mode="nearest")
    att = _conv(hi + lo_up + lo_up, sd, "att_top.0", padding=1)
    if att_divisor is not None:
        att = att / att_divisor
    att = torch.sigmoid(att)
    y = torch.cat((s * att.expand(s.shape), a), dim=1)
    i = 0
    while ("steps.%d.residual.0.weight" % i) in sd:
        y = cam_ref(sd, y, "steps.%d" % i, (1, 2, 3))
        i += 1
    det = _conv(y, sd, "steps.%d" % i, padding=1)
    return att, det
