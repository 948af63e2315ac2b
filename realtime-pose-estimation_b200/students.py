"""The students of rtpe/students.py, executed by libbrtpe.so: ``AttentionStudent`` (:595-771, BASELINE
config 4) with its building blocks ``SELayer`` (:118-142), ``ContextAwareModule`` (:145-201) and
``StemHRNet`` (:206-282); ``CamStudent`` (:502-592); ``SkipConv`` (:37-112) with ``RefinerStudent``
(:302-386) and ``MultistageStudent`` (:389-499); ``AttentionStudentSteps`` (:786-1073).

Module tree and parameter names are the reference's (``stem.1.conv1.weight``,
``att_lo.1.hdcs.3.0.weight``, ``det_top.0.bias`` ...), so ``load_state_dicts`` snapshots and
``StemHRNet.load_pretrained`` work unchanged.  Like ``PoseHigherResolutionNet`` the parameters
are containers: the forward pass is a recorded launch plan (NHWC activations, BN folded into
the convolutions, dilated 3x3 convolutions through the tap table of the conv engine, the five
dilation branches writing channel slices of one buffer instead of ``torch.cat``).

Precision: ``half_precision=False`` (BASELINE config 4) -> fp32 activations on the CUDA-core
path (<= 1e-4 of the tensor max); ``half_precision=True`` -> the whole student runs in bf16 on
the tcgen05 path where the layer shapes allow it (the reference halves only the stem).

``forward(x) -> (att, det)``: att (N,1,H/4,W/4) = sigmoid(attention logits / 20) exactly as the
reference returns it, det (N, num_heatmaps + ae_dims, H/4, W/4).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .hhrnet import BN_MOMENTUM, Bottleneck, _PlanRunner, _Recorder
from .precision import network_to_half

AUX_AVGPOOL, AUX_SE_PARTIAL, AUX_SE_GATE, AUX_CAM_MIX, AUX_ATT_ADD = 1, 2, 3, 4, 5
AUX_RESIZE_NHWC, AUX_IMAGE_NHWC, AUX_S2D, AUX_ATT_MUL = 6, 7, 8, 9


def init_weights(module, init_fn=torch.nn.init.kaiming_normal_, bias_val=0.0):
    """rtpe/students.py:20-31."""
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        init_fn(module.weight)
        if module.bias is not None:
            module.bias.data.fill_(bias_val)


class SELayer(nn.Module):
    """Parameter container of students.py:118-142 (returns the gate, not x * gate)."""

    def __init__(self, in_chans, hidden_chans=None, bn_momentum=0.1):
        super().__init__()
        if hidden_chans is None:
            hidden_chans = in_chans // 4
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(in_chans, hidden_chans, bias=True), nn.ReLU(inplace=True),
                                nn.Linear(hidden_chans, in_chans, bias=True), nn.Sigmoid())


class ContextAwareModule(nn.Module):
    """Parameter container of students.py:145-201."""

    def __init__(self, in_chans, se_chans=None, hdc_dilations=[1, 2, 3, 4], hdc_chans=None,
                 bn_momentum=0.1):
        super().__init__()
        self.residual = nn.Sequential(
            nn.Conv2d(in_chans, in_chans, kernel_size=1, stride=1, bias=False),
            nn.BatchNorm2d(in_chans, momentum=bn_momentum), nn.ReLU(inplace=True))
        self.se = SELayer(in_chans, se_chans, bn_momentum)
        if hdc_chans is None:
            hdc_chans = in_chans // 4
        self.hdcs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(in_chans, hdc_chans, kernel_size=3, stride=1, dilation=d,
                                    padding=d, bias=False),
                          nn.BatchNorm2d(hdc_chans, momentum=bn_momentum), nn.ReLU(inplace=True))
            for d in hdc_dilations])
        self.hdc_top = nn.Sequential(
            nn.Conv2d(hdc_chans * len(hdc_dilations), in_chans, kernel_size=1, bias=False),
            nn.BatchNorm2d(in_chans, momentum=bn_momentum), nn.ReLU(inplace=True))
        self.final_relu = nn.ReLU(inplace=True)


class StemHRNet(nn.Module):
    """Parameter container of students.py:206-264 (the HigherHRNet stem: conv1, conv2, layer1)."""
    INPLANES = 64

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, self.INPLANES, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(self.INPLANES, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(self.INPLANES, self.INPLANES, kernel_size=3, stride=2, padding=1,
                               bias=False)
        self.bn2 = nn.BatchNorm2d(self.INPLANES, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer1(self.INPLANES, 4)

    def _make_layer1(self, planes=64, blocks=4):
        expansion = Bottleneck.expansion
        residual_fn = nn.Sequential(
            nn.Conv2d(self.INPLANES, planes * expansion, kernel_size=1, stride=1, bias=False),
            nn.BatchNorm2d(planes * expansion, momentum=BN_MOMENTUM))
        layers = [Bottleneck(64, 64, 1, residual_fn)]
        layers += [Bottleneck(planes * expansion, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def load_pretrained(self, hhrnet_statedict_path, device="cpu", check=False):
        """students.py:266-282: the stem part of a HigherHRNet checkpoint (``"1."`` prefix)."""
        hhrnet_d = torch.load(hhrnet_statedict_path, map_location=device)
        self.load_state_dict({k: hhrnet_d["1." + k] for k in self.state_dict()})
        if check:
            assert all((hhrnet_d["1." + k].to(device) == v.to(device)).all()
                       for k, v in self.state_dict().items()), "Error loading statedict!"


class AttentionStudent(_PlanRunner, nn.Module):
    """Same constructor, attributes and ``forward(x) -> (att, det)`` as students.py:595-771."""

    def __init__(self, hhrnet_statedict_path=None, device="cuda", inplanes=48, num_heatmaps=17,
                 ae_dims=1, half_precision=True, init_fn=torch.nn.init.kaiming_normal_,
                 trainable_stem=False, bn_momentum=0.1):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.num_heatmaps = num_heatmaps
        self.ae_dims = ae_dims
        self.stem = StemHRNet()
        self.stem_out_chans = self.stem.layer1[-1].bn3.num_features
        self.trainable_stem = trainable_stem
        self.inplanes = inplanes
        mid_inplanes = (self.stem_out_chans + self.inplanes) // 2
        self.mid_stem = nn.Sequential(
            nn.Conv2d(self.stem_out_chans, mid_inplanes, kernel_size=3, stride=1, dilation=1,
                      padding=1, bias=False),
            nn.BatchNorm2d(mid_inplanes, momentum=bn_momentum), nn.ReLU(inplace=True),
            nn.Conv2d(mid_inplanes, inplanes, kernel_size=3, stride=1, dilation=1, padding=1,
                      bias=False),
            nn.BatchNorm2d(inplanes, momentum=bn_momentum), nn.ReLU(inplace=True))
        self.att_lo, self.att_mid, self.att_hi, self.att_top = self._attention_body()
        self.det_lo, self.det_mid, self.det_hi, self.det_top = self._detection_body_v1()
        if init_fn is not None:
            self.apply(lambda module: init_weights(module, init_fn, 0.0))
        self.half_precision = bool(half_precision)
        if half_precision:
            self.stem = network_to_half(self.stem)
        else:
            self.stem = nn.Sequential(nn.Identity(), self.stem)
        if hhrnet_statedict_path is not None:
            self.stem[1].load_pretrained(hhrnet_statedict_path, device, check=False)
        self._init_runner()
        self.to(device)
        self.device = device

    def _body(self, dilations, out_chans):
        def cam():
            return ContextAwareModule(self.inplanes, hdc_dilations=list(dilations))

        def pool():
            return nn.AvgPool2d(kernel_size=3, stride=2, padding=1, count_include_pad=False)
        low_res = nn.Sequential(pool(), cam())
        mid_res = nn.Sequential(pool(), cam())
        high_res = nn.Sequential(cam())
        top = nn.Sequential(nn.Conv2d(self.inplanes, out_chans, kernel_size=3, stride=1, dilation=1,
                                      padding=1, bias=True))
        return nn.ModuleList([low_res, mid_res, high_res, top])

    def _attention_body(self):
        """students.py:653-683."""
        return self._body([1, 2, 3, 4, 5], 1)

    def _detection_body_v1(self):
        """students.py:685-713."""
        return self._body([1, 2, 3, 4], self.num_heatmaps + self.ae_dims)

    def load_state_dicts(self, inpath):
        """students.py:715-731 (same five files)."""
        for name in ("mid_stem", "att_lo", "att_mid", "att_hi", "att_top"):
            getattr(self, name).load_state_dict(
                torch.load(inpath + name + ".statedict", map_location=self.device))
        self.invalidate_plans()

    # ------------------------------------------------------------------ plan
    def _ref_param(self):
        return self.mid_stem[0].weight

    def _mode(self):
        # the reference keeps everything but the stem in fp32; here the precision mode is one
        # switch for the whole student
        return "bf16" if self.half_precision else "fp32"

    def _cam(self, R, cam, x, cin_index=None, padded=False):
        """ContextAwareModule.forward (students.py:181-201) on virtual tensor x.
        ``padded``: x stores the module's channels first and ZERO pad channels up to ``x.ld``
        (widths that are not multiples of 16: the convs then read all ``x.ld`` channels with zero
        weights for the pads, and the output's pad channels are written as zeros, too);
        ``cin_index``: x stores the module's channels in padded slots (see ``_Recorder.conv``)."""
        c = cam.residual[0].in_channels
        kw = {}
        cs = c                                           # channels of x the module reads
        if cin_index is not None:
            kw = {"cin_index": cin_index}
            cs = len(cin_index)
        elif padded:
            kw = {"cin_store": x.ld}
            cs = x.ld
        c16 = (c + 15) // 16 * 16
        wide = cin_index is not None or padded
        res = R.conv(x, cam.residual[0], cam.residual[1], True, **kw) if not wide else \
            R.conv(x, cam.residual[0], cam.residual[1], True, cout_store=c16, pad_cout=True, **kw)
        # squeeze-excitation gate: deterministic two-stage mean, then the two tiny FCs
        hw = x.h * x.w
        chunks = max(1, min(64, hw // 256))
        fc1, fc2 = cam.se.fc[0], cam.se.fc[2]
        hid = fc1.out_features
        dev = R.device
        w1 = fc1.weight.detach().float()                 # (hid, c) -> columns in x's stored order
        if cs != c:
            idx = torch.as_tensor(list(cin_index) if cin_index is not None
                                  else list(range(c)) + [-1] * (cs - c), dtype=torch.long,
                                  device=w1.device)
            w1 = w1[:, idx.clamp(min=0)] * (idx >= 0).to(w1.dtype).view(1, -1)
        packed = torch.cat([w1.reshape(-1), fc1.bias.detach().float(),
                            fc2.weight.detach().float().reshape(-1), fc2.bias.detach().float()]
                           ).to(dev).contiguous()
        partial = torch.empty((x.n, chunks, cs), dtype=torch.float32, device=dev)
        gate = torch.empty((x.n, c), dtype=torch.float32, device=dev)
        R.keepalive += [packed, partial, gate]
        R.aux(AUX_SE_PARTIAL, [x], [], x, None, None, partial, [R.dt, x.n, hw, cs, x.ld, chunks])
        R.aux(AUX_SE_GATE, [], [], partial, packed, None, gate, [R.dt, x.n, c, hid, chunks, hw, cs])
        # hybrid dilated convolutions write channel slices of one buffer (the torch.cat)
        # (every branch gets a 16-channel slot -- zero pad channels, zero weights for them in
        # hdc_top -- so that the slices are chunk aligned and the branches, dilated ones included,
        # run on the tcgen05 per-tap engine instead of the CUDA-core fallback)
        hc = cam.hdcs[0][0].out_channels
        nd = len(cam.hdcs)
        slot = (hc + 15) // 16 * 16
        cat = R.new(x.n, x.h, x.w, slot * nd)
        for i, hdc in enumerate(cam.hdcs):
            R.conv(x, hdc[0], hdc[1], True, out=cat, out_coff=slot * i, cout_store=slot, pad_cout=True,
                   **kw)
        index = [(i * hc + j if j < hc else -1) for i in range(nd) for j in range(slot)]
        top = R.conv(cat, cam.hdc_top[0], cam.hdc_top[1], True, cin_index=index)
        out = R.new(x.n, x.h, x.w, c16)
        R.aux(AUX_CAM_MIX, [res, top], [out], res, top, gate, out,
              [R.dt, x.n, hw, c, res.ld, top.ld, out.ld, c16 if wide else 0])
        return out

    def _pool(self, R, x, c):
        out = R.new(x.n, x.h // 2, x.w // 2, x.ld)
        R.aux(AUX_AVGPOOL, [x], [out], x, None, None, out, [R.dt, x.n, x.h, x.w, c, x.ld, out.ld])
        return out

    def _record(self, n, h, w, mode, device, in_is_half, out_half):
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        stem = self.stem[1]
        if mode == "bf16" and self.conv_engine != L.ENGINE_FFMA:
            x = R.stem_tc(stem.conv1, stem.bn1)
        else:
            x = R.stem(stem.conv1, stem.bn1)
        x = R.conv(x, stem.conv2, stem.bn2, True)
        for blk in stem.layer1:
            res = x
            if blk.downsample is not None:
                res = R.conv(x, blk.downsample[0], blk.downsample[1], False)
            t = R.conv(x, blk.conv1, blk.bn1, True)
            t = R.conv(t, blk.conv2, blk.bn2, True)
            x = R.conv(t, blk.conv3, blk.bn3, True, residual=res)
        # (mid-stem width 152 is stored as 160 channels: zero pad channels / zero weights keep the
        # second conv on the tcgen05 halo engine)
        mid = self.mid_stem[0].out_channels
        mid16 = (mid + 15) // 16 * 16
        x = R.conv(x, self.mid_stem[0], self.mid_stem[1], True, cout_store=mid16, pad_cout=True)
        s = R.conv(x, self.mid_stem[3], self.mid_stem[4], True, cin_store=mid16)
        c = self.inplanes

        # attention pyramid (students.py:741-753): att = hi + 2 * nearest_x4(lo)
        hi = self._cam(R, self.att_hi[0], s)
        mid = self._cam(R, self.att_mid[1], self._pool(R, s, c))
        lo = self._cam(R, self.att_lo[1], self._pool(R, mid, c))
        att_sum = R.fuse([hi, lo, lo], [0, 2, 2], c, False)
        att_logit = R.conv(att_sum, self.att_top[0], None, False)
        att_out = torch.empty((n, 1, s.h, s.w), dtype=torch.float32, device=device)
        s2 = R.new(n, s.h, s.w, s.ld)
        R.aux(AUX_ATT_ADD, [att_logit, s], [s2], att_logit, s, att_out, s2,
              [R.dt, n, s.h * s.w, c, att_logit.ld, s.ld, s2.ld])

        # detection pyramid (students.py:756-766): det_hi is applied twice to the same input and
        # det_mid is never used, so mid == hi and det = hi + 2 * nearest_x2(det_lo(pool(hi)))
        dhi = self._cam(R, self.det_hi[0], s2)
        dlo = self._cam(R, self.det_lo[1], self._pool(R, dhi, c))
        det_sum = R.fuse([dhi, dlo, dlo], [0, 1, 1], c, False)
        det = R.conv(det_sum, self.det_top[0], None, False)
        cout = self.det_top[0].out_channels
        det_out = torch.empty((n, cout, s.h, s.w), dtype=torch.float32, device=device)
        R.to_nchw(det, cout, 0, det_out)
        return R, [att_out, det_out]

    def forward(self, x, out_hw=None, return_intermediate=False):
        """students.py:733-768 -> (att, det)."""
        att, det = self._run_plans(x, 16)
        return att, det


class CamStudent(_PlanRunner, nn.Module):
    """Same constructor, attributes and ``forward(x, out_hw=None) -> [pred]`` as
    ``rtpe.students.CamStudent`` (students.py:502-592): HigherHRNet stem -> one 3x3 mid-stem conv ->
    ``num_stages`` context-aware modules (dilations 1, 2, 3, 5, 8, 12) that all read the mid-stem
    output and are summed -> the LAST 3x3 head (the reference builds one head per stage and uses
    only ``hm_convs[-1]``, :581) -> optional bilinear ``align_corners=True`` resize to ``out_hw``."""

    def __init__(self, hhrnet_statedict_path=None, device="cuda", inplanes=48, num_stages=3,
                 num_heatmaps=17, ae_dims=1, half_precision=True,
                 init_fn=torch.nn.init.kaiming_normal_, trainable_stem=False, bn_momentum=0.1):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.num_stages = num_stages
        self.num_heatmaps = num_heatmaps
        self.ae_dims = ae_dims
        self.stem = StemHRNet()
        self.stem_out_chans = self.stem.layer1[-1].bn3.num_features
        self.trainable_stem = trainable_stem
        self.inplanes = inplanes
        self.mid_stem = nn.Sequential(
            nn.Conv2d(self.stem_out_chans, inplanes, kernel_size=3, stride=1, dilation=1, padding=1,
                      bias=False),
            nn.BatchNorm2d(inplanes, momentum=bn_momentum), nn.ReLU(inplace=True))
        self.cams, self.hm_convs = self._make_body()
        if init_fn is not None:
            self.apply(lambda module: init_weights(module, init_fn, 0.0))
        self.half_precision = bool(half_precision)
        if half_precision:
            self.stem = network_to_half(self.stem)
        else:
            self.stem = nn.Sequential(nn.Identity(), self.stem)
        if hhrnet_statedict_path is not None:
            self.stem[1].load_pretrained(hhrnet_statedict_path, device, check=False)
        self._init_runner()
        self.to(device)
        self.device = device

    def _make_body(self):
        """students.py:555-566."""
        hm_out_ch = self.num_heatmaps + self.ae_dims
        cams, hms = nn.ModuleList(), nn.ModuleList()
        for _ in range(self.num_stages):
            cams.append(ContextAwareModule(self.inplanes, hdc_dilations=[1, 2, 3, 5, 8, 12]))
            hms.append(nn.Conv2d(self.inplanes, hm_out_ch, kernel_size=3, padding=1, bias=True))
        return cams, hms

    def _ref_param(self):
        return self.mid_stem[0].weight

    def _mode(self):
        return "bf16" if self.half_precision else "fp32"

    _cam = AttentionStudent._cam

    def _record(self, n, h, w, mode, device, in_is_half, out_half):
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        stem = self.stem[1]
        if mode == "bf16" and self.conv_engine != L.ENGINE_FFMA:
            x = R.stem_tc(stem.conv1, stem.bn1)
        else:
            x = R.stem(stem.conv1, stem.bn1)
        x = R.conv(x, stem.conv2, stem.bn2, True)
        for blk in stem.layer1:
            res = x
            if blk.downsample is not None:
                res = R.conv(x, blk.downsample[0], blk.downsample[1], False)
            t = R.conv(x, blk.conv1, blk.bn1, True)
            t = R.conv(t, blk.conv2, blk.bn2, True)
            x = R.conv(t, blk.conv3, blk.bn3, True, residual=res)
        s = R.conv(x, self.mid_stem[0], self.mid_stem[1], True)
        c = self.inplanes
        outs = [self._cam(R, cam, s) for cam in self.cams]          # x = cam0(s) + cam1(s) + ...
        acc = outs[0]
        for k in range(1, len(outs), 3):                            # the fuse kernel sums up to 4 terms
            terms = [acc] + outs[k:k + 3]
            acc = R.fuse(terms, [0] * len(terms), c, False)
        head = self.hm_convs[-1]
        y = R.conv(acc, head, None, False)
        cout = head.out_channels
        out = torch.empty((n, cout, s.h, s.w), dtype=torch.float32, device=device)
        R.to_nchw(y, cout, 0, out)
        return R, [out]

    def forward(self, x, out_hw=None, return_intermediate=False):
        """students.py:568-592 -> [pred (N, num_heatmaps + ae_dims, H/4, W/4 or out_hw)]."""
        if return_intermediate:
            raise NotImplementedError                       # as the reference (:579)
        (pred,) = self._run_plans(x, 4)
        if out_hw is not None:
            from .inference import bilinear_resize
            pred = bilinear_resize(pred, out_hw, True)
        return [pred]


class SkipConv(nn.Module):
    """Parameter container of students.py:37-90: a chain of conv + BN + ReLU with a (1x1 conv + BN)
    skip, ``relu(chain(x) + downsample(x))``."""

    def __init__(self, in_chans, out_chans, ksizes, strides=None, dilations=None, paddings=None,
                 downsample=None, bn_momentum=0.1):
        super().__init__()
        if strides is None:
            strides = [1 for _ in in_chans]
        if dilations is None:
            dilations = [1 for _ in in_chans]
        if paddings is None:
            paddings = [0 for _ in in_chans]
        assert len(in_chans) == len(out_chans) == len(ksizes) == len(strides) == len(dilations) == \
            len(paddings), "Channels, ksizes, strides and dilations must be of same length!"
        self.convs = nn.ModuleList(
            [nn.Conv2d(i, o, kernel_size=k, stride=st, dilation=d, padding=p, bias=False)
             for (i, o, k, st, d, p) in zip(in_chans, out_chans, ksizes, strides, dilations, paddings)])
        self.bns = nn.ModuleList([nn.BatchNorm2d(o, momentum=bn_momentum) for o in out_chans])
        self.relus = nn.ModuleList([nn.ReLU(inplace=True) for _ in out_chans])
        self.downsample = downsample
        self.final_relu = nn.ReLU(inplace=True)


def get_straight_skip_conv(in_chans, out_chans, bn_momentum=0.1):
    """students.py:93-112."""
    assert len(in_chans) == len(out_chans), "in_chans and out_chans must have same length!"
    n = len(in_chans)
    downsample = nn.Sequential(
        nn.Conv2d(in_chans[0], out_chans[-1], kernel_size=1, stride=1, padding=0, bias=False),
        nn.BatchNorm2d(out_chans[-1], momentum=bn_momentum))
    return SkipConv(in_chans, out_chans, [3] * n, [1] * n, [1] * n, [1] * n, downsample, bn_momentum)


def _record_stem(R, stem, mode, conv_engine, last_out=None):
    """StemHRNet.forward (students.py:252-264) on the recorder.  ``last_out``: virtual tensor whose
    first channels receive the stem output (a concat buffer) instead of a new tensor."""
    if mode == "bf16" and conv_engine != L.ENGINE_FFMA:
        x = R.stem_tc(stem.conv1, stem.bn1)
    else:
        x = R.stem(stem.conv1, stem.bn1)
    x = R.conv(x, stem.conv2, stem.bn2, True)
    for k, blk in enumerate(stem.layer1):
        res = x
        if blk.downsample is not None:
            res = R.conv(x, blk.downsample[0], blk.downsample[1], False)
        t = R.conv(x, blk.conv1, blk.bn1, True)
        t = R.conv(t, blk.conv2, blk.bn2, True)
        dst = last_out if k == len(stem.layer1) - 1 else None
        x = R.conv(t, blk.conv3, blk.bn3, True, residual=res, out=dst)
    return x


def _record_skip_conv(R, skc, x):
    """SkipConv.forward (students.py:73-90): the last ReLU of the chain comes BEFORE the skip is
    added, so the sum is a separate fuse op (relu(relu(c) + r) != relu(c + r))."""
    def c16(conv):
        return (conv.out_channels + 15) // 16 * 16
    res = R.conv(x, skc.downsample[0], skc.downsample[1], False, cout_store=c16(skc.downsample[0]),
                 pad_cout=True)
    for conv, bn in zip(skc.convs, skc.bns):
        if conv.kernel_size[0] not in (1, 3) or conv.stride[0] != 1:
            raise NotImplementedError("SkipConv with kernel %s / stride %s" % (conv.kernel_size, conv.stride))
        x = R.conv(x, conv, bn, True, cout_store=c16(conv), pad_cout=True)
    return R.fuse([x, res], [0, 0], x.ld, True)          # pad channels are zeros on both sides


class RefinerStudent(_PlanRunner, nn.Module):
    """Same constructor, attributes and ``forward(x, out_hw=None) -> pred`` as
    ``rtpe.students.RefinerStudent`` (students.py:302-386): HigherHRNet stem, then ``SkipConv`` stages
    on the 256-channel stem output, ``x = stage_k(stem_out + x)``; the last stage ends in
    ``num_heatmaps + ae_dims`` channels."""

    def __init__(self, hhrnet_statedict_path=None, device="cuda", layers_per_stage=[3, 3, 3],
                 num_heatmaps=17, ae_dims=1, half_precision=True,
                 init_fn=torch.nn.init.kaiming_normal_, trainable_stem=False, bn_momentum=0.1):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.layers_per_stage = layers_per_stage
        self.num_heatmaps = num_heatmaps
        self.ae_dims = ae_dims
        self.stem = StemHRNet()
        self.stem_out_chans = self.stem.layer1[-1].bn3.num_features
        self.trainable_stem = trainable_stem
        self.stages = self._make_body()
        if init_fn is not None:
            self.apply(lambda module: init_weights(module, init_fn, 0.0))
        self.half_precision = bool(half_precision)
        if half_precision:
            self.stem = network_to_half(self.stem)
        else:
            self.stem = nn.Sequential(nn.Identity(), self.stem)
        if hhrnet_statedict_path is not None:
            self.stem[1].load_pretrained(hhrnet_statedict_path, device, check=False)
        self._init_runner()
        self.to(device)
        self.device = device

    def save_body(self, out_path):
        torch.save(self.stages.state_dict(), out_path)

    def load_body(self, statedict_path):
        self.stages.load_state_dict(torch.load(statedict_path))
        self.invalidate_plans()

    def _make_body(self):
        """students.py:351-372."""
        stages = nn.ModuleList()
        ch = self.stem_out_chans
        for n in self.layers_per_stage[:-1]:
            stages.append(get_straight_skip_conv([ch] * n, [ch] * n, self.bn_momentum))
        n = self.layers_per_stage[-1]
        out_chans = [ch] * n
        out_chans[-1] = self.num_heatmaps + self.ae_dims
        stages.append(get_straight_skip_conv([ch] * n, out_chans, self.bn_momentum))
        return stages

    def _ref_param(self):
        return self.stages[0].convs[0].weight

    def _mode(self):
        return "bf16" if self.half_precision else "fp32"

    def _record(self, n, h, w, mode, device, in_is_half, out_half):
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        s = _record_stem(R, self.stem[1], mode, self.conv_engine)
        x = _record_skip_conv(R, self.stages[0], s)
        for st in self.stages[1:]:
            x = _record_skip_conv(R, st, R.fuse([s, x], [0, 0], s.ld, False))
        cout = self.stages[-1].convs[-1].out_channels
        out = torch.empty((n, cout, s.h, s.w), dtype=torch.float32, device=device)
        R.to_nchw(x, cout, 0, out)
        return R, [out]

    def forward(self, x, out_hw=None):
        """students.py:374-386 -> pred (N, num_heatmaps + ae_dims, H/4, W/4 or out_hw)."""
        (pred,) = self._run_plans(x, 4)
        if out_hw is not None:
            from .inference import bilinear_resize
            pred = bilinear_resize(pred, out_hw, True)
        return pred


def _r16(c):
    return (c + 15) // 16 * 16


class MultistageStudent(_PlanRunner, nn.Module):
    """Same constructor, attributes and ``forward(x, out_hw=None) -> [stage outputs]`` as
    ``rtpe.students.MultistageStudent`` (students.py:389-499): HigherHRNet stem, then ``SkipConv``
    stages with intermediate supervision -- stage 0 reads the 256-channel stem output, every later
    stage reads ``cat(stem_out, previous stage output)``; each stage ends in ``num_heatmaps +
    ae_dims`` channels and every stage output is returned.  With ``out_hw`` the stem output is
    resized (bilinear, ``align_corners=True``) first and the stages run at that resolution, as in
    the reference (:481-498; its later resizes are then identities).

    The ``torch.cat`` is never built: stem output and stage outputs live in one NHWC buffer
    ``[stem 256 | slot A | slot B]``; stage k reads the whole buffer (zero weights for the slot it
    does not use) and its last launch -- the 1x1 skip conv with the chain output as residual,
    ``relu(relu(chain) + skip)`` -- writes the other slot."""

    REMARKS = "Second attempt. We added intermediate supervision"

    def __init__(self, hhrnet_statedict_path=None, device="cuda", layers_per_stage=[3, 3, 3],
                 num_heatmaps=17, ae_dims=1, half_precision=True,
                 init_fn=torch.nn.init.kaiming_normal_, trainable_stem=False, bn_momentum=0.1):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.layers_per_stage = layers_per_stage
        self.num_heatmaps = num_heatmaps
        self.ae_dims = ae_dims
        self.stem = StemHRNet()
        self.stem_out_chans = self.stem.layer1[-1].bn3.num_features
        self.trainable_stem = trainable_stem
        self.stages = self._make_body()
        if init_fn is not None:
            self.apply(lambda module: init_weights(module, init_fn, 0.0))
        self.half_precision = bool(half_precision)
        if half_precision:
            self.stem = network_to_half(self.stem)
        else:
            self.stem = nn.Sequential(nn.Identity(), self.stem)
        if hhrnet_statedict_path is not None:
            self.stem[1].load_pretrained(hhrnet_statedict_path, device, check=False)
        self._init_runner()
        self._out_hw = None
        self.to(device)
        self.device = device

    def save_body(self, out_path):
        torch.save(self.stages.state_dict(), out_path)

    def load_body(self, statedict_path):
        self.stages.load_state_dict(torch.load(statedict_path))
        self.invalidate_plans()

    def _make_body(self):
        """students.py:440-471."""
        stages = nn.ModuleList()
        stem_ch = self.stem_out_chans
        out_ch = self.num_heatmaps + self.ae_dims
        for stage_i, l in enumerate(self.layers_per_stage):
            in_chans = [out_ch + stem_ch for _ in range(l)]
            out_chans = [out_ch + stem_ch for _ in range(l)]
            if stage_i == 0:
                in_chans[0] = stem_ch
            out_chans[-1] = out_ch
            downsample = nn.Sequential(
                nn.Conv2d(in_chans[0], out_chans[-1], kernel_size=1, stride=1, padding=0, bias=False),
                nn.BatchNorm2d(out_chans[-1], momentum=self.bn_momentum))
            stages.append(SkipConv(in_chans, out_chans, [3] * l, [1] * l, [1] * l, [1] * l, downsample,
                                   self.bn_momentum))
        return stages

    def _ref_param(self):
        return self.stages[0].convs[0].weight

    def _mode(self):
        return "bf16" if self.half_precision else "fp32"

    def _plan_extra_key(self):
        return self._out_hw

    def _record(self, n, h, w, mode, device, in_is_half, out_half):
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        sc = self.stem_out_chans
        oc = self.num_heatmaps + self.ae_dims
        slot = _r16(oc)
        if self._out_hw is None:
            buf = R.new(n, h // 4, w // 4, sc + 2 * slot)
            _record_stem(R, self.stem[1], mode, self.conv_engine, last_out=buf)
        else:
            s = _record_stem(R, self.stem[1], mode, self.conv_engine)
            hh, ww = self._out_hw
            buf = R.new(n, hh, ww, sc + 2 * slot)
            R.aux(AUX_RESIZE_NHWC, [s], [buf], s, None, None, buf,
                  [R.dt, n, s.h, s.w, sc, s.ld, buf.ld, hh, ww, 1, 0])
        outs = []
        for k, st in enumerate(self.stages):
            cur = sc + slot * (k % 2)                  # channel offset of the slot this stage writes
            prev = sc + slot * ((k + 1) % 2)
            if k == 0:
                first = {"cin_store": sc}              # stage 0 reads the stem channels only
            else:
                first = {"cin_index": list(range(sc)) + [
                    (sc + j - prev if prev <= j < prev + oc else -1) for j in range(sc, sc + 2 * slot)]}
            # stage 0 zero-fills both slots (the other one is read with zero weights by stage 1)
            wide = 2 * slot if k == 0 else slot
            x = buf
            kw = first
            for i, (conv, bn) in enumerate(zip(st.convs, st.bns)):
                if conv.kernel_size[0] not in (1, 3) or conv.stride[0] != 1:
                    raise NotImplementedError("SkipConv with kernel %s / stride %s"
                                              % (conv.kernel_size, conv.stride))
                last = i == len(st.convs) - 1
                x = R.conv(x, conv, bn, True, cout_store=wide if last else _r16(conv.out_channels),
                           pad_cout=True, **kw)
                kw = {"cin_store": x.ld}
            R.conv(buf, st.downsample[0], st.downsample[1], True, residual=x, out=buf, out_coff=cur,
                   cout_store=wide, pad_cout=True, **first)
            out = torch.empty((n, oc, buf.h, buf.w), dtype=torch.float32, device=device)
            R.to_nchw(buf, oc, cur, out)
            outs.append(out)
        return R, outs

    def forward(self, x, out_hw=None):
        """students.py:473-499 -> list of the stage outputs (N, num_heatmaps + ae_dims, H/4, W/4 or
        out_hw), first stage first."""
        self._out_hw = None if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
        return list(self._run_plans(x, 4))


def _record_conv5s2(R, xs, conv, bn, relu, bs, **kw):
    """5x5 / stride 2 / padding 2 conv (+BN+ReLU) as a 3x3 / stride 1 conv over the space-to-depth
    tensor ``xs``: block (ry*2+rx) of ``bs`` stored channels holds the input pixels (2y+ry, 2x+rx);
    input offset d in [-2, 2] = 2*q + r with q in [-1, 1], r in {0, 1} (d = 3 gets zero weights)."""
    w = conv.weight.detach().float()                                     # (Cout, Cin, 5, 5)
    cout, cin = w.shape[:2]
    assert tuple(w.shape[2:]) == (5, 5) and conv.stride[0] == 2 and conv.padding[0] == 2 and bs >= cin
    w3 = w.new_zeros((cout, 4 * bs, 3, 3))
    for ry in (0, 1):
        for rx in (0, 1):
            b0 = (ry * 2 + rx) * bs
            for qy in (-1, 0, 1):
                for qx in (-1, 0, 1):
                    dy, dx = 2 * qy + ry, 2 * qx + rx
                    if dy <= 2 and dx <= 2:
                        w3[:, b0:b0 + cin, qy + 1, qx + 1] = w[:, :, dy + 2, dx + 2]
    tmp = nn.Conv2d(4 * bs, cout, kernel_size=3, padding=1, bias=conv.bias is not None)
    tmp.weight = nn.Parameter(w3, requires_grad=False)
    if conv.bias is not None:
        tmp.bias = nn.Parameter(conv.bias.detach().float().clone(), requires_grad=False)
    return R.conv(xs, tmp, bn, relu, cin_store=xs.ld, pad_cout=True, **kw)


class AttentionStudentSteps(_PlanRunner, nn.Module):
    """Same constructor, attributes, state-dict layout and ``forward(x, out_hw=None, alt=None,
    att_divisor=None) -> (att, det)`` as ``rtpe.students.AttentionStudentSteps``
    (students.py:786-1073; the student of ``eval_attention.py`` / ``distillation.py``):

    stem -> mid stem (``inplanes`` channels) -> ``cat`` with the bilinearly down-sized ``alt`` image
    -> attention pyramid of context-aware modules -> ``att = sigmoid(att_top(hi + 2 * up4(lo)) /
    att_divisor)`` -> ``cat(stem_cat * att, alt_img_stem(alt))`` -> three context-aware modules
    (dilations 1, 2, 3) and a 3x3 head.

    Neither ``torch.cat`` is built (the parts are written into channel slices of one NHWC buffer,
    odd widths live in zero-padded 16-channel slots), and the two 5x5 / stride-2 convs of
    ``alt_img_stem`` run as 3x3 / stride-1 convs on the 2x2 pixel-parity (space-to-depth) planes,
    so every conv stays on the plan's conv engine."""

    def __init__(self, hhrnet_statedict_path=None, device="cuda", inplanes=48, num_heatmaps=17,
                 ae_dims=1, half_precision=True, init_fn=torch.nn.init.kaiming_normal_,
                 trainable_stem=False, bn_momentum=0.1):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.num_heatmaps = num_heatmaps
        self.ae_dims = ae_dims
        self.stem = StemHRNet()
        self.stem_out_chans = self.stem.layer1[-1].bn3.num_features
        self.trainable_stem = trainable_stem
        self.inplanes = inplanes
        mid_inplanes = (self.stem_out_chans + self.inplanes) // 2
        self.mid_stem = nn.Sequential(
            nn.Conv2d(self.stem_out_chans, mid_inplanes, kernel_size=3, stride=1, dilation=1, padding=1,
                      bias=False),
            nn.BatchNorm2d(mid_inplanes, momentum=bn_momentum), nn.ReLU(inplace=True),
            nn.Conv2d(mid_inplanes, inplanes, kernel_size=3, stride=1, dilation=1, padding=1, bias=False),
            nn.BatchNorm2d(inplanes, momentum=bn_momentum), nn.ReLU(inplace=True))
        self._alt_planes = 50
        self.alt_img_stem = nn.Sequential(
            nn.Conv2d(3, self._alt_planes, kernel_size=5, stride=2, dilation=1, padding=2, bias=False),
            nn.BatchNorm2d(self._alt_planes, momentum=bn_momentum), nn.ReLU(inplace=True),
            nn.Conv2d(self._alt_planes, inplanes, kernel_size=5, stride=2, dilation=1, padding=2,
                      bias=False),
            nn.BatchNorm2d(inplanes, momentum=bn_momentum), nn.ReLU(inplace=True))
        self.att_lo, self.att_mid, self.att_hi, self.att_top = self._attention_body()
        self.steps = self._detection_stage()
        if init_fn is not None:
            self.apply(lambda module: init_weights(module, init_fn, 0.0))
        self.half_precision = bool(half_precision)
        if half_precision:
            self.stem = network_to_half(self.stem)
        else:
            self.stem = nn.Sequential(nn.Identity(), self.stem)
        if hhrnet_statedict_path is not None:
            self.stem[1].load_pretrained(hhrnet_statedict_path, device, check=False)
        self._init_runner()
        self._att_divisor = None
        self.to(device)
        self.device = device

    def _attention_body(self):
        """students.py:868-895."""
        c = self.inplanes + 3

        def pool():
            return nn.AvgPool2d(kernel_size=3, stride=2, padding=1, count_include_pad=False)
        low_res = nn.Sequential(pool(), ContextAwareModule(c, hdc_dilations=[1, 2, 3, 4]))
        mid_res = nn.Sequential(pool(), ContextAwareModule(c, hdc_dilations=[1, 2, 3, 4]))
        high_res = nn.Sequential(ContextAwareModule(c, hdc_dilations=[1, 2, 3, 4]))
        top = nn.Sequential(nn.Conv2d(c, 1, kernel_size=3, stride=1, dilation=1, padding=1, bias=True))
        return nn.ModuleList([low_res, mid_res, high_res, top])

    def _detection_stage(self):
        """students.py:897-948."""
        c = 2 * self.inplanes + 3
        return nn.Sequential(
            ContextAwareModule(c, hdc_dilations=[1, 2, 3]),
            ContextAwareModule(c, hdc_dilations=[1, 2, 3]),
            ContextAwareModule(c, hdc_dilations=[1, 2, 3]),
            nn.Conv2d(c, self.num_heatmaps + self.ae_dims, kernel_size=3, stride=1, dilation=1,
                      padding=1, bias=True))

    def load_state_dicts(self, inpath):
        """students.py:950-964."""
        for name in ("mid_stem", "att_lo", "att_mid", "att_hi", "att_top"):
            getattr(self, name).load_state_dict(
                torch.load(inpath + name + ".statedict", map_location=self.device))
        self.invalidate_plans()

    def _ref_param(self):
        return self.mid_stem[0].weight

    def _mode(self):
        return "bf16" if self.half_precision else "fp32"

    def _plan_extra_key(self):
        return self._att_divisor

    _cam = AttentionStudent._cam
    _pool = AttentionStudent._pool

    def _record(self, n, h, w, mode, device, in_is_half, out_half):
        import struct
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        P = self.inplanes
        p16, c1 = _r16(P), P + 3
        ld1 = _r16(c1)
        alt = torch.zeros((n, 3, h, w), dtype=torch.float32, device=device)
        R.extra_input = alt
        R.keepalive.append(alt)
        s0 = _record_stem(R, self.stem[1], mode, self.conv_engine)
        mid16 = _r16(self.mid_stem[0].out_channels)
        t = R.conv(s0, self.mid_stem[0], self.mid_stem[1], True, cout_store=mid16, pad_cout=True)
        # cat(stem_out, alt down-sized): the mid stem writes channels [0, P), the image op
        # channels [P, P + 3) and zeros up to the slot end
        sc = R.new(n, s0.h, s0.w, ld1)
        R.conv(t, self.mid_stem[3], self.mid_stem[4], True, cin_store=mid16, out=sc, cout_store=p16,
               pad_cout=True)
        R.aux(AUX_IMAGE_NHWC, [], [sc], alt, None, None, sc,
              [R.dt, n, h, w, 3, 0, sc.ld, sc.h, sc.w, P, ld1 - P, 0])
        # attention pyramid (students.py:990-1003): mid and lo are both the up-sampled lo
        hi = self._cam(R, self.att_hi[0], sc, padded=True)
        mid = self._cam(R, self.att_mid[1], self._pool(R, sc, sc.ld), padded=True)
        lo = self._cam(R, self.att_lo[1], self._pool(R, mid, mid.ld), padded=True)
        att_sum = R.fuse([hi, lo, lo], [0, 2, 2], ld1, False)
        att_logit = R.conv(att_sum, self.att_top[0], None, False, cin_store=att_sum.ld)
        att_out = torch.empty((n, 1, sc.h, sc.w), dtype=torch.float32, device=device)
        # cat(stem_cat * att, alt_img_stem(alt))  (students.py:1018-1020)
        cat2 = R.new(n, sc.h, sc.w, ld1 + p16)
        div = 1.0 if self._att_divisor is None else float(self._att_divisor)
        (div_bits,) = struct.unpack("<i", struct.pack("<f", div))
        R.aux(AUX_ATT_MUL, [att_logit, sc], [cat2], att_logit, sc, att_out, cat2,
              [R.dt, n, sc.h * sc.w, ld1, att_logit.ld, sc.ld, cat2.ld, div_bits])
        xs = R.new(n, h // 2, w // 2, 16)
        R.aux(AUX_IMAGE_NHWC, [], [xs], alt, None, None, xs,
              [R.dt, n, h, w, 3, 0, xs.ld, xs.h, xs.w, 0, 16, 1])
        a16 = _r16(self.alt_img_stem[0].out_channels)
        a1 = _record_conv5s2(R, xs, self.alt_img_stem[0], self.alt_img_stem[1], True, 3, cout_store=a16)
        xs2 = R.new(n, h // 4, w // 4, 4 * a16)
        R.aux(AUX_S2D, [a1], [xs2], a1, None, None, xs2, [R.dt, n, a1.h, a1.w, a16, a1.ld, xs2.ld])
        _record_conv5s2(R, xs2, self.alt_img_stem[3], self.alt_img_stem[4], True, a16, out=cat2,
                        out_coff=ld1, cout_store=p16)
        # detection steps (students.py:1026): Sequential of context-aware modules + 3x3 head
        index = [(j if j < c1 else -1) for j in range(ld1)] + \
                [(c1 + k if k < P else -1) for k in range(p16)]
        y, first = cat2, True
        for m in self.steps:
            if isinstance(m, ContextAwareModule):
                y = self._cam(R, m, y, cin_index=index) if first else self._cam(R, m, y, padded=True)
            else:
                y = R.conv(y, m, None, False, cin_index=index) if first else \
                    R.conv(y, m, None, False, cin_store=y.ld)
            first = False
        cout = self.steps[-1].out_channels
        det_out = torch.empty((n, cout, sc.h, sc.w), dtype=torch.float32, device=device)
        R.to_nchw(y, cout, 0, det_out)
        return R, [att_out, det_out]

    def forward(self, x, out_hw=None, alt=None, att_divisor=None):
        """students.py:966-1052 -> (att (N,1,H/4,W/4) after the sigmoid, det (N, num_heatmaps +
        ae_dims, H/4, W/4)); ``out_hw`` is ignored as in the reference."""
        if alt is None:
            raise NotImplementedError("ATM alt is expected")             # students.py:983
        if not isinstance(alt, torch.Tensor) or alt.shape != x.shape:
            raise ValueError("alt must be an image batch of x's shape (N, 3, H, W)")
        self._att_divisor = None if att_divisor is None else float(att_divisor)
        att, det = self._run_plans(x, 16, extra=alt.to(device=x.device, dtype=torch.float32))
        return att, det
