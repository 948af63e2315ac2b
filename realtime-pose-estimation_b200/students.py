"""Context-aware-module student -- drop-in for ``rtpe.students.AttentionStudent``
(rtpe/students.py:595-771) and its building blocks ``SELayer`` (:118-142),
``ContextAwareModule`` (:145-201) and ``StemHRNet`` (:206-282), executed by libbrtpe.so.

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

    def _cam(self, R, cam, x):
        """ContextAwareModule.forward (students.py:181-201) on virtual tensor x."""
        c = cam.residual[0].in_channels
        res = R.conv(x, cam.residual[0], cam.residual[1], True)
        # squeeze-excitation gate: deterministic two-stage mean, then the two tiny FCs
        hw = x.h * x.w
        chunks = max(1, min(64, hw // 256))
        fc1, fc2 = cam.se.fc[0], cam.se.fc[2]
        hid = fc1.out_features
        dev = R.device
        packed = torch.cat([fc1.weight.detach().float().reshape(-1), fc1.bias.detach().float(),
                            fc2.weight.detach().float().reshape(-1), fc2.bias.detach().float()]
                           ).to(dev).contiguous()
        partial = torch.empty((x.n, chunks, c), dtype=torch.float32, device=dev)
        gate = torch.empty((x.n, c), dtype=torch.float32, device=dev)
        R.keepalive += [packed, partial, gate]
        R.aux(AUX_SE_PARTIAL, [x], [], x, None, None, partial, [R.dt, x.n, hw, c, x.ld, chunks])
        R.aux(AUX_SE_GATE, [], [], partial, packed, None, gate, [R.dt, x.n, c, hid, chunks, hw])
        # hybrid dilated convolutions write channel slices of one buffer (the torch.cat)
        # (every branch gets a 16-channel slot -- zero pad channels, zero weights for them in
        # hdc_top -- so that the slices are chunk aligned and the branches, dilated ones included,
        # run on the tcgen05 per-tap engine instead of the CUDA-core fallback)
        hc = cam.hdcs[0][0].out_channels
        nd = len(cam.hdcs)
        slot = (hc + 15) // 16 * 16
        cat = R.new(x.n, x.h, x.w, slot * nd)
        for i, hdc in enumerate(cam.hdcs):
            R.conv(x, hdc[0], hdc[1], True, out=cat, out_coff=slot * i, cout_store=slot, pad_cout=True)
        index = [(i * hc + j if j < hc else -1) for i in range(nd) for j in range(slot)]
        top = R.conv(cat, cam.hdc_top[0], cam.hdc_top[1], True, cin_index=index)
        out = R.new(x.n, x.h, x.w, (c + 15) // 16 * 16)
        R.aux(AUX_CAM_MIX, [res, top], [out], res, top, gate, out,
              [R.dt, x.n, hw, c, res.ld, top.ld, out.ld])
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


def _record_stem(R, stem, mode, conv_engine):
    """StemHRNet.forward (students.py:252-264) on the recorder."""
    if mode == "bf16" and conv_engine != L.ENGINE_FFMA:
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
