"""ORACLE (test infrastructure only) -- functional CPU restatement of
``PoseHigherResolutionNet.forward`` (rtpe/third_party/pose_higher_hrnet.py:637-686
and the blocks it calls: BasicBlock :46-75, Bottleneck :78-116,
HighResolutionModule.forward :238-256, transition layers :548-583, final layers
:447-483, deconv layers :485-533).

It consumes a plain ``state_dict`` with the reference's key names and evaluates the
network with torch CPU operators in a chosen dtype (float32 by default, float64
for error budgets), eval-mode BatchNorm (eps 1e-5).  The layer structure is read
off the state-dict keys, so any W32/W48 variant of the same topology works.

Pinned against the reference module itself (same weights, same input) in
tests/test_oracle_vs_reference.py and through tests/golden/hhrnet_*.npz.
Only tests/, smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import re

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class _SD:
    def __init__(self, sd, dtype):
        self.sd = {k: v.detach().to("cpu").to(dtype) for k, v in sd.items()
                   if v.dtype.is_floating_point}
        self.keys = set(sd.keys())

    def has(self, name):
        return name in self.keys

    def conv(self, x, name, stride=1, padding=0):
        b = self.sd.get(name + ".bias")
        return F.conv2d(x, self.sd[name + ".weight"], b, stride=stride, padding=padding)

    def bn(self, x, name):
        return F.batch_norm(x, self.sd[name + ".running_mean"], self.sd[name + ".running_var"],
                            self.sd[name + ".weight"], self.sd[name + ".bias"],
                            training=False, eps=BN_EPS)


def _basic_block(s: _SD, x, pfx):
    out = F.relu(s.bn(s.conv(x, pfx + ".conv1", 1, 1), pfx + ".bn1"))
    out = s.bn(s.conv(out, pfx + ".conv2", 1, 1), pfx + ".bn2")
    res = x
    if s.has(pfx + ".downsample.0.weight"):          # :64-65 (branches whose width changes)
        res = s.bn(s.conv(x, pfx + ".downsample.0"), pfx + ".downsample.1")
    return F.relu(out + res)


def _block(s: _SD, x, pfx):
    """BasicBlock (:46-75) or Bottleneck (:78-116), told apart by the third conv."""
    return _bottleneck(s, x, pfx) if s.has(pfx + ".conv3.weight") else _basic_block(s, x, pfx)


def _bottleneck(s: _SD, x, pfx):
    out = F.relu(s.bn(s.conv(x, pfx + ".conv1"), pfx + ".bn1"))
    out = F.relu(s.bn(s.conv(out, pfx + ".conv2", 1, 1), pfx + ".bn2"))
    out = s.bn(s.conv(out, pfx + ".conv3"), pfx + ".bn3")
    res = x
    if s.has(pfx + ".downsample.0.weight"):
        res = s.bn(s.conv(x, pfx + ".downsample.0"), pfx + ".downsample.1")
    return F.relu(out + res)


def _count(s: _SD, pattern):
    """number of distinct integer indices matching ``pattern`` (one group)."""
    rx = re.compile(pattern)
    idx = {int(m.group(1)) for k in s.keys for m in [rx.match(k)] if m}
    return (max(idx) + 1) if idx else 0


def _hr_module(s: _SD, xs, pfx):
    nb = len(xs)
    for i in range(nb):
        nblk = _count(s, re.escape(pfx) + r"\.branches\.%d\.(\d+)\.conv1\.weight" % i)
        for b in range(nblk):
            xs[i] = _block(s, xs[i], "%s.branches.%d.%d" % (pfx, i, b))
    nout = _count(s, re.escape(pfx) + r"\.fuse_layers\.(\d+)\.")
    outs = []
    for i in range(nout):
        y = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (pfx, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = s.bn(s.conv(xs[j], fp + ".0"), fp + ".1")
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = s.bn(s.conv(t, "%s.%d.0" % (fp, k), 2, 1), "%s.%d.1" % (fp, k))
                    if k != i - j - 1:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def _transition(s: _SD, ys, pfx, nb_new):
    xs = []
    for i in range(nb_new):
        tp = "%s.%d" % (pfx, i)
        if i < len(ys):
            if s.has(tp + ".0.weight"):
                xs.append(F.relu(s.bn(s.conv(ys[i], tp + ".0", 1, 1), tp + ".1")))
            else:
                xs.append(ys[i])
        else:
            t = ys[-1]
            k = 0
            while s.has("%s.%d.0.weight" % (tp, k)):
                t = F.relu(s.bn(s.conv(t, "%s.%d.0" % (tp, k), 2, 1), "%s.%d.1" % (tp, k)))
                k += 1
            xs.append(t)
    return xs


def hhrnet_forward_ref(state_dict, x: torch.Tensor, dtype=torch.float32):
    """-> [y0 (N, J+A, H/4, W/4), y1 (N, J, H/2, W/2)] in ``dtype`` on CPU."""
    sd = {(k[2:] if k.startswith("1.") else k): v for k, v in state_dict.items()}
    s = _SD(sd, dtype)
    x = x.detach().to("cpu").to(dtype)
    x = F.relu(s.bn(s.conv(x, "conv1", 2, 1), "bn1"))
    x = F.relu(s.bn(s.conv(x, "conv2", 2, 1), "bn2"))
    for b in range(_count(s, r"layer1\.(\d+)\.conv1\.weight")):
        x = _bottleneck(s, x, "layer1.%d" % b)

    ys = [x]
    for stage, trans in (("stage2", "transition1"), ("stage3", "transition2"),
                         ("stage4", "transition3")):
        nb = _count(s, stage + r"\.0\.branches\.(\d+)\.")
        if trans == "transition1":
            xs = []
            for i in range(nb):
                tp = "%s.%d" % (trans, i)
                if i == 0:
                    xs.append(F.relu(s.bn(s.conv(x, tp + ".0", 1, 1), tp + ".1")))
                else:
                    xs.append(F.relu(s.bn(s.conv(x, tp + ".0.0", 2, 1), tp + ".0.1")))
        else:
            xs = _transition(s, ys, trans, nb)
        for m in range(_count(s, stage + r"\.(\d+)\.branches")):
            xs = _hr_module(s, xs, "%s.%d" % (stage, m))
        ys = xs

    outs = []
    x = ys[0]
    fpad = (s.sd["final_layers.0.weight"].shape[-1] - 1) // 2
    y = s.conv(x, "final_layers.0", 1, fpad)
    outs.append(y)
    ndeconv = _count(s, r"deconv_layers\.(\d+)\.0\.0\.weight")
    for i in range(ndeconv):
        w = s.sd["deconv_layers.%d.0.0.weight" % i]
        k = w.shape[-1]                              # _get_deconv_cfg (:535-546)
        pad, opad = {4: (1, 0), 3: (1, 1), 2: (0, 0)}[k]
        if w.shape[0] != x.shape[1]:                 # deconv_cat
            x = torch.cat((x, y), 1)
        x = F.conv_transpose2d(x, w, None, stride=2, padding=pad, output_padding=opad)
        x = F.relu(s.bn(x, "deconv_layers.%d.0.1" % i))
        k = 1
        while s.has("deconv_layers.%d.%d.0.conv1.weight" % (i, k)):
            x = _basic_block(s, x, "deconv_layers.%d.%d.0" % (i, k))
            k += 1
        y = s.conv(x, "final_layers.%d" % (i + 1), 1, fpad)
        outs.append(y)
    return outs
