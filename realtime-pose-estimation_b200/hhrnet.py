"""``PoseHigherResolutionNet`` -- drop-in for
rtpe/third_party/pose_higher_hrnet.py:259-739, executed by libbrtpe.so.

The module tree (names, shapes, buffers) is the reference's, so
``network_to_half(model).load_state_dict(ckpt, strict=True)`` (rtpe/helpers.py:69-70)
works unchanged and ``state_dict()`` has the same 1810 entries.  The parameters are only
*containers*: ``forward`` never calls a torch operator on them.  At the first call for a
given (chunk, H, W, precision) the model

  1. folds every eval BatchNorm into its convolution (w' = w*gamma/sqrt(var+eps),
     b' = beta - mean*gamma/sqrt(var+eps)), re-lays the weights out per tap / K-major and
     rounds them to bf16 for the tcgen05 path (fp32 for the CUDA-core path),
  2. records the whole network as a launch plan in the native library
     (brtpe_plan_*: NHWC activations from a liveness-packed arena, one fused launch per
     conv+BN[+residual][+ReLU], one per cross-resolution fuse sum, the 4x4/s2 transposed
     conv as four parity phases writing the 2x map, the 48+34 channel concat as two
     writers of one 96-channel buffer),
  3. replays that plan as a CUDA graph per chunk of images.

Precision follows the parameters like the reference: fp32 parameters -> fp32 activations
on the FFMA path (error ~1e-6 of the tensor max); half/bfloat16 parameters (what
``network_to_half`` produces) -> bf16 activations, bf16 tcgen05 MMA with fp32 accumulate.
"""
from __future__ import annotations

import ctypes as C
import logging

import torch
import torch.nn as nn

from . import _lib as L

BN_MOMENTUM = 0.1
logger = logging.getLogger(__name__)

_TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
import os as _os
_REVERSE_CONV2 = _os.environ.get("BRTPE_CONV_REVERSE", "0") == "1"


class NoOpModule(nn.Module):
    """Parameter-free placeholder (pose_higher_hrnet.py:24-32)."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        return args


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias)


class BasicBlock(nn.Module):
    """Parameter container of pose_higher_hrnet.py:46-75."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(planes, planes, 3)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.downsample = downsample
        self.stride = stride


class Bottleneck(nn.Module):
    """Parameter container of pose_higher_hrnet.py:78-116."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = _conv(planes, planes, 3, stride)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv3 = _conv(planes, planes * self.expansion, 1)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class HighResolutionModule(nn.Module):
    """Parameter container of pose_higher_hrnet.py:119-256 (BASIC or BOTTLENECK blocks)."""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels,
                 fuse_method, multi_scale_output=True):
        super().__init__()
        for what, seq in (("NUM_BLOCKS", num_blocks), ("NUM_CHANNELS", num_channels),
                          ("NUM_INCHANNELS", num_inchannels)):
            if num_branches != len(seq):
                msg = "NUM_BRANCHES({}) <> {}({})".format(num_branches, what, len(seq))
                logger.error(msg)
                raise ValueError(msg)
        self.num_inchannels = num_inchannels
        self.fuse_method = fuse_method
        self.num_branches = num_branches
        self.multi_scale_output = multi_scale_output
        branches = []
        for i in range(num_branches):
            cin, cout = num_inchannels[i], num_channels[i] * blocks.expansion
            down = None
            if cin != cout:
                down = nn.Sequential(_conv(cin, cout, 1),
                                     nn.BatchNorm2d(cout, momentum=BN_MOMENTUM))
            layers = [blocks(cin, num_channels[i], 1, down)]
            num_inchannels[i] = cout
            layers += [blocks(cout, num_channels[i]) for _ in range(1, num_blocks[i])]
            branches.append(nn.Sequential(*layers))
        self.branches = nn.ModuleList(branches)
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(True)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        ch = self.num_inchannels
        rows = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:       # 1x1 + BN, then nearest upsample x2^(j-i)
                    row.append(nn.Sequential(
                        _conv(ch[j], ch[i], 1), nn.BatchNorm2d(ch[i]),
                        nn.Upsample(scale_factor=2 ** (j - i), mode="nearest")))
                elif j == i:
                    row.append(None)
                else:           # chain of (i-j) stride-2 3x3 convs
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        cout = ch[i] if last else ch[j]
                        mods = [_conv(ch[j], cout, 3, 2), nn.BatchNorm2d(cout)]
                        if not last:
                            mods.append(nn.ReLU(True))
                        chain.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*chain))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def get_num_inchannels(self):
        return self.num_inchannels


# ---------------------------------------------------------------------------------------
# plan recording (a tiny IR -> liveness-packed arena -> native plan)
# ---------------------------------------------------------------------------------------
class _T:
    """virtual NHWC tensor"""
    __slots__ = ("n", "h", "w", "ld", "cl", "first", "last", "buf", "keep", "lane", "touch", "inherit")

    def __init__(self, n, h, w, ld, cl=None):
        self.n, self.h, self.w, self.ld = n, h, w, ld     # ld: pixel stride in stored elements
        self.cl = ld if cl is None else cl                # logical channels per pixel
        self.first = None
        self.last = None
        self.buf = None
        self.keep = False
        self.lane = 0          # lane of the op that creates it (arena pools are per lane)
        self.touch = []        # (op index, is_write, phase group) of every op using it
        self.inherit = []      # ops that used the previous tenants of its arena buffer

    def numel(self):
        return self.n * self.h * self.w * self.ld


class _Recorder:
    def __init__(self, model, n, h, w, mode, engine, device, in_is_half):
        self.model = model
        self.n, self.h, self.w = n, h, w
        self.mode = mode                          # "fp32" | "bf16"
        self.engine = engine                      # ENGINE_AUTO | FFMA | UMMA
        self.device = device
        # fp32 mode on the tensor cores ("split"): activations are (hi, lo) bf16 pairs, every
        # product is three bf16 MMAs (include/brtpe.h BRTPE_DT_BF16X2); models opt in with
        # ``_split_fp32`` and ENGINE_FFMA keeps the CUDA-core float32 path
        self.split = bool(mode == "fp32" and engine != L.ENGINE_FFMA and
                          getattr(model, "_split_fp32", False))
        self.tc = mode == "bf16" or self.split    # tcgen05 engines
        self.dt = L.DT_BF16X2 if self.split else (L.DT_F32 if mode == "fp32" else L.DT_BF16)
        self.tdtype = torch.bfloat16 if self.tc else torch.float32
        self.ops = []
        self.tensors = []
        self.keepalive = []                       # packed weights / biases
        self.in_is_half = in_is_half
        self.stem_mode = 0                        # im2col mode bits 1 / 2 (flip pair, via fp16)
        self.lane = 0                             # lane of the ops being recorded
        self.op_lane = []                         # per op: lane
        self.op_rw = []                           # per op: (reads, writes, phase group)

    # -- tensors
    def new(self, n, h, w, ld):
        """virtual tensor with ``ld`` logical channels per pixel (split mode stores twice as many)"""
        t = _T(n, h, w, 2 * ld if self.split else ld, ld)
        self.tensors.append(t)
        return t

    def _touch(self, t, idx):
        if t.first is None:
            t.first = idx
            t.lane = self.lane
        t.last = idx

    def _sched(self, reads, writes, group=None):
        """Register the op being appended: lane + the tensors it reads / writes."""
        self.op_lane.append(self.lane)
        self.op_rw.append(([t for t in reads if t is not None], list(writes), group))

    # -- weights: BN fold + engine layout in ONE launch per layer (brtpe_prepack_weights)
    _WT = {torch.float32: L.WT_F32, torch.bfloat16: L.WT_BF16, torch.float16: L.WT_F16}

    def _dev(self, t, dtype=None):
        """parameter / buffer -> contiguous device tensor (no launch when it already is one)"""
        t = t.detach()
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if t.device != torch.device(self.device):
            t = t.to(self.device)
        return t.contiguous()

    def _prepack(self, d, weight, transposed, bn, conv_bias, khkw, cin_store, cout_module,
                 cin_index=None, im2col=False):
        """Folded, packed weights + bias of one conv launch.  ``weight``: the module's weight
        ((Cout,Cin,KH,KW), or (Cin,Cout,KH,KW) if ``transposed``); ``khkw``: kernel position of every
        packed tap; ``cin_index``: see ``conv``.  -> (packed, bias float32 [d.Cout])."""
        lib = L.load(require_cuda=False)
        w = self._dev(weight)
        if w.dtype not in self._WT:
            w = w.float()
        pd = L.PrepackDesc()
        pd.w_dtype = self._WT[w.dtype]
        pd.transposed = int(bool(transposed))
        pd.Cout = cout_module
        pd.Cin = w.shape[0] if transposed else w.shape[1]
        pd.KH, pd.KW = w.shape[2], w.shape[3]
        pd.ntaps = len(khkw)
        for i, (kh, kw) in enumerate(khkw):
            pd.tap_kh[i], pd.tap_kw[i] = kh, kw
        pd.im2col = int(bool(im2col))
        pd.Cin_store = cin_store
        pd.bn_eps = float(bn.eps) if bn is not None else 0.0
        eng = lib.brtpe_conv_select_engine(C.byref(d)) if d is not None else L.ENGINE_FFMA
        if eng < 0:
            L.check(eng, "brtpe_conv_select_engine")
        cout = d.Cout if d is not None else cout_module
        if eng in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO):
            cin_pad, cout_pad = C.c_int(0), C.c_int(0)
            lib.brtpe_umma_weight_dims(cin_store, d.Cout_store, C.byref(cin_pad), C.byref(cout_pad))
            pd.layout = L.PACK_KMAJOR_BF16
            pd.split = int(self.split)                     # K = [w_hi | w_hi | w_lo]
            pd.cin_pad = cin_pad.value * (3 if self.split else 1)
            pd.cout_pad = cout_pad.value
            packed = torch.empty((pd.ntaps, pd.cout_pad, pd.cin_pad), dtype=torch.bfloat16,
                                 device=self.device)
        else:
            pd.layout = L.PACK_CIN_COUT_F32
            pd.Cout_pack = cout
            if self.split:
                raise L.BrtpeError("fp32 mode on the tensor cores: layer not supported by a tcgen05 "
                                   "engine (%s)" % lib.brtpe_last_error().decode("utf-8", "replace"))
            pd.round_bf16 = int(self.mode == "bf16" and d is not None)   # tcgen05's operand rounding
            packed = torch.empty((pd.ntaps, cin_store, cout), dtype=torch.float32,
                                 device=self.device)
        bias = torch.empty((cout,), dtype=torch.float32, device=self.device)
        cb = None
        if conv_bias is not None:
            cb = self._dev(conv_bias, w.dtype)
        g = beta = mu = var = None
        if bn is not None:
            g, beta = self._dev(bn.weight, torch.float32), self._dev(bn.bias, torch.float32)
            mu = self._dev(bn.running_mean, torch.float32)
            var = self._dev(bn.running_var, torch.float32)
        ci = None
        if cin_index is not None:
            ci = torch.as_tensor(list(cin_index), dtype=torch.int32).to(self.device)
        if torch.device(self.device).type != "cuda":
            # host-side recording only (tests of the plan structure): nothing can run the plan
            return packed, bias
        with torch.cuda.device(self.device):
            L.check(lib.brtpe_prepack_weights(
                C.byref(pd), L.ptr(w), L.ptr(cb), L.ptr(g), L.ptr(beta), L.ptr(mu), L.ptr(var),
                L.ptr(ci), L.ptr(packed), L.ptr(bias), cout, L.stream_ptr(self.device)),
                "brtpe_prepack_weights")
        # the launch is asynchronous: its inputs must outlive it
        self.keepalive += [w, cb, g, beta, mu, var, ci]
        return packed, bias

    # -- ops
    def conv(self, x, conv, bn, relu, residual=None, out=None, out_coff=0, in_coff=0,
             cin_store=None, cout_store=None, pad_cout=False, cin_index=None, addends=None,
             out2=None):
        """conv (+BN) (+residual) (+ReLU) over virtual tensor x -> virtual tensor.
        ``cin_index``: for every stored input channel the module's input channel it carries, or -1
        for a pad channel (zero weights) -- inputs whose real channels sit in padded slots.
        ``addends``: [(tensor, shift), ...] HRNet fuse terms added in the epilogue, term k read with
        nearest-neighbour upsampling by 2^shift; ``out2``: second output -- then ``out`` receives
        act(conv + residual) and ``out2`` relu(out + sum of the addends) (brtpe_conv_desc.n_add)."""
        k = conv.kernel_size[0]
        s = conv.stride[0]
        dil = conv.dilation[0]                       # dilated 3x3 = the same taps, spread out
        cin, cout = conv.in_channels, conv.out_channels
        cin_store = cin if cin_store is None else cin_store
        cout_store = cout if cout_store is None else cout_store
        ho, wo = x.h // s, x.w // s
        if out is None:
            out = self.new(x.n, ho, wo, (cout_store + 15) // 16 * 16)
        taps = [(dy * dil, dx * dil) for dy, dx in _TAPS3] if k == 3 else [(0, 0)]
        ktaps = _TAPS3 if k == 3 else [(0, 0)]
        khkw = [(dy + k // 2, dx + k // 2) for dy, dx in ktaps]
        if cin_index is not None:
            cin = cin_store = len(cin_index)
        dcout = cout
        if pad_cout and cout_store % 16 == 0 and cout < cout_store and self.tc:
            # heads with 17 / 34 real channels: zero weights + zero bias for the pad channels, so
            # that the layer is a whole number of 16-channel chunks (vectorised epilogue; the pad
            # channels were written as zeros before, too)
            dcout = cout_store
        d = self._desc(x, in_coff, cin_store, taps, s, ho, wo, out, 1, 0, 0, dcout, cout_store,
                       out_coff, residual, relu)
        addends = list(addends or [])
        d.n_add = len(addends)
        for k, (t, sh) in enumerate(addends):
            d.add_ld[k], d.add_shift[k] = t.ld, sh
        d.out2_ld = out2.ld if out2 is not None else 0
        # conv2 of a residual block reads what conv1 has just written (and the block input conv1 has
        # just read): walking the images backwards it would start in the part of both that is still in
        # L2.  Measured neutral (27.09 / 27.21 / 26.96 vs 27.15 / 27.07 / 27.13 ms per 64 forwards,
        # same box): opt-in through BRTPE_CONV_REVERSE=1
        d.reverse_order = int(residual is not None and k == 3 and s == 1 and self.tc and
                              _REVERSE_CONV2)
        packed, b = self._prepack(d, conv.weight, False, bn, conv.bias, khkw, cin_store, cout,
                                  cin_index=cin_index)
        self._emit_conv(d, x, packed, b, residual, out, addends=addends, out2=out2)
        return out

    # output-parity phases of ConvTranspose2d(stride 2): output row 2*y + a takes the kernel rows kh from
    # input rows y + dy, as (dy, kh) lists per kernel size (padding / output_padding of
    # pose_higher_hrnet.py:535-546: k4 p1, k3 p1 op1, k2 p0); columns alike
    _DECONV_SEL = {4: ([(0, 1), (-1, 3)], [(1, 0), (0, 2)]),
                   3: ([(0, 1)], [(1, 0), (0, 2)]),
                   2: ([(0, 0)], [(0, 1)])}

    def deconv_s2(self, x, deconv, bn, cin_store, lanes=(0, 1, 2, 3)):
        """ConvTranspose2d(k, stride 2)+BN+ReLU as four output-parity convs of up to 2x2 taps
        (independent: they write disjoint pixels of the output, one lane each)."""
        cout = deconv.out_channels
        out = self.new(x.n, 2 * x.h, 2 * x.w, cout)
        sel = self._DECONV_SEL[deconv.kernel_size[0]]
        for a in (0, 1):
            for bb in (0, 1):
                taps, khkw = [], []
                for dy, kh in sel[a]:
                    for dx, kw in sel[bb]:
                        taps.append((dy, dx))
                        khkw.append((kh, kw))
                d = self._desc(x, 0, cin_store, taps, 1, x.h, x.w, out, 2, a, bb, cout, cout, 0,
                               None, True)
                packed, b = self._prepack(d, deconv.weight, True, bn, deconv.bias, khkw, cin_store,
                                          cout)
                keep = self.lane
                self.lane = lanes[(2 * a + bb) % len(lanes)]
                self._emit_conv(d, x, packed, b, None, out, group=("deconv", id(out)))
                self.lane = keep
        return out

    deconv4x4s2 = deconv_s2

    def _desc(self, x, in_coff, cin, taps, stride, hm, wm, out, out_scale, oy, ox, cout,
              cout_store, out_coff, residual, relu):
        d = L.ConvDesc()
        d.dtype = self.dt
        d.engine = self.engine if self.tc else L.ENGINE_FFMA
        d.N, d.Hin, d.Win = x.n, x.h, x.w
        d.Cin, d.in_ld, d.in_coff = cin, x.ld, in_coff
        d.Hm, d.Wm, d.in_stride = hm, wm, stride
        d.ntaps = len(taps)
        for i, (dy, dx) in enumerate(taps):
            d.tap_dy[i], d.tap_dx[i] = dy, dx
        d.Hout, d.Wout = out.h, out.w
        d.out_scale, d.out_oy, d.out_ox = out_scale, oy, ox
        d.Cout, d.out_ld, d.out_coff = cout, out.ld, out_coff
        d.res_ld = residual.ld if residual is not None else 0
        d.res_coff = 0
        d.relu = int(bool(relu))
        d.Cout_store = cout_store
        return d

    def _emit_conv(self, d, x, packed, bias, residual, out, group=None, addends=(), out2=None):
        self.keepalive += [packed, bias]
        idx = len(self.ops)
        adds = [t for t, _ in addends]
        self.ops.append(("conv", d, x, packed, bias, residual, out, adds, out2))
        self._sched([x, residual] + adds, [out] + ([out2] if out2 is not None else []), group)
        for t in [x, residual, out, out2] + adds:
            if t is not None:
                self._touch(t, idx)

    def block(self, x, blk, out=None):
        """One residual block (BasicBlock :46-75 or Bottleneck :78-116) -> output tensor; ``out``:
        the virtual tensor the block's last conv writes (a concat buffer)."""
        res = x
        if blk.downsample is not None:
            res = self.conv(x, blk.downsample[0], blk.downsample[1], False)
        if isinstance(blk, Bottleneck):
            t = self.conv(x, blk.conv1, blk.bn1, True)
            t = self.conv(t, blk.conv2, blk.bn2, True)
            return self.conv(t, blk.conv3, blk.bn3, True, residual=res, out=out)
        t = self.conv(x, blk.conv1, blk.bn1, True)
        return self.conv(t, blk.conv2, blk.bn2, True, residual=res, out=out)

    def stem(self, conv, bn):
        cout = conv.out_channels                                           # (64,3,3,3)
        # (ky,kx,ci) x Cout float32: the im2col ordering with one 27-channel "tap"
        wp, b = self._prepack(None, conv.weight, False, bn, conv.bias, [(0, 0)], 27, cout,
                              im2col=True)
        wp = wp.view(27, cout)
        out = self.new(self.n, self.h // 2, self.w // 2, cout)
        self.keepalive += [wp, b]
        idx = len(self.ops)
        self.ops.append(("stem", wp, b, out))
        self._sched([], [out])
        self._touch(out, idx)
        return out

    def stem_tc(self, conv, bn):
        """conv1 + BN + ReLU on the tensor cores: im2col (27 -> 32 channels per output pixel)
        followed by a 1x1 tcgen05 conv with K = 32."""
        cout = conv.out_channels
        cols = self.new(self.n, self.h // 2, self.w // 2, 32)
        idx = len(self.ops)
        self.ops.append(("im2col", cols))
        self._sched([], [cols])
        self._touch(cols, idx)
        out = self.new(self.n, self.h // 2, self.w // 2, cout)
        d = self._desc(cols, 0, 32, [(0, 0)], 1, cols.h, cols.w, out, 1, 0, 0, cout, cout, 0,
                       None, True)
        packed, b = self._prepack(d, conv.weight, False, bn, conv.bias, [(0, 0)], 32, cout,
                                  im2col=True)                             # k = (ky, kx, ci)
        self._emit_conv(d, cols, packed, b, None, out)
        return out

    def fuse(self, terms, shifts, c, relu, out=None):
        t0 = terms[0]
        if out is None:
            out = self.new(t0.n, t0.h << shifts[0], t0.w << shifts[0], c)
        idx = len(self.ops)
        self.ops.append(("fuse", list(terms), list(shifts), c, relu, out))
        self._sched(list(terms), [out])
        for t in terms:
            self._touch(t, idx)
        self._touch(out, idx)
        return out

    def aux(self, kind, reads, writes, in0, in1, in2, out, iparams):
        """Student-only NHWC op (brtpe_plan_add_aux).  in*/out: virtual tensors or raw torch
        tensors (weights, float side buffers)."""
        idx = len(self.ops)
        self.ops.append(("aux", kind, in0, in1, in2, out, [int(v) for v in iparams]))
        self._sched(list(reads), list(writes))
        for t in list(reads) + list(writes):
            self._touch(t, idx)

    def to_nchw(self, x, c, coff, dst):
        idx = len(self.ops)
        self.ops.append(("nchw", x, c, coff, dst))
        self._sched([x], [])
        self._touch(x, idx)

    # -- arena + native plan
    def build(self, in_buf):
        lib = L.load(require_cuda=False)
        esize = 2 if self.tc else 4
        # Liveness-packed arena.  Released buffers go back to the pool of the lane that
        # created them and are only re-used by that lane, so arena re-use does not serialise
        # independent lanes; the ops that used the previous tenants are still recorded as
        # dependencies of the new tenant's first writer (`inherit`).
        free = {}                                  # lane -> [(bytes, storage, users)]
        release_at = {}
        for t in self.tensors:
            if t.first is not None and not t.keep:
                release_at.setdefault(t.last, []).append(t)
        first_at = {}
        for t in self.tensors:
            if t.first is not None:
                first_at.setdefault(t.first, []).append(t)
        total = 0
        for idx, op in enumerate(self.ops):
            for t in first_at.get(idx, []):
                if t.buf is not None:
                    continue
                need = t.numel() * esize
                pool = free.setdefault(t.lane, [])
                best = None
                for k, (nb, buf, users) in enumerate(pool):
                    if nb >= need and (best is None or nb < pool[best][0]):
                        best = k
                if best is not None:
                    nb, buf, users = pool.pop(best)
                    t.inherit = list(users)
                else:
                    nb = (need + 255) // 256 * 256
                    buf = torch.zeros(nb, dtype=torch.uint8, device=self.device)
                    total += nb
                    t.inherit = []
                t.buf = (nb, buf)
            reads, writes, group = self.op_rw[idx]
            for t in reads:
                t.touch.append((idx, False, group))
            for t in writes:
                t.touch.append((idx, True, group))
            for t in release_at.get(idx, []):
                users = sorted(set(t.inherit) | set(i for i, _, _ in t.touch))
                free.setdefault(t.lane, []).append((t.buf[0], t.buf[1], users))
        self.arena_bytes = total

        # dependencies: RAW on every earlier writer, WAR/WAW on every earlier user (ops of one
        # phase group write disjoint parts of a tensor and do not order each other), plus the
        # users of the arena buffer's previous tenants
        self.deps = []
        for idx in range(len(self.ops)):
            reads, writes, group = self.op_rw[idx]
            dep = set()
            for t in reads:
                dep.update(i for i, wr, _ in t.touch if wr and i < idx)
                dep.update(i for i in t.inherit if i < idx)
            for t in writes:
                dep.update(i for i, _, g in t.touch
                           if i < idx and not (group is not None and g == group))
                dep.update(i for i in t.inherit if i < idx)
            lane = self.op_lane[idx]
            self.deps.append(sorted(i for i in dep if self.op_lane[i] != lane))

        plan = lib.brtpe_plan_create()
        try:
            for op in self.ops:
                kind = op[0]
                if kind == "conv":
                    _, d, x, packed, bias, residual, out, adds, out2 = op
                    L.check(lib.brtpe_plan_add_conv(
                        plan, C.byref(d), L.ptr(x.buf[1]), L.ptr(packed), L.ptr(bias),
                        L.ptr(residual.buf[1]) if residual is not None else None,
                        L.ptr(out.buf[1])), "brtpe_plan_add_conv")
                    if adds or out2 is not None:
                        ap = (C.c_void_p * max(len(adds), 1))(*[t.buf[1].data_ptr() for t in adds])
                        L.check(lib.brtpe_plan_set_conv_fuse(
                            plan, ap, L.ptr(out2.buf[1]) if out2 is not None else None),
                            "brtpe_plan_set_conv_fuse")
                elif kind == "stem":
                    _, wp, b, out = op
                    L.check(lib.brtpe_plan_add_stem(
                        plan, L.ptr(in_buf), int(self.in_is_half), self.n, self.h, self.w,
                        L.ptr(wp), L.ptr(b), wp.shape[1], L.ptr(out.buf[1]), self.dt),
                        "brtpe_plan_add_stem")
                elif kind == "fuse":
                    _, terms, shifts, c, relu, out = op
                    nt = len(terms)
                    tp = (C.c_void_p * nt)(*[t.buf[1].data_ptr() for t in terms])
                    sh = (C.c_int32 * nt)(*shifts)
                    ld = (C.c_int32 * nt)(*[t.ld for t in terms])
                    L.check(lib.brtpe_plan_add_fuse(
                        plan, self.dt, nt, tp, sh, ld, out.n, out.h, out.w, c,
                        L.ptr(out.buf[1]), out.ld, int(bool(relu))), "brtpe_plan_add_fuse")
                elif kind == "nchw":
                    _, x, c, coff, dst = op
                    L.check(lib.brtpe_plan_add_nhwc_to_nchw(
                        plan, self.dt, L.ptr(x.buf[1]), x.n, x.h, x.w, c, x.ld, coff,
                        L.ptr(dst), int(dst.dtype == torch.float16)),
                        "brtpe_plan_add_nhwc_to_nchw")
                elif kind == "aux":
                    _, akind, in0, in1, in2, out, ip = op

                    def _p(t):
                        if t is None:
                            return None
                        return L.ptr(t.buf[1]) if isinstance(t, _T) else L.ptr(t)
                    arr = (C.c_int32 * len(ip))(*ip)
                    L.check(lib.brtpe_plan_add_aux(plan, akind, _p(in0), _p(in1), _p(in2), _p(out),
                                                   arr, len(ip)), "brtpe_plan_add_aux")
                elif kind == "im2col":
                    _, cols = op
                    L.check(lib.brtpe_plan_add_stem_im2col(
                        plan, L.ptr(in_buf), int(self.in_is_half) | int(self.stem_mode) |
                        (8 if self.split else 0), self.n,
                        self.h, self.w, L.ptr(cols.buf[1])), "brtpe_plan_add_stem_im2col")
                k = lib.brtpe_plan_num_ops(plan) - 1
                deps = self.deps[k]
                arr = (C.c_int32 * max(len(deps), 1))(*deps)
                L.check(lib.brtpe_plan_set_sched(plan, self.op_lane[k], arr, len(deps)),
                        "brtpe_plan_set_sched")
        except Exception:
            lib.brtpe_plan_destroy(plan)
            raise
        return plan


class _CompiledPlan:
    def __init__(self, handle, in_buf, outs, recorder):
        self.handle = handle
        self.in_buf = in_buf
        self.outs = outs
        self.recorder = recorder            # keeps arena + packed weights alive
        self.num_ops = L.load(require_cuda=False).brtpe_plan_num_ops(handle)
        self.conv_flops = L.load(require_cuda=False).brtpe_plan_conv_flops(handle)

    def __del__(self):
        try:
            if self.handle:
                L.load(require_cuda=False).brtpe_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PlanRunner:
    """Shared execution machinery of the drop-in modules: a module records itself once per
    (chunk, H, W, precision) as a native launch plan (``_record``) and ``forward`` replays the
    plan (CUDA graph) per chunk of images.  Sub-classes provide ``_record`` and ``_ref_param``
    (any parameter: its dtype selects the precision mode, its device the GPU)."""

    def _init_runner(self):
        # execution options (not part of the reference API)
        self.chunk_size = 8            # images per plan replay
        self.conv_engine = L.ENGINE_AUTO
        self.use_cuda_graph = True
        self.parallel_branches = True  # resolution branches = parallel branches of the CUDA graph
        self._plans = {}
        self._sig = None
        self._frozen = False

    # compiled plans own native handles and device arenas: never copied / pickled
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_plans"] = {}
        state["_sig"] = None
        return state

    def invalidate_plans(self):
        """Drop the compiled plans (packed weights are rebuilt at the next forward)."""
        self._plans = {}
        self._sig = None

    def load_state_dict(self, *args, **kwargs):
        """nn.Module.load_state_dict + the compiled plans (BN-folded, packed weights) are dropped."""
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_plans()
        return out

    def _apply(self, fn, *args, **kwargs):
        """.to() / .cuda() / .half() / .float(): the packed weights follow the parameters."""
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_plans"):
            self.invalidate_plans()
        return out

    def _signature(self):
        """Cheap fingerprint of every parameter / buffer (storage address + in-place version
        counter): load_state_dict, .half(), .to() and optimizer steps all change it.  Writes
        through ``.data`` (``p.data.copy_()``, the reference's fp16_utils
        ``master_params_to_model_params``) do NOT bump the version counter: call
        ``invalidate_plans()`` after such an update, the cached plans keep the old folded weights
        otherwise (a content checksum would cost a device read + host sync per forward)."""
        acc = 0
        for t in list(self.parameters()) + list(self.buffers()):
            acc = (acc * 1000003 + t.data_ptr() * 31 + t._version) & 0xFFFFFFFFFFFFFFFF
        return acc

    def freeze(self, frozen=True):
        """Skip the per-forward parameter fingerprint (weights promised not to change)."""
        self._frozen = bool(frozen)
        return self

    def _mode(self):
        return "fp32" if self._ref_param().dtype == torch.float32 else "bf16"

    def _plan_extra_key(self):
        """Forward arguments that change the recorded plan (e.g. a student's ``out_hw``)."""
        return None

    def _get_plan(self, n, h, w, mode, device, in_dtype, stem_mode=0):
        key = (n, h, w, mode, str(device), in_dtype, self.conv_engine, self.parallel_branches,
               stem_mode, self._plan_extra_key(), getattr(self, "fuse_in_epilogue", False))
        plan = self._plans.get(key)
        if plan is None:
            in_is_half = in_dtype == torch.float16
            # flip-pair plans (stem_mode bit 1) read n / 2 images and run n forwards
            in_buf = torch.empty((n // 2 if stem_mode & 2 else n, 3, h, w), dtype=in_dtype,
                                 device=device)
            with torch.cuda.device(device):
                if stem_mode:
                    R, outs = self._record(n, h, w, mode, device, in_is_half, in_is_half,
                                           stem_mode=stem_mode)
                else:
                    R, outs = self._record(n, h, w, mode, device, in_is_half, in_is_half)
                handle = R.build(in_buf)
            plan = _CompiledPlan(handle, in_buf, outs, R)
            self._plans[key] = plan
        return plan

    def _run_plans(self, x, size_multiple, flip_pair=False, via_half=False, borrow=False,
                   extra=None):
        """``flip_pair``: run the network on cat(x, flip(x, [3])) without materialising the batch
        (rows [0:N] of every output belong to x, rows [N:2N] to the mirrored images);
        ``via_half``: a float32 input is rounded through fp16 first (the tofp16 wrapper);
        ``borrow``: when one plan replay covers the whole batch, return the plan's own output
        buffers instead of copies -- valid only until the next forward of this module (callers that
        consume the outputs in stream order before that, like ``TeacherPipeline``);
        ``extra``: second per-image input (N, ...) copied chunk by chunk into the buffer the
        recorder exposes as ``extra_input`` (``AttentionStudentSteps``' ``alt`` image)."""
        lib = L.load()
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise L.BrtpeError("%s.forward needs a CUDA tensor (no CPU fallback); got %r"
                               % (type(self).__name__, getattr(x, "device", type(x)),))
        if self.training:
            raise L.BrtpeError("this implementation is inference-only: call .eval() "
                               "(BatchNorm is folded with its running statistics)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input (N, 3, H, W), got %s" % (tuple(x.shape),))
        n, _, h, w = x.shape
        if h % size_multiple or w % size_multiple:
            raise ValueError("H and W must be multiples of %d (got %dx%d)" % (size_multiple, h, w))
        if x.dtype == torch.bfloat16:
            x = x.to(torch.float32)
        if x.dtype not in (torch.float32, torch.float16):
            raise ValueError("input dtype must be float32 or float16")
        mode = self._mode()
        dev = x.device
        x = x.contiguous()
        if not self._frozen or self._sig is None:
            sig = self._signature()
            if sig != self._sig:
                self._plans = {}
                self._sig = sig
        results = None
        if flip_pair:
            stem_mode = 2 | (4 if via_half and x.dtype == torch.float32 else 0)
            nb = max(1, min(self.chunk_size, 2 * n) // 2)          # originals per plan replay
            with torch.cuda.device(dev):
                st = L.stream_ptr(dev)
                for s0 in range(0, n, nb):
                    cn = min(nb, n - s0)
                    plan = self._get_plan(2 * cn, h, w, mode, dev, x.dtype, stem_mode)
                    plan.in_buf.copy_(x[s0:s0 + cn])
                    if self.use_cuda_graph:
                        L.check(lib.brtpe_plan_graph_launch(plan.handle, st), "brtpe_plan_graph_launch")
                    else:
                        L.check(lib.brtpe_plan_run(plan.handle, st), "brtpe_plan_run")
                    if borrow and cn == n:
                        return list(plan.outs)
                    if results is None:
                        results = [torch.empty((2 * n,) + tuple(o.shape[1:]), dtype=o.dtype, device=dev)
                                   for o in plan.outs]
                    for r, o in zip(results, plan.outs):
                        r[s0:s0 + cn].copy_(o[:cn])
                        r[n + s0:n + s0 + cn].copy_(o[cn:])
            return results
        nb = min(self.chunk_size, n)
        with torch.cuda.device(dev):
            st = L.stream_ptr(dev)
            for s0 in range(0, n, nb):
                cn = min(nb, n - s0)
                plan = self._get_plan(cn, h, w, mode, dev, x.dtype)
                plan.in_buf.copy_(x[s0:s0 + cn])
                if extra is not None:
                    plan.recorder.extra_input.copy_(extra[s0:s0 + cn])
                if self.use_cuda_graph:
                    L.check(lib.brtpe_plan_graph_launch(plan.handle, st), "brtpe_plan_graph_launch")
                else:
                    L.check(lib.brtpe_plan_run(plan.handle, st), "brtpe_plan_run")
                if results is None:
                    results = [torch.empty((n,) + tuple(o.shape[1:]), dtype=o.dtype, device=dev)
                               for o in plan.outs]
                for r, o in zip(results, plan.outs):
                    r[s0:s0 + cn].copy_(o)
        return results

    def plan_profile(self, n, h, w, in_dtype=torch.float32):
        """Per-launch device times of one plan replay: (ms[], kind[], flops[])."""
        lib = L.load()
        dev = self._ref_param().device
        plan = self._get_plan(n, h, w, self._mode(), dev, in_dtype)
        k = plan.num_ops
        ms = (C.c_float * k)()
        kinds = (C.c_int32 * k)()
        fl = (C.c_double * k)()
        with torch.cuda.device(dev):
            L.check(lib.brtpe_plan_profile(plan.handle, L.stream_ptr(dev), ms, kinds, fl),
                    "brtpe_plan_profile")
        return list(ms), list(kinds), list(fl)


class PoseHigherResolutionNet(_PlanRunner, nn.Module):
    """Same constructor and ``forward(x) -> [y0, y1]`` as the reference
    (pose_higher_hrnet.py:266-287, :637-686)."""

    BLOCK_TYPES = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}

    def __init__(self, num_joints=17, tag_per_joint=True, final_conv_ksize=1,
                 pretrained_layers=["*"], inplanes=64,
                 s2_modules=1, s2_branches=2, s2_block_type="BASIC",
                 s2_blocks=[4, 4], s2_chans=[48, 96], s2_fuse_method="SUM",
                 s3_modules=4, s3_branches=3, s3_block_type="BASIC",
                 s3_blocks=[4, 4, 4], s3_chans=[48, 96, 192], s3_fuse_method="SUM",
                 s4_modules=3, s4_branches=4, s4_block_type="BASIC",
                 s4_blocks=[4, 4, 4, 4], s4_chans=[48, 96, 192, 384], s4_fuse_method="SUM",
                 deconvs=1, deconv_chans=[48], deconv_ksize=[4], deconv_num_blocks=4,
                 deconv_cat=[True], with_ae_loss=(True, False)):
        super().__init__()
        for bt in (s2_block_type, s3_block_type, s4_block_type):
            if bt not in self.BLOCK_TYPES:
                raise KeyError(bt)                      # like the reference's BLOCK_TYPES lookup
        if final_conv_ksize not in (1, 3):
            raise ValueError("final_conv_ksize must be 1 or 3 (pose_higher_hrnet.py:460-482)")
        if any(k not in (2, 3, 4) for k in deconv_ksize[:deconvs]):
            raise ValueError("deconv kernel sizes must be 2, 3 or 4 (pose_higher_hrnet.py:535-546)")
        self.inplanes = inplanes
        self.cfg = {"NUM_JOINTS": num_joints, "TAG_PER_JOINT": tag_per_joint,
                    "FINAL_CONV_KSIZE": final_conv_ksize, "PRETRAINED_LAYERS": pretrained_layers}
        stages = [("STAGE2", s2_modules, s2_branches, s2_blocks, s2_chans, s2_fuse_method, s2_block_type),
                  ("STAGE3", s3_modules, s3_branches, s3_blocks, s3_chans, s3_fuse_method, s3_block_type),
                  ("STAGE4", s4_modules, s4_branches, s4_blocks, s4_chans, s4_fuse_method, s4_block_type)]
        for name, nm, nb, blocks, chans, fm, bt in stages:
            self.cfg[name] = {"num_modules": nm, "num_branches": nb,
                              "block_cls": self.BLOCK_TYPES[bt], "num_blocks": blocks,
                              "num_channels": chans, "fuse_method": fm}
        self.cfg["DECONV"] = {"num_deconvs": deconvs, "num_channels": deconv_chans,
                              "kernel_size": deconv_ksize, "num_basic_blocks": deconv_num_blocks,
                              "cat_output": deconv_cat}

        # stem
        self.conv1 = _conv(3, 64, 3, 2)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.conv2 = _conv(64, 64, 3, 2)
        self.bn2 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(Bottleneck, 64, 4)

        pre = [256]
        for si, (name, nm, nb, blocks, chans, fm, bt) in enumerate(stages):
            block_cls = self.BLOCK_TYPES[bt]
            cur = [c * block_cls.expansion for c in chans]
            setattr(self, "transition%d" % (si + 1), self._make_transition_layer(pre, cur))
            last_stage = si == len(stages) - 1
            mods = []
            inch = cur
            for m in range(nm):
                multi = not (last_stage and m == nm - 1)
                mods.append(HighResolutionModule(nb, block_cls, blocks, inch, chans, fm, multi))
                inch = mods[-1].get_num_inchannels()
            setattr(self, "stage%d" % (si + 2), nn.Sequential(*mods))
            pre = inch

        ae_dims = num_joints if tag_per_joint else 1
        heads = []
        cin = pre[0]
        for i in range(deconvs + 1):
            if i > 0:
                cin = deconv_chans[i - 1]
            cout = num_joints + (ae_dims if with_ae_loss[i] else 0)
            heads.append(nn.Conv2d(cin, cout, final_conv_ksize, 1,
                                   1 if final_conv_ksize == 3 else 0))
        self.final_layers = nn.ModuleList(heads)

        dls = []
        cin = pre[0]
        for i in range(deconvs):
            if deconv_cat[i]:
                cin += num_joints + (ae_dims if with_ae_loss[i] else 0)
            cout = deconv_chans[i]
            dk = deconv_ksize[i]                        # _get_deconv_cfg (:535-546)
            layers = [nn.Sequential(
                nn.ConvTranspose2d(cin, cout, kernel_size=dk, stride=2, padding=0 if dk == 2 else 1,
                                   output_padding=1 if dk == 3 else 0, bias=False),
                nn.BatchNorm2d(cout, momentum=BN_MOMENTUM), nn.ReLU(inplace=True))]
            layers += [nn.Sequential(BasicBlock(cout, cout)) for _ in range(deconv_num_blocks)]
            dls.append(nn.Sequential(*layers))
            cin = cout
        self.deconv_layers = nn.ModuleList(dls)

        self.num_deconvs = deconvs
        self.pretrained_layers = pretrained_layers
        self.deconv_cat = deconv_cat
        self.num_joints = num_joints

        self._init_runner()
        self._split_fp32 = True        # float32 parameters -> split-bf16 tcgen05 path (<= 1e-4)
        # HRNet fuse sums folded into conv epilogues (bf16 tcgen05 path): parity-green and the
        # fuse_sum launches disappear, but measured 1.3 % SLOWER than the stand-alone fuse_sum kernel
        # (26.63 vs 26.98 ms per 64 forwards, same box, profiles/r02_fuse_epilogue.md: the epilogue of
        # the 48-channel layers is the critical stage and the addend loads lengthen it), hence opt-in:
        # ``net.fuse_in_epilogue = True`` (+ ``invalidate_plans()``) or BRTPE_FUSE_EPILOGUE=1
        import os
        self.fuse_in_epilogue = os.environ.get("BRTPE_FUSE_EPILOGUE", "0") == "1"

    # ---------------------------------------------------------------- construction helpers
    def _make_layer(self, block, planes, blocks, stride=1):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride,
                          bias=False),
                nn.BatchNorm2d(planes * block.expansion, momentum=BN_MOMENTUM))
        layers = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _make_transition_layer(self, pre, cur):
        out = []
        for i, c in enumerate(cur):
            if i < len(pre):
                if c != pre[i]:
                    out.append(nn.Sequential(_conv(pre[i], c, 3), nn.BatchNorm2d(c),
                                             nn.ReLU(inplace=True)))
                else:
                    out.append(NoOpModule())
            else:
                chain = []
                for j in range(i + 1 - len(pre)):
                    cout = c if j == i - len(pre) else pre[-1]
                    chain.append(nn.Sequential(_conv(pre[-1], cout, 3, 2), nn.BatchNorm2d(cout),
                                               nn.ReLU(inplace=True)))
                out.append(nn.Sequential(*chain))
        return nn.ModuleList(out)

    def _ref_param(self):
        return self.conv1.weight

    def init_weights(self, pretrained="", verbose=True):
        """pose_higher_hrnet.py:688-727: N(0, 0.001) convs, unit BN, optional partial load."""
        import os
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if os.path.isfile(pretrained):
            sd = torch.load(pretrained)
            own = set(n for n, _ in self.named_parameters()) | set(n for n, _ in self.named_buffers())
            pick = {n: v for n, v in sd.items()
                    if (n.split(".")[0] in self.pretrained_layers or self.pretrained_layers[0] == "*")
                    and n in own}
            self.load_state_dict(pick, strict=False)
        self.invalidate_plans()

    # ---------------------------------------------------------------- plan
    def _record(self, n, h, w, mode, device, in_is_half, out_half, stem_mode=0):
        R = _Recorder(self, n, h, w, mode, self.conv_engine, device, in_is_half)
        R.stem_mode = stem_mode
        par = self.parallel_branches
        if R.tc and self.conv_engine != L.ENGINE_FFMA:
            x = R.stem_tc(self.conv1, self.bn1)
        else:
            x = R.stem(self.conv1, self.bn1)
        x = R.conv(x, self.conv2, self.bn2, True)
        for blk in self.layer1:
            res = x
            if blk.downsample is not None:
                res = R.conv(x, blk.downsample[0], blk.downsample[1], False)
            t = R.conv(x, blk.conv1, blk.bn1, True)
            t = R.conv(t, blk.conv2, blk.bn2, True)
            x = R.conv(t, blk.conv3, blk.bn3, True, residual=res)

        ys = [x]
        nj = self.num_joints
        # HRNet cross-resolution fuse-add in the conv epilogues (tcgen05 bf16 path); the stand-alone
        # fuse_sum kernel stays for the fp32 modes and as the A/B reference (fuse_in_epilogue = False)
        fuse_epi = bool(self.fuse_in_epilogue and R.tc and not R.split and
                        self.conv_engine != L.ENGINE_FFMA and
                        all(self.cfg[k]["block_cls"] is BasicBlock for k in ("STAGE2", "STAGE3", "STAGE4")))
        head0 = self.final_layers[0]
        cat = self.deconv_cat[0] if self.num_deconvs else False
        cat_ld = None
        for si in (1, 2, 3):
            trans = getattr(self, "transition%d" % si)
            stage = getattr(self, "stage%d" % (si + 1))
            xs = []
            for i, tr in enumerate(trans):
                R.lane = i if par else 0              # lane = resolution branch
                if isinstance(tr, NoOpModule):
                    xs.append(ys[i])
                elif i < len(ys) and not isinstance(tr[0], nn.Sequential):
                    xs.append(R.conv(ys[i], tr[0], tr[1], True))
                else:
                    t = ys[-1]
                    for sub in tr:
                        t = R.conv(t, sub[0], sub[1], True)
                    xs.append(t)
            for mi, mod in enumerate(stage):
                final_module = si == 3 and mi == len(stage) - 1
                dst0 = None
                if final_module and cat:
                    # stage-4 output lands in channels [0, c) of the concat buffer
                    c = mod.num_inchannels[0]
                    cat_ld = (c + head0.out_channels + 15) // 16 * 16
                    dst0 = R.new(n, xs[0].h, xs[0].w, cat_ld)
                if fuse_epi and mod.fuse_layers is not None:
                    xs = self._record_module_fused(R, mod, xs, par, dst0)
                    continue
                # interleave the branches block by block: the recording order is also the
                # issue order of the capture, so independent launches sit next to each other
                nblk = max(len(b) for b in mod.branches)
                for bi in range(nblk):
                    for i in range(mod.num_branches):
                        if bi >= len(mod.branches[i]):
                            continue
                        R.lane = i if par else 0
                        xs[i] = R.block(xs[i], mod.branches[i][bi])
                outs = []
                nout = len(mod.fuse_layers)
                terms_of = [[None] * mod.num_branches for _ in range(nout)]
                shifts_of = [[0] * mod.num_branches for _ in range(nout)]
                for j in range(mod.num_branches):     # fuse convs run on their source lane
                    R.lane = j if par else 0
                    for i in range(nout):
                        f = mod.fuse_layers[i][j]
                        if j == i:
                            terms_of[i][j] = xs[j]
                        elif j > i:
                            terms_of[i][j] = R.conv(xs[j], f[0], f[1], False)
                            shifts_of[i][j] = j - i
                        else:
                            t = xs[j]
                            for k, sub in enumerate(f):
                                t = R.conv(t, sub[0], sub[1], k != len(f) - 1)
                            terms_of[i][j] = t
                for i in range(nout):
                    R.lane = i if par else 0
                    c = mod.num_inchannels[i]
                    outs.append(R.fuse(terms_of[i], shifts_of[i], c, True,
                                       out=dst0 if i == 0 else None))
                xs = outs
            ys = xs
        R.lane = 0

        x = ys[0]
        odt = torch.float16 if out_half else torch.float32
        outs = []
        nd = self.num_deconvs
        for i in range(nd + 1):
            head = self.final_layers[i]
            c0 = head.in_channels
            yo = torch.empty((n, head.out_channels, x.h, x.w), dtype=odt, device=device)
            outs.append(yo)
            if i < nd and self.deconv_cat[i]:
                # torch.cat((x, y), 1) is never built: x already lives in channels [0, c0) of a buffer
                # that is wide enough for the head's channels, which the head writes behind it
                assert x.cl >= c0 + head.out_channels, "concat buffer was not reserved"
                R.lane = 0
                R.conv(x, head, None, False, out=x, out_coff=c0, in_coff=0, cin_store=c0,
                       cout_store=x.cl - c0, pad_cout=True)
                R.lane = 1 if par else 0
                R.to_nchw(x, head.out_channels, c0, yo)
                R.lane = 0
                dc_cin = x.cl
            else:
                y = R.conv(x, head, None, False, cin_store=c0,
                           cout_store=(head.out_channels + 15) // 16 * 16, pad_cout=True)
                R.to_nchw(y, head.out_channels, 0, yo)
                dc_cin = c0
            if i == nd:
                break
            dl = self.deconv_layers[i]
            x = R.deconv_s2(x, dl[0][0], dl[0][1], dc_cin, lanes=(0, 1, 2, 3) if par else (0,))
            nblocks = len(dl) - 1
            for k in range(1, len(dl)):
                dst = None
                if k == nblocks and i + 1 < nd and self.deconv_cat[i + 1]:
                    nxt_head = self.final_layers[i + 1]
                    dst = R.new(n, x.h, x.w, (x.cl + nxt_head.out_channels + 15) // 16 * 16)
                x = R.block(x, dl[k][0], out=dst)
            if nblocks == 0 and i + 1 < nd and self.deconv_cat[i + 1]:
                raise NotImplementedError("deconv_cat after a deconv stage without blocks")
        return R, outs

    def _record_module_fused(self, R, mod, xs, par, dst0):
        """One HighResolutionModule with the fuse sums (pose_higher_hrnet.py:245-254) folded into
        conv epilogues.  Every conv takes at most ONE same-resolution addend (its residual) and
        ONE nearest-upsampled addend (shift 1), so the sums are built as chains:
          * upsampled terms of output i (branches j > i): W_{B-1} = c_{i,B-1}(x_{B-1}) and
            W_j = c_ij(x_j) + up2(W_{j+1}) in the epilogue of the 1x1 conv c_ij -- nearest
            upsampling composes exactly (up4 = up2 o up2), the partial sums live at low resolution;
          * stride-2 chains of output i >= 1 (branches j < i): the last conv of the chain from
            branch j takes the running sum S (starting with x_i) as residual; the chain from
            branch 0 comes last, adds up2(W_{i+1}) and applies the ReLU;
          * output 0: y_0 = relu(x_0 + up2(W_1)) leaves the epilogue of branch 0's LAST conv2, which
            writes x_0 (the stride-2 chains read it) and y_0.
        Recording order = dependency order: the other branches finish first, then their fuse
        convs, then branch 0's last conv2, then the chains that start from x_0."""
        nb = mod.num_branches
        nout = len(mod.fuse_layers)
        xs = list(xs)
        nblk = max(len(b) for b in mod.branches)
        last0 = None
        for bi in range(nblk):
            for i in range(nb):
                if bi >= len(mod.branches[i]):
                    continue
                R.lane = i if par else 0
                blk = mod.branches[i][bi]
                t = R.conv(xs[i], blk.conv1, blk.bn1, True)
                if i == 0 and bi == len(mod.branches[0]) - 1:
                    last0 = (t, blk, xs[0])              # conv2 of branch 0's last block comes later
                else:
                    xs[i] = R.conv(t, blk.conv2, blk.bn2, True, residual=xs[i])
        # upsampled terms: W[i] = sum_{j > i} up_{2^(j-i-1)}(c_ij(x_j)) at the resolution of branch i+1
        W = {}
        for i in range(min(nout, nb - 1)):
            w = None
            for j in range(nb - 1, i, -1):
                R.lane = j if par else 0
                f = mod.fuse_layers[i][j]
                w = R.conv(xs[j], f[0], f[1], False, addends=[(w, 1)] if w is not None else None)
            W[i] = w
        # stride-2 chains that do not start from branch 0: running sums S[i] = x_i + chains
        S = {i: xs[i] for i in range(1, nout)}
        for j in range(1, nb):
            R.lane = j if par else 0
            for i in range(j + 1, nout):
                t = xs[j]
                f = mod.fuse_layers[i][j]
                for k, sub in enumerate(f):
                    last = k == len(f) - 1
                    t = R.conv(t, sub[0], sub[1], not last, residual=S[i] if last else None)
                S[i] = t
        # branch 0: last conv2 -> x_0 and y_0
        R.lane = 0
        t, blk, res0 = last0
        c0 = mod.num_inchannels[0]
        y0 = dst0 if dst0 is not None else R.new(res0.n, res0.h, res0.w, c0)
        adds0 = [(W[0], 1)]
        if nout > 1:
            x0 = R.conv(t, blk.conv2, blk.bn2, True, residual=res0, addends=adds0, out2=y0)
        else:
            # nobody reads x_0 after the last module: both stores go to the same pixels, the
            # second one (y_0) stays
            x0 = R.conv(t, blk.conv2, blk.bn2, True, residual=res0, out=y0, addends=adds0, out2=y0)
        outs = [y0]
        # chains from x_0; their last conv adds the running sum and the upsampled terms, then ReLU
        for i in range(1, nout):
            f = mod.fuse_layers[i][0]
            t = x0
            for k, sub in enumerate(f):
                if k < len(f) - 1:
                    R.lane = 0
                    t = R.conv(t, sub[0], sub[1], True)
                else:
                    R.lane = i if par else 0
                    t = R.conv(t, sub[0], sub[1], True, residual=S[i],
                               addends=[(W[i], 1)] if i in W else None)
            outs.append(t)
        return outs

    # ---------------------------------------------------------------- forward
    def forward(self, x):
        """pose_higher_hrnet.py:637-686: (N,3,H,W) -> [(N,34,H/4,W/4), (N,17,H/2,W/2)]."""
        return self._run_plans(x, 32)

    def supports_flip_pair(self, x):
        """The fused flip-test batch needs the tensor-core stem (im2col) and N * H / 2 <= 65535 rows
        per plan replay."""
        return (self.conv_engine != L.ENGINE_FFMA and x.dim() == 4 and
                min(self.chunk_size, 2 * x.shape[0]) * (x.shape[2] // 2) <= 65535)

    def forward_flip_pair(self, x, via_half=False, borrow=False):
        """``forward(torch.cat((x, torch.flip(x, [3])), 0))`` of the flip test (upstream
        get_multi_stage_outputs callers, legacy/valid_ae_avg.py:176-185) without building that batch:
        the stem's im2col reads the mirrored half straight from ``x``.  Not part of the reference
        API (``TeacherPipeline`` uses it); bit-identical to the materialised batch.  ``via_half``:
        ``x`` is the float32 input of a ``network_to_half`` wrapper (rounded through fp16 first);
        ``borrow``: see ``_run_plans`` (outputs valid until the next forward of this module)."""
        return self._run_plans(x, 32, flip_pair=True, via_half=via_half, borrow=borrow)
