"""GPU parity of the students' streaming NHWC kernels (brtpe_aux_run kinds 1, 4, 6-9) against the
torch operators the reference uses (rtpe/students.py:137-142, :197-201, :481-498, :835-846, :980-1018).
fp32 mode within 1e-6 of the tensor max; bf16 mode within one bf16 rounding of the fp32 result."""
import ctypes as C
import struct

import pytest
import torch
import torch.nn.functional as F

from rtpe_b200 import _lib as L

pytestmark = pytest.mark.gpu

MODES = [("fp32", torch.float32, 1e-6), ("bf16", torch.bfloat16, 8e-3)]


def _aux(kind, in0, in1, in2, out, ip):
    lib = L.load()
    arr = (C.c_int32 * len(ip))(*[int(v) for v in ip])
    L.check(lib.brtpe_aux_run(kind, L.ptr(in0), L.ptr(in1) if in1 is not None else None,
                              L.ptr(in2) if in2 is not None else None, L.ptr(out), arr, len(ip),
                              L.stream_ptr()), "brtpe_aux_run")
    torch.cuda.synchronize()


def _dt(mode):
    return L.DT_F32 if mode == "fp32" else L.DT_BF16


def _nhwc(x, ld, dtype):
    """NCHW float -> NHWC tensor with leading dimension ld (pad channels NaN: must not be read)."""
    n, c, h, w = x.shape
    t = torch.full((n, h, w, ld), float("nan"), dtype=dtype, device="cuda")
    t[..., :c] = x.permute(0, 2, 3, 1).to(dtype)
    return t


def _err(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max()).item()


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("c,ld", [(48, 48), (51, 64), (176, 176)])
def test_avgpool(cuda_device, mode, dtype, tol, c, ld):
    x = torch.randn(2, c, 12, 20, generator=torch.Generator().manual_seed(1)).cuda()
    xin = _nhwc(x, ld, dtype)
    out = torch.zeros((2, 6, 10, ld), dtype=dtype, device="cuda")
    _aux(1, xin, None, None, out, [_dt(mode), 2, 12, 20, c, ld, ld])
    ref = F.avg_pool2d(xin[..., :c].float().permute(0, 3, 1, 2), 3, 2, 1, count_include_pad=False)
    assert _err(out[..., :c].permute(0, 3, 1, 2), ref) <= tol


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("c,cz", [(48, 48), (83, 96), (163, 176)])
def test_cam_mix_zero_fills_pad_channels(cuda_device, mode, dtype, tol, c, cz):
    g = torch.Generator().manual_seed(2)
    res, hdc = torch.randn(2, c, 9, 7, generator=g).cuda(), torch.randn(2, c, 9, 7, generator=g).cuda()
    gate = torch.rand(2, c, generator=g).cuda()
    a, b = _nhwc(res, cz, dtype), _nhwc(hdc, cz, dtype)
    a[..., c:] = 1.0                                       # finite garbage in the pad channels
    b[..., c:] = -2.0
    out = torch.full((2, 9, 7, cz), float("nan"), dtype=dtype, device="cuda")
    _aux(4, a, b, gate, out, [_dt(mode), 2, 63, c, cz, cz, cz, cz if cz > c else 0])
    ref = F.relu(a[..., :c].float() + b[..., :c].float() * gate[:, None, None, :])
    assert _err(out[..., :c], ref) <= tol
    if cz > c:
        assert (out[..., c:] == 0).all()


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("align", [True, False])
def test_resize_nhwc(cuda_device, mode, dtype, tol, align):
    x = torch.randn(2, 40, 9, 13, generator=torch.Generator().manual_seed(3)).cuda()
    xin = _nhwc(x, 48, dtype)
    out = torch.zeros((2, 21, 17, 64), dtype=dtype, device="cuda")
    _aux(6, xin, None, None, out, [_dt(mode), 2, 9, 13, 40, 48, 64, 21, 17, int(align), 16])
    ref = F.interpolate(xin[..., :40].float().permute(0, 3, 1, 2), (21, 17), mode="bilinear",
                        align_corners=align)
    assert _err(out[..., 16:56].permute(0, 3, 1, 2), ref) <= tol
    assert (out[..., :16] == 0).all() and (out[..., 56:] == 0).all()
    # same size + align_corners=True is the identity (MultistageStudent's per-stage resizes)
    same = torch.zeros((2, 9, 13, 48), dtype=dtype, device="cuda")
    _aux(6, xin, None, None, same, [_dt(mode), 2, 9, 13, 40, 48, 48, 9, 13, 1, 0])
    assert torch.equal(same[..., :40], xin[..., :40])


@pytest.mark.parametrize("mode,dtype,tol", MODES)
def test_image_to_nhwc_resize_and_space_to_depth(cuda_device, mode, dtype, tol):
    img = torch.randn(2, 3, 32, 48, generator=torch.Generator().manual_seed(4)).cuda()
    out = torch.full((2, 8, 12, 64), 7.0, dtype=dtype, device="cuda")
    _aux(7, img, None, None, out, [_dt(mode), 2, 32, 48, 3, 0, 64, 8, 12, 48, 16, 0])
    ref = F.interpolate(img, (8, 12), mode="bilinear")          # students.py:986-987
    assert _err(out[..., 48:51].permute(0, 3, 1, 2), ref) <= tol
    assert (out[..., 51:] == 0).all() and (out[..., :48] == 7.0).all()
    s2d = torch.full((2, 16, 24, 16), float("nan"), dtype=dtype, device="cuda")
    _aux(7, img, None, None, s2d, [_dt(mode), 2, 32, 48, 3, 0, 16, 16, 24, 0, 16, 1])
    for ry in (0, 1):
        for rx in (0, 1):
            b = (ry * 2 + rx) * 3
            want = img[:, :, ry::2, rx::2].to(dtype)
            assert torch.equal(s2d[..., b:b + 3].permute(0, 3, 1, 2), want)
    assert (s2d[..., 12:] == 0).all()


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("c", [64, 20])
def test_space_to_depth_nhwc(cuda_device, mode, dtype, tol, c):
    x = torch.randn(2, c, 12, 20, generator=torch.Generator().manual_seed(5)).cuda()
    xin = _nhwc(x, 64, dtype)
    out = torch.full((2, 6, 10, 4 * c), float("nan"), dtype=dtype, device="cuda")
    _aux(8, xin, None, None, out, [_dt(mode), 2, 12, 20, c, 64, 4 * c])
    for ry in (0, 1):
        for rx in (0, 1):
            b = (ry * 2 + rx) * c
            assert torch.equal(out[..., b:b + c], xin[:, ry::2, rx::2, :c])


@pytest.mark.parametrize("mode,dtype,tol", MODES)
@pytest.mark.parametrize("div", [1.0, 20.0])
def test_attention_product(cuda_device, mode, dtype, tol, div):
    g = torch.Generator().manual_seed(6)
    att = (torch.randn(2, 1, 9, 7, generator=g) * 4).cuda()
    x = torch.randn(2, 96, 9, 7, generator=g).cuda()
    a, xin = _nhwc(att, 16, dtype), _nhwc(x, 96, dtype)
    att_out = torch.zeros((2, 1, 9, 7), dtype=torch.float32, device="cuda")
    out = torch.zeros((2, 9, 7, 176), dtype=dtype, device="cuda")
    (bits,) = struct.unpack("<i", struct.pack("<f", div))
    _aux(9, a, xin, att_out, out, [_dt(mode), 2, 63, 96, 16, 96, 176, bits])
    s = torch.sigmoid(a[..., 0].float() / div)
    assert _err(att_out[:, 0], s) <= 1e-6
    assert _err(out[..., :96], xin.float() * s[..., None]) <= tol
    assert (out[..., 96:] == 0).all()


def test_aux_rejects_bad_arguments(cuda_device):
    lib = L.load()
    x = torch.zeros(64, device="cuda")
    bad = (C.c_int32 * 11)(L.DT_F32, 1, 4, 4, 8, 4, 8, 2, 2, 1, 0)        # C > in_ld
    assert lib.brtpe_aux_run(6, L.ptr(x), None, None, L.ptr(x), bad, 11, L.stream_ptr()) < 0
    assert b"resize_nhwc" in lib.brtpe_last_error()
    unk = (C.c_int32 * 4)(L.DT_F32, 1, 1, 1)
    assert lib.brtpe_aux_run(42, L.ptr(x), None, None, L.ptr(x), unk, 4, L.stream_ptr()) < 0


@pytest.mark.parametrize("c,h,w,shifts,relu", [
    (48, 16, 160, (0, 1, 2, 3), True),      # HRNet branch 0: identity + three up-sampled terms
    (96, 8, 40, (0, 0, 1), True),
    (192, 4, 20, (0, 0), False),
    (384, 2, 20, (0, 0, 0, 0), True),
    (48, 6, 50, (0, 1), True),              # width that is not a multiple of the pixel block
    (64, 3, 7, (0,), False),
    (176, 5, 9, (0, 0), True),
])
def test_fuse_sum_bf16_bit_exact(cuda_device, c, h, w, shifts, relu):
    """HighResolutionModule's cross-resolution sum (pose_higher_hrnet.py:245-254): nearest
    up-sampling by 2^shift, fp32 left-to-right sum, ReLU, one bf16 rounding -- bit-exact."""
    lib = L.load()
    g = torch.Generator().manual_seed(7)
    n = 3
    hh = h * 8                                                    # every shift divides H and W
    ww = (w + 7) // 8 * 8 if max(shifts) > 0 else w
    terms, ref = [], None
    for k, sh in enumerate(shifts):
        ld = c + 16 * (k % 2)                                     # padded leading dimensions too
        t = torch.full((n, hh >> sh, ww >> sh, ld), float("nan"), dtype=torch.bfloat16, device="cuda")
        t[..., :c] = torch.randn((n, hh >> sh, ww >> sh, c), generator=g).to("cuda", torch.bfloat16)
        terms.append(t)
        up = t[..., :c].float().repeat_interleave(1 << sh, 1).repeat_interleave(1 << sh, 2)
        ref = up if ref is None else ref + up
    if relu:
        ref = torch.relu(ref)
    out = torch.zeros((n, hh, ww, c + 16), dtype=torch.bfloat16, device="cuda")
    nt = len(terms)
    tp = (C.c_void_p * nt)(*[t.data_ptr() for t in terms])
    sh = (C.c_int32 * nt)(*shifts)
    ld = (C.c_int32 * nt)(*[t.shape[3] for t in terms])
    L.check(lib.brtpe_fuse_sum(L.DT_BF16, nt, tp, sh, ld, n, hh, ww, c, L.ptr(out), out.shape[3],
                               int(relu), L.stream_ptr()), "brtpe_fuse_sum")
    torch.cuda.synchronize()
    assert torch.equal(out[..., :c], ref.to(torch.bfloat16))
    assert (out[..., c:] == 0).all()
