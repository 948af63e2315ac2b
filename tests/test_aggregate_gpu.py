"""GPU parity of the aggregation kernels against the CPU oracle (oracle/aggregate_ref.py =
torch CPU F.interpolate restatements).  Floating point: max|delta| <= 1e-5 * max|ref|
(both sides are fp32 bilinear arithmetic; only the FMA contraction order differs)."""
import pytest
import torch
import torch.nn.functional as F

from rtpe_b200 import inference
from oracle import aggregate_ref as A

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(got, ref):
    return ((got.cpu() - ref).abs().max() <= TOL * ref.abs().max()).item()


@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("hi,wi,ho,wo", [(40, 56, 480, 640), (80, 80, 555, 640), (32, 48, 64, 96),
                                         (20, 20, 7, 33)])
def test_bilinear_resize(cuda_device, ac, hi, wi, ho, wo):
    x = torch.randn(2, 5, hi, wi, generator=torch.Generator().manual_seed(hi + wo))
    got = inference.bilinear_resize(x.cuda(), (ho, wo), ac)
    ref = F.interpolate(x, (ho, wo), mode="bilinear", align_corners=ac)
    assert got.shape == ref.shape and close(got, ref)


def test_aggregate_intree(cuda_device):
    g = torch.Generator().manual_seed(0)
    y0 = torch.randn(2, 34, 40, 56, generator=g)
    y1 = torch.randn(2, 17, 80, 112, generator=g)
    det, tag = inference.aggregate_intree(y0.cuda(), y1.cuda(), (123, 200))
    rdet, rtag = A.aggregate_intree_ref(y0, y1, (123, 200))
    assert det.shape == rdet.shape and tag.shape == rtag.shape
    assert close(det, rdet) and close(tag, rtag)


@pytest.mark.parametrize("flip", [True, False])
@pytest.mark.parametrize("base", [(96, 64), (100, 72)])
def test_aggregate_single_scale(cuda_device, flip, base):
    g = torch.Generator().manual_seed(1)
    outs = [torch.randn(2, 34, 16, 24, generator=g), torch.randn(2, 17, 32, 48, generator=g)]
    outs_f = [torch.randn(2, 34, 16, 24, generator=g), torch.randn(2, 17, 32, 48, generator=g)] \
        if flip else None
    rdet, rtag = A.aggregate_flip_multiscale_ref([(1.0, outs, outs_f)], base)
    det, tag = inference.aggregate_flip_multiscale(
        [(1.0, [o.cuda() for o in outs], [o.cuda() for o in outs_f] if flip else None)], base)
    assert det.shape == rdet.shape and tag.shape == rtag.shape
    assert close(det, rdet) and close(tag, rtag)


def test_aggregate_multi_scale_flip(cuda_device):
    """scales visited descending (2.0, 1.0, 0.5); tags from scale 1.0 only; /3 at the end."""
    g = torch.Generator().manual_seed(2)
    per = []
    for s, (h4, w4) in ((2.0, (32, 48)), (1.0, (16, 24)), (0.5, (8, 12))):
        o = [torch.randn(1, 34, h4, w4, generator=g), torch.randn(1, 17, 2 * h4, 2 * w4, generator=g)]
        f = [torch.randn(1, 34, h4, w4, generator=g), torch.randn(1, 17, 2 * h4, 2 * w4, generator=g)]
        per.append((s, o, f))
    rdet, rtag = A.aggregate_flip_multiscale_ref(per, (96, 64))
    det, tag = inference.aggregate_flip_multiscale(
        [(s, [t.cuda() for t in o], [t.cuda() for t in f]) for s, o, f in per], (96, 64))
    assert tag.shape == rtag.shape == (1, 17, 64, 96, 2)
    assert close(det, rdet) and close(tag, rtag)


@pytest.mark.parametrize("flip", [True, False])
@pytest.mark.parametrize("h4,w4", [(40, 56), (44, 40), (10, 100)])
def test_aggregate_exact_x4_tma_path(cuda_device, flip, h4, w4):
    """base = 4 x the 1/4-resolution map and W4 % 4 == 0: the TMA-staged 2x -> 2x cascade kernel
    (tiles with aprons on every border, partial tiles in both directions)."""
    g = torch.Generator().manual_seed(3 + h4)
    outs = [torch.randn(3, 34, h4, w4, generator=g), torch.randn(3, 17, 2 * h4, 2 * w4, generator=g)]
    outs_f = [torch.randn(3, 34, h4, w4, generator=g), torch.randn(3, 17, 2 * h4, 2 * w4, generator=g)] \
        if flip else None
    base = (4 * w4, 4 * h4)
    rdet, rtag = A.aggregate_flip_multiscale_ref([(1.0, outs, outs_f)], base)
    det, tag = inference.aggregate_flip_multiscale(
        [(1.0, [o.cuda() for o in outs], [o.cuda() for o in outs_f] if flip else None)], base)
    assert det.shape == rdet.shape and tag.shape == rtag.shape
    assert close(det, rdet) and close(tag, rtag)
